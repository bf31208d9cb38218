// Internal (non-ABI) entry points shared between the translation units of libcngp.
#pragma once
#include "../../include/cngp.h"

// cngp_predict_batch with the output transform of the node callback: sigma_mode = 1 writes 2 sqrt(var) into `var`.
int cngp_predict_impl(cngp_ctx* ctx, const cngp_kernel* kernel, const double* theta, int64_t theta_stride,
                      const double* x, const double* y, const double* xstar, int64_t xstar_stride, int64_t B, int32_t N,
                      int32_t M, double* mean, double* var, double* lml, int32_t* status, int32_t mem, int sigma_mode);

// cngp_lml_grad_batch (win_map == nullptr) / cngp_lml_grad_windows (win_map given).
int cngp_lml_grad_impl(cngp_ctx* ctx, const cngp_kernel* kernel, const double* theta, int64_t C, const double* x,
                       const double* y, int64_t B, int32_t N, double* lml, double* grad, int32_t* status, int32_t mem,
                       const int32_t* win_map, const int32_t* skip = nullptr);   // skip[p] != 0: problem p is left untouched

// error text of a context (cngp_api.cu)
int cngp_set_error(cngp_ctx* ctx, int code, const char* text);

// ---- context internals for the other translation units (chol_large.cu) ----
#ifdef __CUDACC__
#include <cuda_runtime.h>
namespace cngp { struct KProg; }
cudaStream_t cngp_ctx_stream(cngp_ctx* ctx);
int cngp_ctx_device(cngp_ctx* ctx);
void cngp_ctx_begin(cngp_ctx* ctx, int prof_id);   // counts one kernel launch (+ opens a timing span when profiling)
void cngp_ctx_end(cngp_ctx* ctx);
void* cngp_ctx_buf(cngp_ctx* ctx, size_t slot, size_t bytes);   // grow-only device buffer `slot`
int cngp_build_kprog(const cngp_kernel* k, cngp::KProg* kp);
int cngp_sm_count(void);
#endif
