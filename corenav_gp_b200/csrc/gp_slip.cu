// Host orchestration of the two entry points that sit directly behind the node callback
// core_navigation/script/gp_slip_node.py:16-63:
//
//   cngp_optimize_batch   m.optimize() (gp_slip_node.py:36) for B windows at once.  GPy hands the objective to paramz'
//                         "lbfgsb" optimiser = scipy.optimize.fmin_l_bfgs_b (L-BFGS-B 3.0, m = 10, factr = 1e7,
//                         pgtol = 1e-5, maxfun = maxiter = 1000) on Logexp (softplus) transformed positives.  With no
//                         bounds L-BFGS-B is L-BFGS with the More-Thuente line search (MINPACK-2 dcsrch/dcstep,
//                         ftol 1e-3, gtol 0.9, xtol 0.1, first step 1/||g||, at most 20 trial points per search).  That
//                         published algorithm is restated below as an explicit per-window state machine so that all B
//                         optimisers advance in lock step: every round evaluates ONE trial point per still-active
//                         window in a single batched GPU launch (cngp_lml_grad_windows).  Objective, gradient and
//                         all O(N^3) work are on the GPU; the host only runs the O(m P) two-loop recursions.
//   cngp_gp_slip_batch    the whole callback (rows a1-a7): 90 % train split, [optional fit], prediction grid
//                         arange(min, max + horizon, 1), predict, keep [n:], sigma = 2 sqrt(var) (fused into the
//                         variance kernel's epilogue).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "cngp_internal.h"
#include "lbfgsb_host.h"

using namespace cngp_host;

namespace {

// grow-only device buffers of the context (slots 30..36 belong to the optimiser): no cudaMalloc / cudaFree per call
enum { SLOT_OPT_X = 30, SLOT_OPT_Y, SLOT_OPT_TH, SLOT_OPT_MAP, SLOT_OPT_LML, SLOT_OPT_GRAD, SLOT_OPT_ST };

}  // namespace

extern "C" int cngp_optimize_batch(cngp_ctx* ctx, const cngp_kernel* kernel, const double* theta0, int64_t theta0_stride,
                                   const double* x, const double* y, int64_t B, int32_t N, int32_t max_iters,
                                   double* theta_out, double* lml_out, int32_t* iters_out) {
  if (!ctx) return CNGP_ERR_INVALID;
  if (!kernel || !x || !y || !theta_out || B < 0 || N <= 0) return cngp_set_error(ctx, CNGP_ERR_INVALID, "optimize: bad argument");
  if (B == 0) return CNGP_OK;
  cngp_kernel kk = *kernel;
  if (cngp_kernel_finalize(&kk) != CNGP_OK) return cngp_set_error(ctx, CNGP_ERR_INVALID, "optimize: invalid kernel expression");
  const int P = kk.n_params + 1;
  if (max_iters <= 0) max_iters = 1000;

  std::vector<Optimizer> opt((size_t)B);
  std::vector<double> ones((size_t)P, 1.0);   // GPy initialises every hyper-parameter and the noise at 1.0
  for (int64_t b = 0; b < B; ++b)
    opt[b].init(theta0 ? theta0 + (theta0_stride ? b * theta0_stride : 0) : ones.data(), P, max_iters);

  // Everything below is ordered on the context's stream on the context's device: the uploads are cudaMemcpyAsync on
  // that stream (the kernels run there, and it is a non-blocking stream, so legacy-stream copies would not order them).
  if (cudaSetDevice(cngp_ctx_device(ctx)) != cudaSuccess) return cngp_set_error(ctx, CNGP_ERR_CUDA, "optimize: cudaSetDevice failed");
  cudaStream_t s = cngp_ctx_stream(ctx);
  struct { void* p; } dx, dy, dth, dmap, dlml, dgrad, dst;
  const size_t xy = sizeof(double) * (size_t)B * N;
  dx.p = cngp_ctx_buf(ctx, SLOT_OPT_X, xy);
  dy.p = cngp_ctx_buf(ctx, SLOT_OPT_Y, xy);
  dth.p = cngp_ctx_buf(ctx, SLOT_OPT_TH, sizeof(double) * B * P);
  dmap.p = cngp_ctx_buf(ctx, SLOT_OPT_MAP, sizeof(int) * B);
  dlml.p = cngp_ctx_buf(ctx, SLOT_OPT_LML, sizeof(double) * B);
  dgrad.p = cngp_ctx_buf(ctx, SLOT_OPT_GRAD, sizeof(double) * B * P);
  dst.p = cngp_ctx_buf(ctx, SLOT_OPT_ST, sizeof(int) * B);
  if (!dx.p || !dy.p || !dth.p || !dmap.p || !dlml.p || !dgrad.p || !dst.p)
    return cngp_set_error(ctx, CNGP_ERR_NOMEM, "optimize: device allocation failed");
  if (cudaMemcpyAsync(dx.p, x, xy, cudaMemcpyHostToDevice, s) != cudaSuccess ||
      cudaMemcpyAsync(dy.p, y, xy, cudaMemcpyHostToDevice, s) != cudaSuccess)
    return cngp_set_error(ctx, CNGP_ERR_CUDA, "optimize: upload failed");

  std::vector<int> active;
  std::vector<double> th((size_t)B * P), lml((size_t)B), grad((size_t)B * P), gz((size_t)P);
  std::vector<int> status((size_t)B);
  for (;;) {
    active.clear();
    for (int64_t b = 0; b < B; ++b)
      if (!opt[b].done) active.push_back((int)b);
    if (active.empty()) break;
    const size_t na = active.size();
    for (size_t a = 0; a < na; ++a)
      for (int i = 0; i < P; ++i) th[a * P + i] = softplus(opt[active[a]].zt[i]);
    if (cudaMemcpyAsync(dth.p, th.data(), sizeof(double) * na * P, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaMemcpyAsync(dmap.p, active.data(), sizeof(int) * na, cudaMemcpyHostToDevice, s) != cudaSuccess)
      return cngp_set_error(ctx, CNGP_ERR_CUDA, "optimize: upload failed");
    int rc = cngp_lml_grad_impl(ctx, &kk, (const double*)dth.p, (int64_t)na, (const double*)dx.p, (const double*)dy.p, B, N,
                                (double*)dlml.p, (double*)dgrad.p, (int*)dst.p, CNGP_MEM_DEVICE, (const int*)dmap.p);
    if (rc) return rc;
    if (cudaMemcpyAsync(lml.data(), dlml.p, sizeof(double) * na, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(grad.data(), dgrad.p, sizeof(double) * na * P, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(status.data(), dst.p, sizeof(int) * na, cudaMemcpyDeviceToHost, s) != cudaSuccess)
      return cngp_set_error(ctx, CNGP_ERR_CUDA, "optimize: download failed");
    if ((rc = cngp_sync(ctx))) return rc;     // the host state machines need the values; th / active are re-written next
    for (size_t a = 0; a < na; ++a) {
      Optimizer& o = opt[active[a]];
      double fv;
      if (status[a] < 0 || !std::isfinite(lml[a])) {   // not positive definite: a wall, as in the CPU path
        fv = 1e300;
        for (int i = 0; i < P; ++i) gz[i] = 0.0;
      } else {
        fv = -lml[a];
        for (int i = 0; i < P; ++i) gz[i] = -grad[a * P + i] * softplus_gradfactor(th[a * P + i]);
      }
      o.feed(fv, gz.data());
    }
  }
  for (int64_t b = 0; b < B; ++b) {
    for (int i = 0; i < P; ++i) theta_out[b * P + i] = softplus(opt[b].z[i]);
    if (lml_out) lml_out[b] = -opt[b].f;
    if (iters_out) iters_out[b] = opt[b].nfev;
  }
  return CNGP_OK;
}

extern "C" int cngp_gp_slip_batch(cngp_ctx* ctx, const cngp_kernel* kernel, const double* theta, int64_t theta_stride,
                                  const double* time_array, const double* slip_array, int64_t B, int32_t n,
                                  int32_t horizon, int32_t m_cap, double* mean, double* sigma, int32_t* m_out,
                                  int32_t* status) {
  if (!ctx) return CNGP_ERR_INVALID;
  if (!kernel || !time_array || !slip_array || !mean || !sigma || !m_out || B < 0 || n <= 0 || horizon < 0)
    return cngp_set_error(ctx, CNGP_ERR_INVALID, "gp_slip: bad argument");
  *m_out = 0;
  if (B == 0) return CNGP_OK;
  if (cudaSetDevice(cngp_ctx_device(ctx)) != cudaSuccess) return cngp_set_error(ctx, CNGP_ERR_CUDA, "gp_slip: cudaSetDevice failed");
  cngp_kernel kk = *kernel;
  if (cngp_kernel_finalize(&kk) != CNGP_OK) return cngp_set_error(ctx, CNGP_ERR_INVALID, "gp_slip: invalid kernel expression");
  const int P = kk.n_params + 1;
  // gp_slip_node.py:27-30: per = 0.9; train = first int(per * len(X)) samples (double product, then truncation)
  const int ntr = (int)(0.9 * (double)n);
  if (ntr < 1) return cngp_set_error(ctx, CNGP_ERR_INVALID, "gp_slip: fewer than 2 samples");
  if (ntr > CNGP_MAX_N) return cngp_set_error(ctx, CNGP_ERR_UNSUPPORTED, "gp_slip: int(0.9 n) exceeds CNGP_MAX_N");
  // gp_slip_node.py:45: X_ = arange(X.min(), X.max() + horizon, 1) -> start + k, ceil(stop - start) points
  std::vector<double> t0((size_t)B);
  int64_t len = -1;
  for (int64_t b = 0; b < B; ++b) {
    const double* t = time_array + b * n;
    double lo = t[0], hi = t[0];
    for (int i = 1; i < n; ++i) { lo = std::min(lo, t[i]); hi = std::max(hi, t[i]); }
    const int64_t l = (int64_t)std::ceil((hi + (double)horizon) - lo);
    if (len >= 0 && l != len)
      return cngp_set_error(ctx, CNGP_ERR_UNSUPPORTED, "gp_slip: windows of one batch must span the same number of grid points");
    len = l;
    t0[b] = lo;
  }
  const int64_t M = len - n;                      // gp_slip_node.py:59-61: entries [len(X):] are published
  if (M <= 0) return CNGP_OK;
  if (M > m_cap) return cngp_set_error(ctx, CNGP_ERR_INVALID, "gp_slip: m_cap too small");

  std::vector<double> xtr((size_t)B * ntr), ytr((size_t)B * ntr), grid((size_t)B * M);
  for (int64_t b = 0; b < B; ++b) {
    memcpy(&xtr[(size_t)b * ntr], time_array + b * n, sizeof(double) * ntr);
    memcpy(&ytr[(size_t)b * ntr], slip_array + b * n, sizeof(double) * ntr);
    for (int64_t k = 0; k < M; ++k) grid[(size_t)b * M + k] = t0[b] + (double)(n + k);
  }
  std::vector<double> fitted;
  const double* th = theta;
  int64_t th_stride = theta_stride;
  if (!theta) {                                   // gp_slip_node.py:35-36: all-ones start, then m.optimize()
    fitted.resize((size_t)B * P);
    const int rc = cngp_optimize_batch(ctx, &kk, nullptr, 0, xtr.data(), ytr.data(), B, ntr, 1000, fitted.data(), nullptr, nullptr);
    if (rc) return rc;
    th = fitted.data();
    th_stride = P;
  }
  std::vector<double> mu((size_t)B * M), sg((size_t)B * M);
  const int rc = cngp_predict_impl(ctx, &kk, th, th_stride, xtr.data(), ytr.data(), grid.data(), M, B, ntr, (int32_t)M,
                                   mu.data(), sg.data(), nullptr, status, CNGP_MEM_HOST, /*sigma_mode=*/1);
  if (rc) return rc;
  for (int64_t b = 0; b < B; ++b) {
    memcpy(mean + b * m_cap, &mu[(size_t)b * M], sizeof(double) * M);
    memcpy(sigma + b * m_cap, &sg[(size_t)b * M], sizeof(double) * M);
  }
  *m_out = (int32_t)M;
  return CNGP_OK;
}
