/* TEST DOUBLE of the part of the C ABI that GpPredictor uses, backed by the C oracle (oracle/stop_oracle.c), so that the
 * host logic of corenav_gp_b200/host/gp_predictor.cpp can be tested on a machine without a GPU.  Lives under tests/:
 * the product never links this. */
#include <string.h>
#include "../../include/cngp.h"

typedef struct {
  double v_nom, floor_a, floor_b, track, scale, thresh;
  int ratio, fix_h_packing;
  double init_llh[3], init_ecef[3];
} stop_cfg;
void stop_oracle_default_cfg(stop_cfg* c);
void stop_oracle_llh_to_enu(double phi, double lambda, double h, const stop_cfg* c, double enu[3]);
int stop_oracle_lookahead_batch(const double* mean, const double* sigma, int B, int M, const double* Pvec,
                                const double* Qvec, const double* STMvec, const double* Hvec, const double* pos,
                                int per_window, const stop_cfg* c, int* triggered, int* i_stop, int* step_stop,
                                double* xy_err);

struct cngp_ctx { int dummy; };
static struct cngp_ctx g_ctx;

int cngp_create(const cngp_config* cfg, cngp_ctx** out) { (void)cfg; *out = &g_ctx; return CNGP_OK; }
void cngp_destroy(cngp_ctx* ctx) { (void)ctx; }
const char* cngp_last_error(cngp_ctx* ctx) { (void)ctx; return ""; }
void cngp_default_stop_config(cngp_stop_config* c) { stop_oracle_default_cfg((stop_cfg*)c); }
int cngp_llh_to_enu(cngp_ctx* ctx, const double* llh, int64_t n, const cngp_stop_config* cfg, double* enu, int32_t mem) {
  (void)ctx; (void)mem;
  for (int64_t i = 0; i < n; ++i) stop_oracle_llh_to_enu(llh[3 * i], llh[3 * i + 1], llh[3 * i + 2], (const stop_cfg*)cfg, enu + 3 * i);
  return CNGP_OK;
}
int cngp_zupt_lookahead_batch(cngp_ctx* ctx, const double* mean, const double* sigma, int64_t B, int32_t M,
                              const double* P, const double* Q, const double* STM, const double* Hvec, const double* pos,
                              int32_t per_window, const cngp_stop_config* cfg, int32_t* triggered, int32_t* i_stop,
                              int32_t* step_stop, double* xy_err, int32_t mem) {
  (void)ctx; (void)mem;
  int step_tmp[1]; double xy_tmp[1];
  if (B != 1 && (!step_stop || !xy_err)) return CNGP_ERR_INVALID;
  return stop_oracle_lookahead_batch(mean, sigma, (int)B, M, P, Q, STM, Hvec, pos, per_window, (const stop_cfg*)cfg,
                                     triggered, i_stop, step_stop ? step_stop : step_tmp, xy_err ? xy_err : xy_tmp);
}
