/* TEST DOUBLE of the part of the C ABI that GpPredictor uses, backed by the C oracle (oracle/stop_oracle.c), so that the
 * host logic of corenav_gp_b200/host/gp_predictor.cpp can be tested on a machine without a GPU.  Lives under tests/:
 * the product never links this. */
#include <string.h>
#include "../../include/cngp.h"

typedef struct {     /* oracle/stop_oracle.c stop_cfg */
  double v_nom, floor_a, floor_b, track, scale, thresh;
  int ratio, fix_h_packing, trig_mode, pad_;
  double init_llh[3], init_ecef[3];
} stop_cfg;
void stop_oracle_default_cfg(stop_cfg* c);
void stop_oracle_llh_to_enu(double phi, double lambda, double h, const stop_cfg* c, double enu[3]);
int stop_oracle_lookahead_ex(const double* mean, const double* sigma, int M, const double* Pvec, const double* Qvec,
                             const double* STMvec, const double* Hvec, const double* pos, const stop_cfg* c,
                             int* triggered, int* i_stop, int* step_stop, double* xy_err, double* xy_trace, double* P_out,
                             double* K_out, double* R_out);

/* field-by-field: the two structs do not share a layout (the oracle carries its trigonometry switch) */
static stop_cfg to_oracle(const cngp_stop_config* c) {
  stop_cfg o;
  o.v_nom = c->v_nom; o.floor_a = c->floor_a; o.floor_b = c->floor_b; o.track = c->track; o.scale = c->scale;
  o.thresh = c->thresh; o.ratio = c->ratio; o.fix_h_packing = c->fix_h_packing;
  o.trig_mode = 1;      /* the deterministic sin / cos the CUDA kernels use */
  o.pad_ = 0;
  memcpy(o.init_llh, c->init_llh, sizeof o.init_llh);
  memcpy(o.init_ecef, c->init_ecef, sizeof o.init_ecef);
  return o;
}

struct cngp_ctx { int dummy; };
static struct cngp_ctx g_ctx;

int cngp_create(const cngp_config* cfg, cngp_ctx** out) { (void)cfg; *out = &g_ctx; return CNGP_OK; }
void cngp_destroy(cngp_ctx* ctx) { (void)ctx; }
const char* cngp_last_error(cngp_ctx* ctx) { (void)ctx; return ""; }
void cngp_default_stop_config(cngp_stop_config* c) {
  stop_cfg o;
  stop_oracle_default_cfg(&o);
  c->v_nom = o.v_nom; c->floor_a = o.floor_a; c->floor_b = o.floor_b; c->track = o.track; c->scale = o.scale;
  c->thresh = o.thresh; c->ratio = o.ratio; c->fix_h_packing = o.fix_h_packing;
  memcpy(c->init_llh, o.init_llh, sizeof o.init_llh);
  memcpy(c->init_ecef, o.init_ecef, sizeof o.init_ecef);
}
int cngp_llh_to_enu(cngp_ctx* ctx, const double* llh, int64_t n, const cngp_stop_config* cfg, double* enu, int32_t mem) {
  (void)ctx; (void)mem;
  const stop_cfg o = to_oracle(cfg);
  for (int64_t i = 0; i < n; ++i) stop_oracle_llh_to_enu(llh[3 * i], llh[3 * i + 1], llh[3 * i + 2], &o, enu + 3 * i);
  return CNGP_OK;
}
int cngp_zupt_lookahead_batch_ex(cngp_ctx* ctx, const double* mean, const double* sigma, int64_t B, int32_t M,
                                 const double* P, const double* Q, const double* STM, const double* Hvec,
                                 const double* pos, int32_t per_window, const cngp_stop_config* cfg, int32_t* triggered,
                                 int32_t* i_stop, int32_t* step_stop, double* xy_err, double* P_final, double* K_final,
                                 double* R_final, int32_t mem) {
  (void)ctx; (void)mem;
  const stop_cfg o = to_oracle(cfg);
  for (int64_t b = 0; b < B; ++b) {
    int step_tmp; double xy_tmp;
    if (K_final) memset(K_final + b * 60, 0, 60 * sizeof(double));
    if (R_final) memset(R_final + b * 16, 0, 16 * sizeof(double));
    stop_oracle_lookahead_ex(mean + b * M, sigma + b * M, M, P + ((per_window & CNGP_PERWIN_P) ? b * 225 : 0),
                             Q + ((per_window & CNGP_PERWIN_Q) ? b * 225 : 0),
                             STM + ((per_window & CNGP_PERWIN_STM) ? b * 225 : 0),
                             Hvec + ((per_window & CNGP_PERWIN_H) ? b * 60 : 0),
                             pos + ((per_window & CNGP_PERWIN_POS) ? b * 3 : 0), &o, triggered + b, i_stop + b,
                             step_stop ? step_stop + b : &step_tmp, xy_err ? xy_err + b : &xy_tmp, 0,
                             P_final ? P_final + b * 225 : 0, K_final ? K_final + b * 60 : 0,
                             R_final ? R_final + b * 16 : 0);
  }
  return CNGP_OK;
}
int cngp_zupt_lookahead_batch(cngp_ctx* ctx, const double* mean, const double* sigma, int64_t B, int32_t M,
                              const double* P, const double* Q, const double* STM, const double* Hvec, const double* pos,
                              int32_t per_window, const cngp_stop_config* cfg, int32_t* triggered, int32_t* i_stop,
                              int32_t* step_stop, double* xy_err, int32_t mem) {
  return cngp_zupt_lookahead_batch_ex(ctx, mean, sigma, B, M, P, Q, STM, Hvec, pos, per_window, cfg, triggered, i_stop,
                                      step_stop, xy_err, 0, 0, 0, mem);
}
