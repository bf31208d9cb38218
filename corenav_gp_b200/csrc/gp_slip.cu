// placeholder - replaced below in this round
#include "../../include/cngp.h"
extern "C" int cngp_optimize_batch(cngp_ctx*, const cngp_kernel*, const double*, int64_t, const double*, const double*,
                                   int64_t, int32_t, int32_t, double*, double*, int32_t*) { return CNGP_ERR_UNSUPPORTED; }
extern "C" int cngp_gp_slip_batch(cngp_ctx*, const cngp_kernel*, const double*, int64_t, const double*, const double*,
                                  int64_t, int32_t, int32_t, int32_t, double*, double*, int32_t*, int32_t*) { return CNGP_ERR_UNSUPPORTED; }
