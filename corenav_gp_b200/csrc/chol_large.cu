// placeholder - replaced below in this round
#include "../../include/cngp.h"
extern "C" int cngp_chol_large(cngp_ctx*, const cngp_kernel*, const double*, const double*, const double*, int64_t,
                               double*, double*, double*, double*, int32_t) { return CNGP_ERR_UNSUPPORTED; }
