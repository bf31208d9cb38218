// CPU test shim: drives the product's L-BFGS-B state machine (corenav_gp_b200/csrc/lbfgsb_host.h) with a caller-supplied
// objective, exactly as cngp_optimize_batch drives it with the GPU objective.  Built by tests/test_host_lbfgsb.py.
#include <vector>

#include "../../corenav_gp_b200/csrc/lbfgsb_host.h"

extern "C" int lbfgsb_shim_minimize(int P, const double* theta0, int max_iters,
                                    void (*fg)(const double* theta, double* neg_lml, double* neg_grad_theta),
                                    double* theta_out, double* f_out, int* nfev_out, int* iters_out) {
  using namespace cngp_host;
  Optimizer o;
  o.init(theta0, P, max_iters);
  std::vector<double> th(P), gth(P), gz(P);
  while (!o.done) {
    for (int i = 0; i < P; ++i) th[i] = softplus(o.zt[i]);
    double f = 0.0;
    fg(th.data(), &f, gth.data());
    for (int i = 0; i < P; ++i) gz[i] = gth[i] * softplus_gradfactor(th[i]);
    o.feed(f, gz.data());
  }
  for (int i = 0; i < P; ++i) theta_out[i] = softplus(o.z[i]);
  *f_out = o.f;
  *nfev_out = o.nfev;
  *iters_out = o.iters;
  return 0;
}
