// gp_slip_predict: the reference's GP node callback as a C++ function over the C ABI.
// Mirrors core_navigation/script/gp_slip_node.py:16-63; all arithmetic is cngp_gp_slip_batch (libcngp.so).
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gp_predictor_b200.hpp"

core_nav::GP_Output gp_slip_predict(cngp_ctx* ctx, const core_nav::GP_Input& data, const char* kernel, const double* theta,
                                    int horizon) {
  if (data.time_array.size() != data.slip_array.size() || data.time_array.empty())
    throw std::runtime_error("gp_slip_predict: time_array and slip_array must be non-empty and of equal length");
  cngp_kernel k;
  if (cngp_kernel_parse(kernel, &k) != CNGP_OK) throw std::runtime_error(std::string("gp_slip_predict: bad kernel expression ") + kernel);
  const int n = (int)data.time_array.size();
  double lo = data.time_array[0], hi = data.time_array[0];
  for (double t : data.time_array) { lo = std::fmin(lo, t); hi = std::fmax(hi, t); }
  const int m_cap = (int)std::ceil(hi - lo) + horizon + 2;       // len(arange(min, max + horizon, 1)) - n, with slack
  core_nav::GP_Output out;
  out.header = data.header;
  out.mean.resize(m_cap);
  out.sigma.resize(m_cap);
  int32_t m = 0, status = 0;
  const int rc = cngp_gp_slip_batch(ctx, &k, theta, 0, data.time_array.data(), data.slip_array.data(), 1, n, horizon, m_cap,
                                    out.mean.data(), out.sigma.data(), &m, &status);
  if (rc != CNGP_OK) throw std::runtime_error(std::string("gp_slip_predict: ") + cngp_last_error(ctx));
  if (status < 0) throw std::runtime_error("gp_slip_predict: kernel matrix not positive definite (numpy.linalg.LinAlgError in the reference)");
  out.mean.resize(m);
  out.sigma.resize(m);
  return out;
}
