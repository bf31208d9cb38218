// C entry points around the UNMODIFIED reference class GpPredictor (compiled from
// /root/reference/gp_predictor/src/gp_predictor.cpp with the stand-in headers of oracle/ref_stubs/).
//
// TEST INFRASTRUCTURE ONLY: built into oracle/_ref/libgp_predictor_ref.so by oracle/ref_build.py; loaded by tests/ to
// pin oracle/stop_oracle.c and the CUDA look-ahead kernel to the reference's own code.  Nothing under
// corenav_gp_b200/ links or loads it.
//
// What is injected, and why it does not change the reference's arithmetic:
//   * the SetStopping response (P, Q, STM, Hvec as CoreNav::setStopping_ packs them, CoreNav.cpp:652-676; PosData),
//   * a constant clock, so stop_cmd = gp_arrived + i/10.0 - now = i/10.0 exactly (gp_predictor.cpp:107-116) and the
//     reference's `i` at the trigger is recovered from the message it publishes,
//   * init_llh / init_ecef: public members the reference reads in llh_to_enu but never sets (LoadParameters is not
//     called anywhere); here they are set through the reference's own LoadParameters().
#include <cmath>
#include <cstring>
#include <gp_predictor/gp_predictor.h>

extern "C" {

// One GPCallBack.  mean/sigma: GP_Output arrays [M].  Outputs: n_published (0 or 1), stop_cmd (the published
// std_msgs/Float64, NaN if none), xy_err_last (member xy_errSlip after the callback), n_steps = look-ahead steps
// executed (one ROS_ERROR_THROTTLE each), xy_trace[cap_trace] (optional) the per-step xy error, P_out[225] /
// K_out[60] / R_out[16] (optional, row-major) = members P_pred, K_pred, R_IP after the callback.
int ref_gp_callback(const double* mean, const double* sigma, int M, const double* Pvec, const double* Qvec,
                    const double* STMvec, const double* Hvec, const double* pos, const double* init_llh,
                    const double* init_ecef, double clock_arrive, double clock_later, int* n_published,
                    double* stop_cmd, double* xy_err_last, int* n_steps, double* xy_trace, int cap_trace, double* P_out,
                    double* K_out, double* R_out) {
  ros::stub::Bench& b = ros::stub::bench();
  b.published.clear();
  b.throttle_values.clear();
  b.params = {{"init_llh/x", init_llh[0]},   {"init_llh/y", init_llh[1]},   {"init_llh/z", init_llh[2]},
              {"init_ecef/x", init_ecef[0]}, {"init_ecef/y", init_ecef[1]}, {"init_ecef/z", init_ecef[2]}};
  int clock_calls = 0;
  b.now = [&]() { return clock_calls++ == 0 ? clock_arrive : clock_later; };
  b.service = [&](void* p) {
    core_nav::SetStopping* srv = static_cast<core_nav::SetStopping*>(p);
    std::memcpy(srv->response.PvecData.data(), Pvec, 225 * sizeof(double));
    std::memcpy(srv->response.QvecData.data(), Qvec, 225 * sizeof(double));
    std::memcpy(srv->response.STMvecData.data(), STMvec, 225 * sizeof(double));
    std::memcpy(srv->response.HvecData.data(), Hvec, 60 * sizeof(double));
    srv->response.PosData.x = pos[0];
    srv->response.PosData.y = pos[1];
    srv->response.PosData.z = pos[2];
    return true;
  };
  ros::NodeHandle nh("");
  GpPredictor gp(nh);
  gp.new_gp_data_arrived_ = false;
  if (!gp.LoadParameters(nh)) return -1;
  std::shared_ptr<core_nav::GP_Output> msg = std::make_shared<core_nav::GP_Output>();
  msg->mean.assign(mean, mean + M);
  msg->sigma.assign(sigma, sigma + M);
  gp.GPCallBack(msg);
  *n_published = (int)b.published.size();
  *stop_cmd = b.published.empty() ? std::nan("") : b.published.back();
  *xy_err_last = gp.xy_errSlip;
  *n_steps = (int)b.throttle_values.size();
  if (xy_trace)
    for (int k = 0; k < (int)b.throttle_values.size() && k < cap_trace; ++k) xy_trace[k] = b.throttle_values[k];
  if (P_out)
    for (int r = 0; r < 15; ++r)
      for (int c = 0; c < 15; ++c) P_out[r * 15 + c] = gp.P_pred(r, c);
  if (K_out)
    for (int r = 0; r < 15; ++r)
      for (int c = 0; c < 4; ++c) K_out[r * 4 + c] = gp.K_pred(r, c);
  if (R_out)
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) R_out[r * 4 + c] = gp.R_IP(r, c);
  b.service = nullptr;
  b.now = nullptr;
  return 0;
}

// GpPredictor::llh_to_enu (gp_predictor.cpp:144-178)
int ref_llh_to_enu(double lat, double lon, double h, const double* init_llh, const double* init_ecef, double* enu) {
  ros::stub::Bench& b = ros::stub::bench();
  b.params = {{"init_llh/x", init_llh[0]},   {"init_llh/y", init_llh[1]},   {"init_llh/z", init_llh[2]},
              {"init_ecef/x", init_ecef[0]}, {"init_ecef/y", init_ecef[1]}, {"init_ecef/z", init_ecef[2]}};
  ros::NodeHandle nh("");
  GpPredictor gp(nh);
  if (!gp.LoadParameters(nh)) return -1;
  GpPredictor::Vector3 p = gp.llh_to_enu(lat, lon, h);
  enu[0] = p(0); enu[1] = p(1); enu[2] = p(2);
  return 0;
}

}  // extern "C"
