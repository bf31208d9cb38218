// ROS-free GpPredictor over the C ABI (include/gp_predictor_b200.hpp).  Host glue only: see the header.
#include "../../include/gp_predictor_b200.hpp"

#include <algorithm>
#include <cmath>

GpPredictor::GpPredictor(cngp_ctx* ctx, StoppingService stopping_service, StopPublisher stop_cmd_pub, Clock now)
    : ctx_(ctx), clt_setStopping_(std::move(stopping_service)), stop_cmd_pub_(std::move(stop_cmd_pub)), now_(std::move(now)) {
  cngp_default_stop_config(&stop_config);
  init_x = stop_config.init_llh[0]; init_y = stop_config.init_llh[1]; init_z = stop_config.init_llh[2];
  init_ecef_x = stop_config.init_ecef[0]; init_ecef_y = stop_config.init_ecef[1]; init_ecef_z = stop_config.init_ecef[2];
}

void GpPredictor::sync_init() {
  stop_config.init_llh[0] = init_x; stop_config.init_llh[1] = init_y; stop_config.init_llh[2] = init_z;
  stop_config.init_ecef[0] = init_ecef_x; stop_config.init_ecef[1] = init_ecef_y; stop_config.init_ecef[2] = init_ecef_z;
}

bool GpPredictor::LoadParameters(const std::map<std::string, double>& params) {
  const char* names[6] = {"init_llh/x", "init_llh/y", "init_llh/z", "init_ecef/x", "init_ecef/y", "init_ecef/z"};
  double* dst[6] = {&init_x, &init_y, &init_z, &init_ecef_x, &init_ecef_y, &init_ecef_z};
  for (int k = 0; k < 6; ++k) {
    auto it = params.find(names[k]);
    if (it == params.end()) return false;
    *dst[k] = it->second;
  }
  return true;
}

GpPredictor::Vector3 GpPredictor::llh_to_enu(const double latitude, const double longitude, const double height) {
  sync_init();
  const double llh[3] = {latitude, longitude, height};
  Vector3 enu{};
  cngp_llh_to_enu(ctx_, llh, 1, &stop_config, enu.data(), CNGP_MEM_HOST);
  return enu;
}

int GpPredictor::GPCallBackBatch(const double* mean, const double* sigma, int64_t B, int32_t M, int32_t* triggered,
                                 int32_t* i_stop, double* xy_err) {
  sync_init();
  return cngp_zupt_lookahead_batch(ctx_, mean, sigma, B, M, P_pred.data(), Q_.data(), STM_.data(), H_.data(), savePos.data(),
                                   /*per_window=*/0, &stop_config, triggered, i_stop, nullptr, xy_err, CNGP_MEM_HOST);
}

bool GpPredictor::GPCallBack(const core_nav::GP_Output& gp_data_in_) {
  gp_data_.mean = gp_data_in_.mean;
  gp_data_.sigma = gp_data_in_.sigma;
  gp_arrived_time_ = now_();
  core_nav::SetStopping srv_set_stopping;
  srv_set_stopping.request.stopping = true;
  if (clt_setStopping_(srv_set_stopping)) {
    P_pred = srv_set_stopping.response.PvecData;      // row-major [row*15+col], gp_predictor.cpp:30-36
    Q_ = srv_set_stopping.response.QvecData;
    STM_ = srv_set_stopping.response.STMvecData;
    H_ = srv_set_stopping.response.HvecData;
    savePos = {srv_set_stopping.response.PosData.x, srv_set_stopping.response.PosData.y, srv_set_stopping.response.PosData.z};
    new_gp_data_arrived_ = true;
  }
  // the reference logs a failed service call and only runs the look-ahead when the flag is set (gp_predictor.cpp:53-58)
  bool published = false;
  if (new_gp_data_arrived_) {
    const int32_t M = (int32_t)std::min(gp_data_.mean.size(), gp_data_.sigma.size());
    int32_t trig = 0, i_upd = 0, step = 0;
    double xy = 0.0;
    sync_init();
    std::array<double, 225> P_final{};
    const int rc = M > 0 ? cngp_zupt_lookahead_batch_ex(ctx_, gp_data_.mean.data(), gp_data_.sigma.data(), 1, M,
                                                        P_pred.data(), Q_.data(), STM_.data(), H_.data(), savePos.data(), 0,
                                                        &stop_config, &trig, &i_upd, &step, &xy, P_final.data(),
                                                        K_pred.data(), R_IP.data(), CNGP_MEM_HOST)
                         : CNGP_OK;
    if (rc == CNGP_OK && M > 0) {
      xy_errSlip = xy;
      i = i_upd;
      slip_i = step;
      P_pred = P_final;                 // the reference propagates P_pred in place (gp_predictor.cpp:66,91)
      // R_IP_1 / R_IP_2 of the last update (gp_predictor.cpp:69-87); ENU of the saved position and of its +-3 sigma
      // corners at the last step (gp_predictor.cpp:95-97)
      const double it = 1.0 / stop_config.track;
      R_IP_1 = {0.5, 0.5, 0.0, 0.0, it, -it, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0};
      if (i_upd > 0) {
        const double mu = gp_data_.mean[i_upd - 1], sg = gp_data_.sigma[i_upd - 1];
        const double c0 = stop_config.v_nom / (1.0 - mu), c1 = stop_config.v_nom / (1.0 - (mu + sg)),
                     c2 = stop_config.v_nom / (1.0 - (mu - sg));
        const double est = (c0 + c1 + c2) / 3.0;
        const double cov = ((c0 - est) * (c0 - est) + (c1 - est) * (c1 - est) + (c2 - est) * (c2 - est)) / 3.0;
        const double fa = stop_config.floor_a * stop_config.floor_a, fb = stop_config.floor_b * stop_config.floor_b;
        R_IP_2.fill(0.0);
        R_IP_2[0] = std::max(fa, cov * cov); R_IP_2[5] = std::max(fa, cov * cov); R_IP_2[10] = std::max(fb, cov * cov);
        R_IP_2[15] = fb;
      }
      const double s6 = 3.0 * std::sqrt(std::fabs(P_pred[6 * 15 + 6])), s7 = 3.0 * std::sqrt(std::fabs(P_pred[7 * 15 + 7])),
                   s8 = 3.0 * std::sqrt(std::fabs(P_pred[8 * 15 + 8]));
      const double llh3[9] = {savePos[0], savePos[1], savePos[2], savePos[0] - s6, savePos[1] - s7, savePos[2] - s8,
                              savePos[0] + s6, savePos[1] + s7, savePos[2] + s8};
      double enu3[9];
      if (cngp_llh_to_enu(ctx_, llh3, 3, &stop_config, enu3, CNGP_MEM_HOST) == CNGP_OK) {
        ins_enu_slip = {enu3[0], enu3[1], enu3[2]};
        ins_enu_slip_3p = {enu3[3], enu3[4], enu3[5]};
        ins_enu_slip3p = {enu3[6], enu3[7], enu3[8]};
      }
      if (trig) {
        // gp_predictor.cpp:107-118: i/10.0 seconds of odometry updates after the GP result arrived; late => 0.5 s
        const double dt = gp_arrived_time_ + i / 10.0 - now_();
        stop_cmd_msg_.data = dt < 0.0 ? 0.5 : dt;
        stop_cmd_pub_(stop_cmd_msg_);
        published = true;
      }
    }
    new_gp_data_arrived_ = false;   // gp_predictor.cpp:126-128 (i and slip_i keep the values of this call for inspection)
  }
  return published;
}
