// ROS-free C++ drop-in of the reference's stop-predictor node class (row a13 of SURVEY.md section 8).
//
// Mirrors gp_predictor/include/gp_predictor/gp_predictor.h:18-62 - same class name, same public method names
// (GPCallBack, LoadParameters, llh_to_enu) and the same public data members the callback fills (gp_data_, H_, P_pred,
// STM_, Q_, savePos, xy_errSlip, i, slip_i, gp_arrived_time_, new_gp_data_arrived_, stop_cmd_msg_) - with the three ROS
// attachments replaced by plain callables:
//     ros::ServiceClient clt_setStopping_  (gp_predictor.cpp:12,26)  -> StoppingService
//     ros::Publisher     stop_cmd_pub_     (gp_predictor.cpp:13,118) -> StopPublisher
//     ros::Time::now().toSec()             (gp_predictor.cpp:22,107) -> Clock
// and the message types by PODs of the same names and fields (core_navigation/msg/GP_Input.msg, GP_Output.msg,
// srv/SetStopping.srv).  All arithmetic of the callback (gp_predictor.cpp:64-124) runs in the CUDA look-ahead kernel
// behind cngp_zupt_lookahead_batch; this class only unpacks the service response the way the reference does
// (row-major 15x15, the H aliasing index of gp_predictor.cpp:38-42 is applied inside the kernel) and turns the kernel's
// (triggered, i) into the stop command (gp_predictor.cpp:102-122).
#ifndef GP_PREDICTOR_B200_HPP_
#define GP_PREDICTOR_B200_HPP_

#include <algorithm>
#include <array>
#include <cstdint>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "cngp.h"

namespace core_nav {
struct Header {
  uint32_t seq = 0;
  double stamp = 0.0;          // seconds (ros::Time::toSec); the convenient view
  std::string frame_id;
  // the integer pair the message arrived with (cngp_wire.hpp).  A double at epoch scale resolves ~240 ns, so the pair
  // cannot be recovered from `stamp`; the serialiser re-uses it while `stamp` still equals the double it decoded to.
  uint32_t stamp_sec = 0, stamp_nsec = 0;
  bool stamp_raw = false;
};
struct GP_Input {   // core_navigation/msg/GP_Input.msg:1-3
  Header header;
  std::vector<double> time_array, slip_array;
};
struct GP_Output {  // core_navigation/msg/GP_Output.msg:1-3
  Header header;
  std::vector<double> mean, sigma;
};
struct Point {
  double x = 0, y = 0, z = 0;
};
struct SetStopping {  // core_navigation/srv/SetStopping.srv:1-7
  struct Request { bool stopping = false; } request;
  struct Response {
    std::array<double, 225> PvecData{}, QvecData{}, STMvecData{};
    std::array<double, 60> HvecData{};
    Point PosData;
  } response;
};
}  // namespace core_nav

namespace std_msgs {
struct Float64 { double data = 0.0; };
}

class GpPredictor {
 public:
  typedef std::array<double, 3> Vector3;
  using StoppingService = std::function<bool(core_nav::SetStopping&)>;
  using StopPublisher = std::function<void(const std_msgs::Float64&)>;
  using Clock = std::function<double()>;

  // ctx: a cngp context owned by the caller (one per GPU).
  GpPredictor(cngp_ctx* ctx, StoppingService stopping_service, StopPublisher stop_cmd_pub, Clock now);

  // gp_predictor.cpp:17-132.  Returns true when a stop command was published.
  bool GPCallBack(const core_nav::GP_Output& gp_data_in_);
  // Many predictions against the same service response (Monte-Carlo, configs[3]): mean/sigma [B][M] row-major.
  // triggered/i_stop [B]; nothing is published.
  int GPCallBackBatch(const double* mean, const double* sigma, int64_t B, int32_t M, int32_t* triggered, int32_t* i_stop,
                      double* xy_err);
  // gp_predictor.cpp:134-142: init_llh/{x,y,z}, init_ecef/{x,y,z}.  The reference never calls it;
  // without it the defaults of core_navigation/config/init_params.yaml:9-16 apply.
  bool LoadParameters(const std::map<std::string, double>& params);
  // gp_predictor.cpp:144-178 (evaluated on the device)
  Vector3 llh_to_enu(const double latitude, const double longitude, const double height);

  core_nav::GP_Output gp_data_;
  // gp_predictor.h:36-43.  Row-major arrays stand in for the Eigen matrices.  After GPCallBack they hold what the
  // reference's members hold: P_pred the covariance after the last look-ahead step executed, K_pred / R_IP / R_IP_2 the
  // gain and the odometry noise of the last update, R_IP_1 the constant wheel-geometry matrix (gp_predictor.cpp:84-87).
  std::array<double, 16> R_IP{}, R_IP_1{}, R_IP_2{};
  std::array<double, 60> K_pred{};    // 15 x 4
  std::array<double, 60> H_{};        // as received: HvecData (the kernel applies the reference's index)
  std::array<double, 225> P_pred{}, STM_{}, Q_{};
  Vector3 savePos{}, ins_enu_slip{}, ins_enu_slip3p{}, ins_enu_slip_3p{};   // gp_predictor.cpp:95-97
  std_msgs::Float64 stop_cmd_msg_;
  bool new_gp_data_arrived_ = false;
  double gp_arrived_time_ = 0.0;
  double xy_errSlip = 0.0;
  double init_ecef_x, init_ecef_y, init_ecef_z, init_x, init_y, init_z;
  int slip_i = 0;
  int i = 0;
  cngp_stop_config stop_config;       // every hard-coded constant of gp_predictor.cpp:73-88,102 (defaults = reference)

 private:
  cngp_ctx* ctx_;
  StoppingService clt_setStopping_;
  StopPublisher stop_cmd_pub_;
  Clock now_;
  void sync_init();
};

// The GP node's callback for a C++ caller (core_navigation/script/gp_slip_node.py:16-63, rows a1-a7): train on the first
// int(0.9 n) samples, fit the hyper-parameters from all-ones (theta == nullptr, as m.optimize() does) or take them as
// given (noise last), predict on arange(min(time), max(time) + horizon, 1) and return mean[n:], sigma = 2 sqrt(var[n:]).
// Throws std::runtime_error with the library's message when the window cannot be processed (GPy would raise).
// Implemented in corenav_gp_b200/host/gp_slip_predict.cpp (libgp_predictor_b200.so).
core_nav::GP_Output gp_slip_predict(cngp_ctx* ctx, const core_nav::GP_Input& data, const char* kernel = "rbf*brownian",
                                    const double* theta = nullptr, int horizon = 600);

#endif  // GP_PREDICTOR_B200_HPP_
