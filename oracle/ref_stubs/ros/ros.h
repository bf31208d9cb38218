// STAND-IN for <ros/ros.h> - TEST INFRASTRUCTURE ONLY (oracle/_ref build, see oracle/ref_build.py).
//
// ROS is not installed in this image.  This header gives the UNMODIFIED reference translation unit
// /root/reference/gp_predictor/src/gp_predictor.cpp the handful of roscpp names it uses (NodeHandle::subscribe /
// advertise / serviceClient, ServiceClient::call, Publisher::publish, ros::Time::now, ros::param::get, ros::init,
// ros::spin, the ROS_* logging macros) wired to an in-process test bench instead of a ROS graph:
//   * ServiceClient::call(srv)  -> bench().service(&srv)      (the harness fills the SetStopping response)
//   * Publisher::publish(msg)   -> bench().published.push_back(msg.data)
//   * ros::Time::now()          -> bench().now()              (an injected clock)
//   * ROS_ERROR_THROTTLE(p, fmt, v...) -> bench().throttle_values gets every numeric argument at full precision
//     (gp_predictor.cpp:100 logs xy_errSlip once per look-ahead step: that is the per-step trace the tests compare).
// Nothing here comes from roscpp's sources.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

namespace ros {

namespace stub {
struct Bench {
  std::function<bool(void*)> service;            // receives a pointer to the service object passed to call()
  std::function<double()> now;                   // wall clock in seconds
  std::vector<double> published;                 // .data of every message published
  std::vector<double> throttle_values;           // numeric arguments of every ROS_*_THROTTLE call
  std::map<std::string, double> params;          // ros::param::get
};
inline Bench& bench() {
  static Bench b;
  return b;
}
template <typename... A>
inline void ignore(const A&...) {}
inline void record() {}
template <typename V, typename... A>
inline void record(const V& v, const A&... rest) {
  if constexpr (std::is_arithmetic<V>::value) bench().throttle_values.push_back((double)v);
  record(rest...);
}
}  // namespace stub

struct Time {
  double t;
  static Time now() { return Time{stub::bench().now ? stub::bench().now() : 0.0}; }
  double toSec() const { return t; }
};

class Subscriber {};

class Publisher {
 public:
  template <typename M>
  void publish(const M& m) const { stub::bench().published.push_back((double)m.data); }
};

class ServiceClient {
 public:
  template <typename S>
  bool call(S& srv) { return stub::bench().service ? stub::bench().service(static_cast<void*>(&srv)) : false; }
};

class NodeHandle {
 public:
  NodeHandle() {}
  NodeHandle(const std::string&) {}
  template <typename M, typename T>
  Subscriber subscribe(const std::string&, unsigned, void (T::*)(const std::shared_ptr<M const>&), T*) { return Subscriber(); }
  template <typename M>
  Publisher advertise(const std::string&, unsigned) { return Publisher(); }
  template <typename S>
  ServiceClient serviceClient(const std::string&) { return ServiceClient(); }
};

namespace param {
inline bool get(const std::string& key, double& v) {
  auto it = stub::bench().params.find(key);
  if (it == stub::bench().params.end()) return false;
  v = it->second;
  return true;
}
}  // namespace param

inline void init(int&, char**, const std::string&) {}
inline void spin() {}

}  // namespace ros

#define ROS_STUB_LOG(...) do { if (false) ::ros::stub::ignore(__VA_ARGS__); } while (0)
#define ROS_DEBUG(...) ROS_STUB_LOG(__VA_ARGS__)
#define ROS_INFO(...) ROS_STUB_LOG(__VA_ARGS__)
#define ROS_WARN(...) ROS_STUB_LOG(__VA_ARGS__)
#define ROS_ERROR(...) ROS_STUB_LOG(__VA_ARGS__)
#define ROS_FATAL(...) ROS_STUB_LOG(__VA_ARGS__)
#define ROS_INFO_STREAM(x) do { } while (0)
#define ROS_ERROR_THROTTLE(period, ...) ::ros::stub::record(__VA_ARGS__)
#define ROS_INFO_THROTTLE(period, ...) ::ros::stub::record(__VA_ARGS__)
#define ROS_WARN_THROTTLE(period, ...) ::ros::stub::record(__VA_ARGS__)
