// ROS1 wire formats of the messages either side of the hot path, header-only.
//
// The reference moves its data through roscpp/rospy-generated (de)serialisers for
//     core_nav/GP_Input      core_navigation/msg/GP_Input.msg:1-3     (CoreNav.cpp:304 -> gp_slip_node.py:81)
//     core_nav/GP_Output     core_navigation/msg/GP_Output.msg:1-3    (gp_slip_node.py:63 -> gp_predictor.cpp:11)
//     core_nav/SetStopping   core_navigation/srv/SetStopping.srv:1-7  (gp_predictor.cpp:26 <-> CoreNav.cpp:652-676)
//     std_msgs/Float64       stop_cmd                                  (gp_predictor.cpp:118)
// ROS and its message generators are not in this image, so the byte layout is restated here from the ROS1
// serialisation rules: little-endian; uint32 length prefix for strings and variable-length arrays, none for
// fixed-length arrays; Header = uint32 seq, time stamp (uint32 secs, uint32 nsecs), string frame_id; bool = 1 byte.
// A message on a TCPROS connection is framed by a uint32 byte count - frame() / unframe(); a rosbag record's data
// field holds the serialised message without that prefix.  The md5 sums are the ROS "md5 text" sums of the
// definitions above (tests/test_wire.py rebuilds them from the definitions and pins the method on the published sums of std_msgs/Header, geometry_msgs/Point and
// std_msgs/Float64), so a TCPROS connection header written with them is accepted by a live roscore graph.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "gp_predictor_b200.hpp"

namespace cngp_wire {

constexpr const char* MD5_HEADER = "2176decaecbce78abc3b96ef049fabed";
constexpr const char* MD5_POINT = "4a842b65f413084dc2b10fb484ea7f17";
constexpr const char* MD5_FLOAT64 = "fdb28210bfa9d7c91146260178d9a584";
constexpr const char* MD5_GP_INPUT = "9753e28f26b0947dec1baef0e82339bc";
constexpr const char* MD5_GP_OUTPUT = "aa85e91d502deb241dc28762eb372b44";
constexpr const char* MD5_SET_STOPPING = "24fce43738a51f1ac343c3c21c375939";
constexpr const char* TYPE_GP_INPUT = "core_nav/GP_Input";
constexpr const char* TYPE_GP_OUTPUT = "core_nav/GP_Output";
constexpr const char* TYPE_SET_STOPPING = "core_nav/SetStopping";
constexpr size_t SET_STOPPING_RESPONSE_BYTES = (3 * 225 + 60 + 3) * 8;   // 5904

using Bytes = std::vector<uint8_t>;

struct Writer {
  Bytes buf;
  void u8(uint8_t v) { buf.push_back(v); }
  void u32(uint32_t v) { for (int i = 0; i < 4; ++i) buf.push_back((uint8_t)(v >> (8 * i))); }
  void f64(double v) {
    uint64_t b;
    std::memcpy(&b, &v, 8);
    for (int i = 0; i < 8; ++i) buf.push_back((uint8_t)(b >> (8 * i)));
  }
  void f64s(const double* p, size_t n) { for (size_t i = 0; i < n; ++i) f64(p[i]); }
  void str(const std::string& s) { u32((uint32_t)s.size()); buf.insert(buf.end(), s.begin(), s.end()); }
  void vec(const std::vector<double>& v) { u32((uint32_t)v.size()); f64s(v.data(), v.size()); }
};

struct Reader {
  const uint8_t* p;
  size_t n, at = 0;
  Reader(const uint8_t* data, size_t size) : p(data), n(size) {}
  void need(size_t k) const { if (at + k > n) throw std::out_of_range("cngp_wire: truncated message"); }
  uint8_t u8() { need(1); return p[at++]; }
  uint32_t u32() {
    need(4);
    uint32_t v = 0;
    for (int i = 0; i < 4; ++i) v |= (uint32_t)p[at + i] << (8 * i);
    at += 4;
    return v;
  }
  double f64() {
    need(8);
    uint64_t b = 0;
    for (int i = 0; i < 8; ++i) b |= (uint64_t)p[at + i] << (8 * i);
    at += 8;
    double v;
    std::memcpy(&v, &b, 8);
    return v;
  }
  void f64s(double* out, size_t k) { for (size_t i = 0; i < k; ++i) out[i] = f64(); }
  std::string str() {
    const uint32_t k = u32();
    need(k);
    std::string s(reinterpret_cast<const char*>(p + at), k);
    at += k;
    return s;
  }
  std::vector<double> vec() {
    const uint32_t k = u32();
    need((size_t)k * 8);
    std::vector<double> v(k);
    f64s(v.data(), k);
    return v;
  }
  void done() const { if (at != n) throw std::length_error("cngp_wire: trailing bytes"); }
};

// ros::Time(double): sec = floor(t), nsec = round((t - sec) 1e9), carried into sec when it reaches 1e9
inline void stamp_to_ros(double t, uint32_t& sec, uint32_t& nsec) {
  const double fl = std::floor(t);
  long long s = (long long)fl, ns = (long long)std::llround((t - fl) * 1e9);
  s += ns / 1000000000LL;
  ns %= 1000000000LL;
  if (s < 0 || s > 0xffffffffLL) throw std::range_error("cngp_wire: stamp outside the ROS time range");
  sec = (uint32_t)s;
  nsec = (uint32_t)ns;
}
inline double stamp_from_ros(uint32_t sec, uint32_t nsec) { return (double)sec + 1e-9 * (double)nsec; }   // ros::Time::toSec

inline void put(Writer& w, const core_nav::Header& h) {
  uint32_t s, ns;
  if (h.stamp_raw && h.stamp == stamp_from_ros(h.stamp_sec, h.stamp_nsec)) { s = h.stamp_sec; ns = h.stamp_nsec; }   // byte-exact round trip
  else stamp_to_ros(h.stamp, s, ns);
  w.u32(h.seq); w.u32(s); w.u32(ns); w.str(h.frame_id);
}
inline void get(Reader& r, core_nav::Header& h) {
  h.seq = r.u32();
  const uint32_t s = r.u32(), ns = r.u32();
  h.stamp = stamp_from_ros(s, ns);
  h.stamp_sec = s; h.stamp_nsec = ns; h.stamp_raw = true;
  h.frame_id = r.str();
}

inline Bytes serialize(const core_nav::GP_Input& m) {
  Writer w;
  put(w, m.header); w.vec(m.time_array); w.vec(m.slip_array);
  return w.buf;
}
inline Bytes serialize(const core_nav::GP_Output& m) {
  Writer w;
  put(w, m.header); w.vec(m.mean); w.vec(m.sigma);
  return w.buf;
}
inline Bytes serialize(const core_nav::SetStopping::Request& m) { return Bytes{(uint8_t)(m.stopping ? 1 : 0)}; }
inline Bytes serialize(const core_nav::SetStopping::Response& m) {
  Writer w;
  w.f64s(m.PvecData.data(), 225); w.f64s(m.QvecData.data(), 225); w.f64s(m.STMvecData.data(), 225);
  w.f64s(m.HvecData.data(), 60);
  w.f64(m.PosData.x); w.f64(m.PosData.y); w.f64(m.PosData.z);
  return w.buf;
}
inline Bytes serialize(const std_msgs::Float64& m) {
  Writer w;
  w.f64(m.data);
  return w.buf;
}

inline void deserialize(const uint8_t* p, size_t n, core_nav::GP_Input& m) {
  Reader r(p, n);
  get(r, m.header); m.time_array = r.vec(); m.slip_array = r.vec();
  r.done();
}
inline void deserialize(const uint8_t* p, size_t n, core_nav::GP_Output& m) {
  Reader r(p, n);
  get(r, m.header); m.mean = r.vec(); m.sigma = r.vec();
  r.done();
}
inline void deserialize(const uint8_t* p, size_t n, core_nav::SetStopping::Request& m) {
  Reader r(p, n);
  m.stopping = r.u8() != 0;
  r.done();
}
inline void deserialize(const uint8_t* p, size_t n, core_nav::SetStopping::Response& m) {
  Reader r(p, n);
  r.f64s(m.PvecData.data(), 225); r.f64s(m.QvecData.data(), 225); r.f64s(m.STMvecData.data(), 225);
  r.f64s(m.HvecData.data(), 60);
  m.PosData.x = r.f64(); m.PosData.y = r.f64(); m.PosData.z = r.f64();
  r.done();
}
inline void deserialize(const uint8_t* p, size_t n, std_msgs::Float64& m) {
  Reader r(p, n);
  m.data = r.f64();
  r.done();
}
template <class M>
inline void deserialize(const Bytes& b, M& m) { deserialize(b.data(), b.size(), m); }

// TCPROS framing: uint32 byte count, then the serialised message
inline Bytes frame(const Bytes& body) {
  Writer w;
  w.u32((uint32_t)body.size());
  w.buf.insert(w.buf.end(), body.begin(), body.end());
  return w.buf;
}
// returns the number of bytes consumed from p (4 + body), or 0 when the buffer does not yet hold a whole frame
inline size_t unframe(const uint8_t* p, size_t n, Bytes& body) {
  if (n < 4) return 0;
  Reader r(p, n);
  const uint32_t k = r.u32();
  if (n < 4 + (size_t)k) return 0;
  body.assign(p + 4, p + 4 + k);
  return 4 + (size_t)k;
}
// service response on the wire: ok byte, then the framed response (or the framed error string when ok == 0)
inline Bytes frame_service_response(bool ok, const Bytes& body) {
  Bytes out{(uint8_t)(ok ? 1 : 0)};
  const Bytes f = frame(body);
  out.insert(out.end(), f.begin(), f.end());
  return out;
}

}  // namespace cngp_wire
