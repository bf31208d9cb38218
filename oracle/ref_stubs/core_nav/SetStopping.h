// STAND-IN for the header catkin generates from core_navigation/srv/SetStopping.srv:1-7 - oracle/_ref build only.
#pragma once
#include <array>
#include <geometry_msgs/Point.h>
namespace core_nav {
struct SetStoppingRequest { bool stopping = false; };
struct SetStoppingResponse {
  std::array<double, 225> PvecData{};
  std::array<double, 225> QvecData{};
  std::array<double, 225> STMvecData{};
  std::array<double, 60> HvecData{};
  geometry_msgs::Point PosData;
};
struct SetStopping {
  typedef SetStoppingRequest Request;
  typedef SetStoppingResponse Response;
  Request request;
  Response response;
};
}  // namespace core_nav
