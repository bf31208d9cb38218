// STAND-IN for the header catkin generates from core_navigation/msg/GP_Input.msg:1-3 - oracle/_ref build only.
#pragma once
#include <memory>
#include <vector>
#include <std_msgs/Header.h>
namespace core_nav {
struct GP_Input {
  std_msgs::Header header;
  std::vector<double> time_array;
  std::vector<double> slip_array;
  typedef std::shared_ptr<GP_Input> Ptr;
  typedef std::shared_ptr<GP_Input const> ConstPtr;
};
}  // namespace core_nav
