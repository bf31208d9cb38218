// STAND-IN for the header catkin generates from core_navigation/msg/GP_Output.msg:1-3 - oracle/_ref build only.
#pragma once
#include <memory>
#include <vector>
#include <std_msgs/Header.h>
namespace core_nav {
struct GP_Output {
  std_msgs::Header header;
  std::vector<double> mean;
  std::vector<double> sigma;
  typedef std::shared_ptr<GP_Output> Ptr;
  typedef std::shared_ptr<GP_Output const> ConstPtr;
};
}  // namespace core_nav
