// STAND-IN for the generated <std_msgs/Float64.h> (std_msgs/Float64: one field "data") - oracle/_ref build only.
#pragma once
#include <memory>
namespace std_msgs {
struct Float64 {
  double data = 0;
  typedef std::shared_ptr<Float64> Ptr;
  typedef std::shared_ptr<Float64 const> ConstPtr;
};
}  // namespace std_msgs
