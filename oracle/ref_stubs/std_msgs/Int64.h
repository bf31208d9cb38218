// STAND-IN for the generated <std_msgs/Int64.h> (std_msgs/Int64: one field "data") - oracle/_ref build only.
#pragma once
#include <memory>
namespace std_msgs {
struct Int64 {
  long data = 0;
  typedef std::shared_ptr<Int64> Ptr;
  typedef std::shared_ptr<Int64 const> ConstPtr;
};
}  // namespace std_msgs
