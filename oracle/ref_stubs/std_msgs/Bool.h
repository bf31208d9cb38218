// STAND-IN for the generated <std_msgs/Bool.h> (std_msgs/Bool: one field "data") - oracle/_ref build only.
#pragma once
#include <memory>
namespace std_msgs {
struct Bool {
  bool data = 0;
  typedef std::shared_ptr<Bool> Ptr;
  typedef std::shared_ptr<Bool const> ConstPtr;
};
}  // namespace std_msgs
