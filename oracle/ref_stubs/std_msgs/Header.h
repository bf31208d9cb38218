// STAND-IN for the generated <std_msgs/Header.h> - oracle/_ref build only.
#pragma once
#include <cstdint>
#include <string>
namespace std_msgs {
struct Header {
  uint32_t seq = 0;
  struct { uint32_t sec = 0, nsec = 0; } stamp;
  std::string frame_id;
};
}  // namespace std_msgs
