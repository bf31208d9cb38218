// STAND-IN for the generated <geometry_msgs/Point.h> - oracle/_ref build only.
#pragma once
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
}  // namespace geometry_msgs
