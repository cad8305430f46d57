#pragma once
#include <std_msgs/Empty.h>
#include <vector>
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 0; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
struct PoseArray { std_msgs::Header header; std::vector<Pose> poses; };
}
