#pragma once
#include <geometry_msgs/Pose.h>
namespace visualization_msgs {
struct Marker { enum { LINE_LIST = 5, ADD = 0 }; std_msgs::Header header; std::string ns; int id = 0, type = 0, action = 0;
  geometry_msgs::Pose pose; struct { double x = 0, y = 0, z = 0; } scale; struct { float r = 0, g = 0, b = 0, a = 0; } color;
  std::vector<geometry_msgs::Point> points; };
}
