#pragma once
#include <ros/ros.h>
namespace pink_fundamentals {
struct PID_drive { struct { double x = 0, y = 0, degree = 0, speed = 0; } request; };
struct Wanderer { struct { bool isWander = false; } request; };
struct align_call { struct {} request; };
struct Pose { int row = 0, column = 0, orientation = 0; };
struct ExactPose { float x = 0, y = 0, thetaQuaternion = 0, theta = 0; int orientation = 0; };
}
