#pragma once
#include <ros/ros.h>
namespace create_fundamentals {
struct DiffDrive { struct { double left = 0, right = 0; } request; struct {} response; };
struct SensorPacket { double encoderLeft = 0, encoderRight = 0; typedef boost::shared_ptr<SensorPacket const> ConstPtr; };
struct PlaySong { struct { int number = 0; } request; };
struct StoreSong { struct {} request; };
struct ResetEncoders { struct {} request; };
}
