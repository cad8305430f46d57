#pragma once
#include <ros/ros.h>
namespace std_msgs { struct Header { ros::Time stamp; std::string frame_id; }; struct Empty { typedef boost::shared_ptr<Empty const> ConstPtr; }; }
