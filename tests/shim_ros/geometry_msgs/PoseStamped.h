#pragma once
#include <geometry_msgs/Pose.h>
namespace geometry_msgs { struct PoseStamped { std_msgs::Header header; Pose pose; typedef shim::const_ptr<PoseStamped> ConstPtr; }; }
