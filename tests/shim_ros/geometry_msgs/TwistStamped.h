#pragma once
#include <geometry_msgs/Pose.h>
namespace geometry_msgs { struct TwistStamped { std_msgs::Header header; Twist twist; }; }
