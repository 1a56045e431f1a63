#pragma once
#include <geometry_msgs/Pose.h>
namespace gazebo_msgs { struct ModelStates { std::vector<std::string> name; std::vector<geometry_msgs::Pose> pose; std::vector<geometry_msgs::Twist> twist; typedef shim::const_ptr<ModelStates> ConstPtr; }; }
