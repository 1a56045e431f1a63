#pragma once
#include <ros/ros.h>
namespace geometry_msgs {
struct Point { double x, y, z; };
struct Vector3 { double x, y, z; };
struct Quaternion { double x, y, z, w; };
struct Pose { Point position; Quaternion orientation; };
struct Twist { Vector3 linear, angular; };
}
