// TEST INFRASTRUCTURE: declarations-only stand-in (see README.md)
#pragma once
#include <pcl/point_types.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl {
struct PointXYZRGB { float x, y, z; unsigned char r, g, b, a; };
template <typename T> void fromROSMsg(const sensor_msgs::PointCloud2 &, PointCloud<T> &);
template <typename T> void toROSMsg(const PointCloud<T> &, sensor_msgs::PointCloud2 &);
}
