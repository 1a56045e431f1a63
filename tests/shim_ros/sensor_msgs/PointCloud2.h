#pragma once
#include <ros/ros.h>
namespace sensor_msgs { struct PointCloud2 { std_msgs::Header header; unsigned height, width; std::vector<unsigned char> data; typedef shim::const_ptr<PointCloud2> ConstPtr; }; typedef PointCloud2::ConstPtr PointCloud2ConstPtr; }
