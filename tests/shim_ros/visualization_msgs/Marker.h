#pragma once
#include <geometry_msgs/Pose.h>
namespace std_msgs { struct ColorRGBA { float r, g, b, a; }; }
namespace visualization_msgs {
struct Marker {
    enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6, SPHERE_LIST = 7, POINTS = 8, ADD = 0, MODIFY = 0, DELETE = 2, DELETEALL = 3 };
    std_msgs::Header header;
    std::string ns;
    int id, type, action;
    geometry_msgs::Pose pose;
    geometry_msgs::Vector3 scale;
    std_msgs::ColorRGBA color;
    ros::Duration lifetime;
    std::vector<geometry_msgs::Point> points;
    std::vector<std_msgs::ColorRGBA> colors;
};
}
