// TEST INFRASTRUCTURE: declarations-only stand-in for <ros/ros.h> (see README.md)
#pragma once
#include <cstdio>
#include <string>
#include <vector>
#include <boost_like_ptr.h>
namespace ros {
struct Time {
    double toSec() const;
    static Time now();
};
struct Duration { Duration(double = 0); };
struct Rate { Rate(double); bool sleep(); };
struct Publisher { template <typename M> void publish(const M &) const; };
struct Subscriber {};
struct NodeHandle {
    template <typename M> Publisher advertise(const std::string &, unsigned, bool = false);
    template <typename M> Subscriber subscribe(const std::string &, unsigned, void (*)(const M &));
    template <typename M> Subscriber subscribe(const std::string &, unsigned, void (*)(const shim::const_ptr<M> &));
};
struct AsyncSpinner { AsyncSpinner(unsigned); void start(); };
void init(int &, char **, const std::string &);
void spinOnce();
void spin();
bool ok();
void waitForShutdown();
}  // namespace ros
#define ROS_INFO(...) printf(__VA_ARGS__)
#define ROS_WARN(...) printf(__VA_ARGS__)
#define ROS_ERROR(...) printf(__VA_ARGS__)
#define ROS_INFO_THROTTLE(period, ...) printf(__VA_ARGS__)
#define ROS_WARN_THROTTLE(period, ...) printf(__VA_ARGS__)
namespace std_msgs { struct Header { unsigned seq; ros::Time stamp; std::string frame_id; }; }
