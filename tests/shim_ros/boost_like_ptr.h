// TEST INFRASTRUCTURE: the ConstPtr of a message type (boost::shared_ptr<M const> in ROS 1)
#pragma once
#include <memory>
namespace shim { template <typename M> using const_ptr = std::shared_ptr<const M>; }
