#pragma once
#include <pcl/point_types.h>
namespace pcl {
template <typename T> struct VoxelGrid {
    void setInputCloud(const typename PointCloud<T>::Ptr &);
    void setLeafSize(float, float, float);
    void filter(PointCloud<T> &);
};
}
