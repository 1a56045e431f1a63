// Stand-in for <pcl/point_types.h> (+ the PointCloud container) — TEST INFRASTRUCTURE ONLY.
// Surface used by the reference: dsp_dynamic.h:130,134,240,257,385-445,815,1387-1439 and
// map_sim_example.cpp:320-323 (.width, .points.at).
#pragma once
#include <memory>
#include <vector>
namespace pcl {
struct PointXYZ { float x = 0.f, y = 0.f, z = 0.f; };
struct PointXYZINormal {
    float x = 0.f, y = 0.f, z = 0.f;
    float intensity = 0.f;
    float normal_x = 0.f, normal_y = 0.f, normal_z = 0.f;
};
struct PointIndices { std::vector<int> indices; };
template <typename T>
class PointCloud {
public:
    typedef std::shared_ptr<PointCloud<T>> Ptr;
    typedef std::shared_ptr<const PointCloud<T>> ConstPtr;
    std::vector<T> points;
    unsigned width = 0, height = 1;
    void push_back(const T &p) { points.push_back(p); width = (unsigned)points.size(); }
    void clear() { points.clear(); width = 0; }
    bool empty() const { return points.empty(); }
    size_t size() const { return points.size(); }
    T &operator[](size_t i) { return points[i]; }
    const T &operator[](size_t i) const { return points[i]; }
    typename std::vector<T>::iterator begin() { return points.begin(); }
    typename std::vector<T>::iterator end() { return points.end(); }
    typename std::vector<T>::const_iterator begin() const { return points.begin(); }
    typename std::vector<T>::const_iterator end() const { return points.end(); }
};
}  // namespace pcl
