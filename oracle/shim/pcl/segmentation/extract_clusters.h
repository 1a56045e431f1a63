// Stand-in for <pcl/segmentation/extract_clusters.h> — TEST INFRASTRUCTURE ONLY.
// Used by the reference's side thread only (dsp_dynamic.h:1407-1417); NOT on the hot path.
// Semantics (the documented behaviour of pcl::EuclideanClusterExtraction, made deterministic):
//   * clusters are the connected components of the graph "squared distance <= tolerance^2"
//     (fp32: dx*dx + dy*dy + dz*dz, left to right);
//   * components with size < min or > max are dropped;
//   * indices inside a cluster ascend; clusters are ordered by size descending, ties by smallest
//     member index (PCL sorts with an unstable std::sort; the tie rule here pins it).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <unordered_map>
#include <pcl/point_types.h>
namespace pcl {
namespace search {
template <typename T>
class KdTree {
public:
    typedef std::shared_ptr<KdTree<T>> Ptr;
    void setInputCloud(const typename PointCloud<T>::Ptr &) {}
};
}  // namespace search

template <typename T>
class EuclideanClusterExtraction {
public:
    void setClusterTolerance(double t) { tol_ = (float)t; }
    void setMinClusterSize(int n) { min_ = n; }
    void setMaxClusterSize(int n) { max_ = n; }
    void setSearchMethod(const typename search::KdTree<T>::Ptr &) {}
    void setInputCloud(const typename PointCloud<T>::Ptr &c) { cloud_ = c; }
    void extract(std::vector<PointIndices> &out) {
        out.clear();
        const std::vector<T> &pts = cloud_->points;
        const int n = (int)pts.size();
        if (n == 0 || !(tol_ > 0.f)) return;
        const float tol2 = tol_ * tol_;
        auto cell = [&](float v) { return (int64_t)std::floor(v / tol_); };
        auto key = [](int64_t a, int64_t b, int64_t c) {
            return (uint64_t)((a + (1 << 20)) & 0x1FFFFF) << 42 | (uint64_t)((b + (1 << 20)) & 0x1FFFFF) << 21 |
                   (uint64_t)((c + (1 << 20)) & 0x1FFFFF);
        };
        std::unordered_map<uint64_t, std::vector<int>> grid;
        grid.reserve(n);
        for (int i = 0; i < n; ++i) grid[key(cell(pts[i].x), cell(pts[i].y), cell(pts[i].z))].push_back(i);
        std::vector<char> seen(n, 0);
        std::vector<int> queue;
        for (int s = 0; s < n; ++s) {
            if (seen[s]) continue;
            queue.clear();
            queue.push_back(s);
            seen[s] = 1;
            for (size_t h = 0; h < queue.size(); ++h) {
                const T &p = pts[queue[h]];
                int64_t cx = cell(p.x), cy = cell(p.y), cz = cell(p.z);
                for (int64_t a = cx - 1; a <= cx + 1; ++a)
                    for (int64_t b = cy - 1; b <= cy + 1; ++b)
                        for (int64_t c = cz - 1; c <= cz + 1; ++c) {
                            auto it = grid.find(key(a, b, c));
                            if (it == grid.end()) continue;
                            for (int j : it->second) {
                                if (seen[j]) continue;
                                float dx = pts[j].x - p.x, dy = pts[j].y - p.y, dz = pts[j].z - p.z;
                                float d2 = dx * dx + dy * dy + dz * dz;
                                if (d2 <= tol2) { seen[j] = 1; queue.push_back(j); }
                            }
                        }
            }
            if ((int)queue.size() >= min_ && (int)queue.size() <= max_) {
                PointIndices r;
                r.indices = queue;
                std::sort(r.indices.begin(), r.indices.end());
                out.push_back(r);
            }
        }
        std::stable_sort(out.begin(), out.end(), [](const PointIndices &a, const PointIndices &b) {
            return a.indices.size() > b.indices.size();
        });
    }
private:
    typename PointCloud<T>::Ptr cloud_;
    float tol_ = 0.f;
    int min_ = 1, max_ = 1 << 30;
};
}  // namespace pcl
