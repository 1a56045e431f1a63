// velocity_estimator.h — host-side stand-in for the reference's side thread (product code).
// Replaces DSPMap::velocityEstimationThread (g-ch/DSP-map include/dsp_dynamic.h:1377-1544; static variant
// dsp_static.h:1285-1309) without PCL / munkres-cpp: ground split, Euclidean clustering on a hash grid,
// cluster centroids, static / dynamic classification, Hungarian matching to the previous frame's clusters,
// per-point velocity tagging.  Its output is the newborn stage's input cloud (7 floats per point, world frame).
#pragma once
#include <cstdint>
#include <vector>
#include "dspmap_types.h"

struct ClusterFeature {  // dsp_dynamic.h:98-109
    float cx = 0.f, cy = 0.f, cz = 0.f;
    int point_num = 0;
    float vx = -10000.f, vy = -10000.f, vz = -10000.f;
    float v = 0.f;
    float intensity = 0.f;
};

struct VelocityEstimator {
    float filter_res = 0.15f;  // voxel_filtered_resolution (dsp_dynamic.h:132)
    u64 seed = 0, draws = 0;   // helper uniform stream (cluster colours, :1422)
    std::vector<ClusterFeature> last;
    std::vector<float> rotated;  // scratch: in-FOV rotated points (cloud_in_current_view_rotated, :130)
    std::vector<float> statics, nonground;  // scratch: world-frame xyz triples of the ground / non-ground split

    void reset(u64 s) { seed = s ^ 0xA5A5A5A5DEADBEEFull; draws = 0; last.clear(); }
    float uniform(float lo, float hi);
    // pts: n x 3 points in the sensor frame. tagged_out is left untouched when no point is in view (:1379).
    void match(std::vector<ClusterFeature> &cur, float dt);
    void finish_device(const EstFeature *feat, int n_dynamic, int n_clusters, float dt, float *cvel);
    void estimate(const MapConst &mc, const FrameConst &fc, const float *planes0, const float *pts, int n, int model,
                  std::vector<float> &tagged_out);
};

// Euclidean clustering: connected components of "squared distance <= tol^2", sizes in [min_size, max_size], member
// indices ascending, clusters by size descending (ties: smallest member index first).
void euclidean_clusters(const float *xyz, int n, float tol, int min_size, int max_size, std::vector<std::vector<int>> &out);
// path: 0 = automatic (dense grid when the bounding box allows, else hash grid), 1 = hash grid, 2 = dense grid; false when
// the requested path cannot take the cloud.  Both paths return the same clusters.
bool euclidean_clusters_path(const float *xyz, int n, float tol, int min_size, int max_size, int path, std::vector<std::vector<int>> &out);
// Minimum-cost assignment of an R x C cost matrix padded to square with its maximum; assign[r] = column or -1.
void hungarian(const std::vector<float> &cost, int R, int C, std::vector<int> &assign);
