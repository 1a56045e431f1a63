// oracle/ref_driver.cpp — TEST INFRASTRUCTURE ONLY (never linked into, or called by, the product path).
//
// C driver around the UNMODIFIED reference header (g-ch/DSP-map include/dsp_dynamic.h,
// dsp_dynamic_multiple_neighbors.h or dsp_static.h). build_ref.py substitutes the BASELINE config's
// #define / const-int lines the same way script/set_map_parameters.py:392-452 does, writes the result
// to oracle/_ref/gen/ (git-ignored) and compiles this file against it with the stand-in dependency
// headers in oracle/shim/.  The output is oracle/_ref/libdspref_<cfg>.so, loaded by tests/ and by
// bench.py's reference arm through ctypes.
//
// Determinism hooks (no edits to the reference): the header seeds from the wall clock
// (srand(time(0)) dsp_dynamic.h:586, default_random_engine(time(NULL)) :1151) and calls libc rand()
// (:1552).  This file defines time()/srand()/rand() itself (linked with -Bsymbolic so the header's
// calls bind here):
//   * time()  returns the seed given to ref_create();
//   * rand()  is the counter-based stream  u31(seed, k) = splitmix64(seed + (k+1)*GOLDEN) >> 33,
//             one counter for the calling (main) thread = the newborn stream, and a separate one for
//             the reference's helper thread (cluster colours, dsp_dynamic.h:1422).
// The GPU library uses the same counter-based stream (include/dspmap_b200.h, dspmap_config.uniform_seed).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <memory>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "Eigen/Eigen"
#include <pcl/point_types.h>
#include <pcl/common/transforms.h>
#include <pcl_conversions/pcl_conversions.h>
#include <pcl/segmentation/extract_clusters.h>
#include "munkres.h"

// ---- determinism hooks -------------------------------------------------------------------------
static uint64_t g_seed = 1;
static std::thread::id g_main_thread;
static uint64_t g_rand_main = 0;    // calls made from the thread that called ref_create()
static uint64_t g_rand_helper = 0;  // calls made from any other thread (velocityEstimationThread)

static inline uint32_t u31(uint64_t seed, uint64_t k) {
    uint64_t z = seed + (k + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 33);
}
extern "C" {
time_t time(time_t *t) noexcept {
    if (t) *t = (time_t)g_seed;
    return (time_t)g_seed;
}
void srand(unsigned) noexcept {}
int rand() noexcept {
    if (std::this_thread::get_id() == g_main_thread) return (int)u31(g_seed, g_rand_main++);
    return (int)u31(g_seed ^ 0xA5A5A5A5DEADBEEFull, g_rand_helper++);
}
}

// ---- the reference, verbatim --------------------------------------------------------------------
#define private public  // expose DSPMap's private state for dumping; std headers are already included
#include REF_HEADER
#undef private

static DSPMap *g_map = nullptr;

extern "C" {

// dims[0..15]: V, S, P, L, T, Nh, Nv, neighbor slots, MAX ppv, nx, ny, nz, obs max/pyramid, model(0 dyn,1 static)
void ref_dims(int *d) {
    d[0] = VOXEL_NUM;
    d[1] = SAFE_PARTICLE_NUM_VOXEL;
    d[2] = observation_pyramid_num;
    d[3] = SAFE_PARTICLE_NUM_PYRAMID;
    d[4] = PREDICTION_TIMES;
    d[5] = observation_pyramid_num_h;
    d[6] = observation_pyramid_num_v;
    d[7] = (int)(sizeof(observation_pyramid_neighbors[0]) / sizeof(int));
    d[8] = MAX_PARTICLE_NUM_VOXEL;
    d[9] = MAP_LENGTH_VOXEL_NUM;
    d[10] = MAP_WIDTH_VOXEL_NUM;
    d[11] = MAP_HEIGHT_VOXEL_NUM;
    d[12] = observation_max_points_num_one_pyramid;
#ifdef REF_STATIC
    d[13] = 1;
#else
    d[13] = 0;
#endif
    d[14] = half_fov_h;
    d[15] = half_fov_v;
}
float ref_resolution() { return (float)VOXEL_RESOLUTION; }
void ref_future_times(float *out) {
    for (int i = 0; i < PREDICTION_TIMES; ++i) out[i] = prediction_future_time[i];
}

// Constructs the map exactly like the example app does: global-style default ctor, then the setters of
// src/map_sim_example.cpp:522-528 with caller-provided values.
int ref_create(uint64_t seed, int init_particle_num, float init_weight) {
    if (g_map) return 0;  // function-static state inside update() makes a second map meaningless
    g_seed = seed;
    g_main_thread = std::this_thread::get_id();
    std::cout.setstate(std::ios::failbit);  // silence the per-frame "Velocity estimation done"
    g_map = new DSPMap(init_particle_num, init_weight);
    return 1;
}
void ref_set_prediction_variance(float p, float v) { g_map->setPredictionVariance(p, v); }
void ref_set_observation_stddev(float s) { g_map->setObservationStdDev(s); }
void ref_set_newborn_weight(float w) { g_map->setNewBornParticleWeight(w); }
void ref_set_newborn_number(int n) { g_map->setNewBornParticleNumberofEachPoint(n); }
void ref_set_voxel_filter_resolution(float r) { DSPMap::setOriginalVoxelFilterResolution(r); }
// the particle CSV of update() (dsp_dynamic.h:325-350): the header's own writer, into `folder` (the header's global string)
void ref_set_particle_record_flag(int flag, float record_time, const char *folder) {
    particle_save_folder = folder;
    g_map->setParticleRecordFlag(flag, record_time);
}

int ref_update(int n, int stride, float *pts, float px, float py, float pz, double t, float qw, float qx,
               float qy, float qz) {
    return g_map->update(n, stride, pts, px, py, pz, t, qw, qx, qy, qz);
}

// update() + getOccupancyMapWithFutureStatus(), returning seconds spent inside the two calls
// (the region BASELINE.md §3 times).
double ref_timed_frame(int n, int stride, float *pts, float px, float py, float pz, double t, float qw,
                       float qx, float qy, float qz, float threshold, float *future, int *n_occ) {
    static pcl::PointCloud<pcl::PointXYZ> cloud;
    cloud.clear();
    auto t0 = std::chrono::steady_clock::now();
    int ok = g_map->update(n, stride, pts, px, py, pz, t, qw, qx, qy, qz);
    int occ = 0;
    if (ok) g_map->getOccupancyMapWithFutureStatus(occ, cloud, future, threshold);
    auto t1 = std::chrono::steady_clock::now();
    *n_occ = ok ? occ : -1;
    return std::chrono::duration<double>(t1 - t0).count();
}

// threshold reader; xyz_out may be null (count only); future may be null (getOccupancyMap variant)
int ref_get_occupancy(float threshold, float *xyz_out, int cap, float *future) {
    pcl::PointCloud<pcl::PointXYZ> cloud;
    int n = 0;
    if (future) g_map->getOccupancyMapWithFutureStatus(n, cloud, future, threshold);
    else g_map->getOccupancyMap(n, cloud, threshold);
    if (xyz_out)
        for (int i = 0; i < n && i < cap; ++i) {
            xyz_out[3 * i] = cloud[i].x;
            xyz_out[3 * i + 1] = cloud[i].y;
            xyz_out[3 * i + 2] = cloud[i].z;
        }
    return n;
}
void ref_clear_prediction() { g_map->clearOccupancyMapPrediction(); }

// input_cloud_with_velocity (world frame): x,y,z,vx,vy,vz,intensity per point
int ref_tagged_cloud(float *out, int cap) {
    int n = (int)input_cloud_with_velocity->size();
    if (out)
        for (int i = 0; i < n && i < cap; ++i) {
            const auto &p = (*input_cloud_with_velocity)[i];
            float *o = out + 7 * i;
            o[0] = p.x; o[1] = p.y; o[2] = p.z;
            o[3] = p.normal_x; o[4] = p.normal_y; o[5] = p.normal_z;
            o[6] = p.intensity;
        }
    return n;
}

// live particles in sweep order: ids[n][2] = {voxel, slot}; vals[n][8] = flag,vx,vy,vz,px,py,pz,w
int ref_dump_particles(int *ids, float *vals, int cap) {
    int n = 0;
    for (int v = 0; v < VOXEL_NUM; ++v)
        for (int s = 0; s < SAFE_PARTICLE_NUM_VOXEL; ++s)
            if (voxels_with_particle[v][s][0] > 0.1f) {
                if (ids && n < cap) {
                    ids[2 * n] = v;
                    ids[2 * n + 1] = s;
                    std::memcpy(vals + 8 * n, &voxels_with_particle[v][s][0], 8 * sizeof(float));
                }
                ++n;
            }
    return n;
}
// the raw 9 floats of one slot, live or not (a vanished particle keeps its data, only the flag is cleared): for debugging
void ref_raw_slot(int v, int s, float *out9) { std::memcpy(out9, &voxels_with_particle[v][s][0], 9 * sizeof(float)); }
void ref_dump_voxel_objects(float *out) {
    std::memcpy(out, &voxels_objects_number[0][0], sizeof(float) * (size_t)VOXEL_NUM * voxels_objects_number_dimension);
}
// binned observations of the last update: counts[P], maxlen[P], pts[P][obs_max][5]
void ref_dump_observations(int *counts, float *maxlen, float *pts) {
    std::memcpy(counts, g_map->observation_num_each_pyramid, sizeof(int) * observation_pyramid_num);
    std::memcpy(maxlen, g_map->point_cloud_max_length, sizeof(float) * observation_pyramid_num);
    if (pts) std::memcpy(pts, g_map->point_cloud, sizeof(float) * (size_t)observation_pyramid_num * observation_max_points_num_one_pyramid * 5);
}
// pyramid lists of the last update, compact: offsets[P+1], entries[n][2] = {voxel, slot} in list order
int ref_dump_pyramid_lists(int *offsets, int *entries, int cap) {
    int n = 0;
    for (int p = 0; p < observation_pyramid_num; ++p) {
        offsets[p] = n;
        for (int j = 0; j < SAFE_PARTICLE_NUM_PYRAMID; ++j)
            if (pyramids_in_fov[p][j][0] & O_MAKE_VALID) {
                if (entries && n < cap) {
                    entries[2 * n] = pyramids_in_fov[p][j][1];
                    entries[2 * n + 1] = pyramids_in_fov[p][j][2];
                }
                ++n;
            }
    }
    offsets[observation_pyramid_num] = n;
    return n;
}
void ref_dump_neighbors(int *out) {
    std::memcpy(out, &observation_pyramid_neighbors[0][0], sizeof(observation_pyramid_neighbors));
}
// c[0] position cursor, c[1] velocity cursor, c[2] main-thread rand() calls, c[3] helper-thread rand() calls
void ref_cursors(int64_t *c) {
    c[0] = g_map->position_gaussian_random_seq;
    c[1] = g_map->velocity_gaussian_random_seq;
    c[2] = (int64_t)g_rand_main;
    c[3] = (int64_t)g_rand_helper;
}
void ref_gaussian_tables(float *p_out, float *v_out, int n) {
    std::memcpy(p_out, p_gaussian_randoms, sizeof(float) * n);
    std::memcpy(v_out, v_gaussian_randoms, sizeof(float) * n);
}
void ref_pdf_table(float *out) { std::memcpy(out, standard_gaussian_pdf, sizeof(float) * 20000); }
void ref_plane_normals(float *h, float *v) {
    std::memcpy(h, g_map->pyramid_BPnorm_params_h, sizeof(g_map->pyramid_BPnorm_params_h));
    std::memcpy(v, g_map->pyramid_BPnorm_params_v, sizeof(g_map->pyramid_BPnorm_params_v));
}
int ref_voxel_index(float x, float y, float z) {
    int idx = -1;
    return g_map->getPointVoxelsIndexPublic(x, y, z, idx) ? idx : -1;
}
void ref_voxel_center(int idx, float *xyz) { g_map->getVoxelPositionFromIndexPublic(idx, xyz[0], xyz[1], xyz[2]); }
float ref_uniform(float lo, float hi) { return DSPMap::generateRandomFloat(lo, hi); }

}  // extern "C"
