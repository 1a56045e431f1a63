// dspmap_dropin.hpp — `class DSPMap` with the reference's exact public surface (g-ch/DSP-map include/dsp_dynamic.h:142-446,
// 1549-1584) implemented over the C-ABI of the B200 library (include/dspmap_b200.h).  Included by this repository's
// dsp_dynamic.h / dsp_dynamic_multiple_neighbors.h / dsp_static.h after they have set the map parameters; the
// application (src/map_sim_example.cpp of the reference) compiles unchanged against them and links -ldspmap_b200.
//
// Like the reference headers this one pulls in <pcl/point_types.h>, "Eigen/Eigen" and `using namespace std;`, because the
// application relies on them (map_sim_example.cpp:39-57, 342, 443).
#pragma once
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iostream>
#include <random>
#include <string>
#include <thread>
#include <vector>
#include "Eigen/Eigen"
#include <pcl/point_types.h>
#ifndef DSPMAP_NO_PCL_EXTRAS  // the reference includes these too (dsp_dynamic.h:29-31); nothing from them is used here
#include <pcl/common/transforms.h>
#include <pcl_conversions/pcl_conversions.h>
#endif

#include "dspmap_b200.h"

using namespace std;  // dsp_dynamic.h:35

static const float prediction_future_time[PREDICTION_TIMES] = DSPMAP_FUTURE_TIMES;                            // :47
static const int observation_pyramid_num_h = (int)half_fov_h * 2 / ANGLE_RESOLUTION;                          // :58
static const int observation_pyramid_num_v = (int)half_fov_v * 2 / ANGLE_RESOLUTION;                          // :59
static const int observation_pyramid_num = observation_pyramid_num_h * observation_pyramid_num_v;             // :60
static const int VOXEL_NUM = MAP_LENGTH_VOXEL_NUM * MAP_WIDTH_VOXEL_NUM * MAP_HEIGHT_VOXEL_NUM;               // :62

class DSPMap {
public:
    DSPMap(int init_particle_num = 0, float init_weight = 0.01f) {  // :145
        dspmap_config c;
        dspmap_default_config(&c);
        c.nx = MAP_LENGTH_VOXEL_NUM;
        c.ny = MAP_WIDTH_VOXEL_NUM;
        c.nz = MAP_HEIGHT_VOXEL_NUM;
        c.resolution = VOXEL_RESOLUTION;
        c.angle_resolution = ANGLE_RESOLUTION;
        c.half_fov_h = half_fov_h;
        c.half_fov_v = half_fov_v;
        c.max_particles_per_voxel = MAX_PARTICLE_NUM_VOXEL;
        c.pyramid_neighbor_n = DSPMAP_PYRAMID_NEIGHBOR_N;
        c.model = DSPMAP_MODEL;
        c.prediction_times = PREDICTION_TIMES;
        for (int i = 0; i < PREDICTION_TIMES; ++i) c.prediction_future_time[i] = prediction_future_time[i];
        c.occlusion_margin = DSPMAP_OCCLUSION_MARGIN;
        c.pi_is_double = DSPMAP_PI_IS_DOUBLE;
        c.init_particle_num = init_particle_num;
        c.init_weight = init_weight;
        if (dspmap_create(&c, &map_) != DSPMAP_OK) {
            // the reference's constructor cannot fail; there is no CPU path to fall back to, so say why and stop
            cerr << "DSPMap: " << dspmap_last_error() << endl;
            std::abort();
        }
        instances().push_back(map_);
        dspmap_set_voxel_filter_resolution(map_, filter_resolution());
    }
    ~DSPMap() {  // :177
        for (size_t i = 0; i < instances().size(); ++i)
            if (instances()[i] == map_) { instances().erase(instances().begin() + i); break; }
        dspmap_destroy(map_);
        cout << "\n See you ;)" << endl;
    }
    DSPMap(const DSPMap &) = delete;
    DSPMap &operator=(const DSPMap &) = delete;

    int update(int point_cloud_num, int size_of_one_point, float *point_cloud_ptr, float sensor_px, float sensor_py,
               float sensor_pz, double time_stamp_second, float sensor_quaternion_w, float sensor_quaternion_x,
               float sensor_quaternion_y, float sensor_quaternion_z) {  // :181-184
        // the reference reads the global particle_save_folder when it writes the CSV (:333), not when the flag is set
        if (record_flag_) dspmap_set_particle_record_flag(map_, record_flag_, record_time_, (particle_save_folder + DSPMAP_CSV_SEPARATOR).c_str());
        int rc = dspmap_update(map_, point_cloud_num, size_of_one_point, point_cloud_ptr, sensor_px, sensor_py, sensor_pz,
                               time_stamp_second, sensor_quaternion_w, sensor_quaternion_x, sensor_quaternion_y,
                               sensor_quaternion_z);
        if (rc < 0) cout << "DSPMap::update: " << dspmap_last_error() << endl;
        return rc == DSPMAP_OK ? 1 : 0;
    }
    void setPredictionVariance(float p_stddev, float v_stddev) { dspmap_set_prediction_variance(map_, p_stddev, v_stddev); }  // :355
    void setObservationStdDev(float ob_stddev) { dspmap_set_observation_stddev(map_, ob_stddev); }                           // :362
    void setNewBornParticleWeight(float weight) { dspmap_set_newborn_weight(map_, weight); }                                 // :366
    void setNewBornParticleNumberofEachPoint(int num) { dspmap_set_newborn_number(map_, num); }                              // :370
    void setParticleRecordFlag(int record_particle_flag, float record_csv_time = 1.f) {                                      // :375
        record_flag_ = record_particle_flag;
        record_time_ = record_csv_time;
        dspmap_set_particle_record_flag(map_, record_particle_flag, record_csv_time, (particle_save_folder + DSPMAP_CSV_SEPARATOR).c_str());
    }
    static void setOriginalVoxelFilterResolution(float res) {  // :380 (static in the reference: applies to the process)
        filter_resolution() = res;
        for (dspmap *m : instances()) dspmap_set_voxel_filter_resolution(m, res);
    }
    void getOccupancyMap(int &obstacles_num, pcl::PointCloud<pcl::PointXYZ> &cloud, const float threshold = 0.7) {  // :385
        read(obstacles_num, cloud, nullptr, threshold);
    }
    void getOccupancyMapWithFutureStatus(int &obstacles_num, pcl::PointCloud<pcl::PointXYZ> &cloud, float *future_status,
                                         const float threshold = 0.7) {  // :405
        read(obstacles_num, cloud, future_status, threshold);
    }
    void clearOccupancyMapPrediction() { dspmap_clear_prediction(map_); }  // :431
    void getKMClusterResult(pcl::PointCloud<pcl::PointXYZINormal> &cluster_cloud) {  // :441
        int n = dspmap_get_tagged_cloud(map_, nullptr, 0);
        std::vector<float> buf((size_t)7 * (n > 0 ? n : 1));
        dspmap_get_tagged_cloud(map_, buf.data(), n);
        for (int i = 0; i < n; ++i) {
            pcl::PointXYZINormal p;
            p.x = buf[7 * i]; p.y = buf[7 * i + 1]; p.z = buf[7 * i + 2];
            p.normal_x = buf[7 * i + 3]; p.normal_y = buf[7 * i + 4]; p.normal_z = buf[7 * i + 5];
            p.intensity = buf[7 * i + 6];
            cluster_cloud.push_back(p);
        }
    }
    // public in the reference (dsp_dynamic.h:795-796); the newborn step is part of update() here, so this is a no-op kept
    // for source compatibility
    void mapAddNewBornParticlesByObservation() {}
    static float generateRandomFloat(float min, float max) {  // :1551
        return min + static_cast<float>(rand()) / (static_cast<float>(RAND_MAX / (max - min)));
    }
    void getVoxelPositionFromIndexPublic(const int &index, float &px, float &py, float &pz) const {  // :1556
        float c[3];
        dspmap_voxel_center(map_, index, c);
        px = c[0]; py = c[1]; pz = c[2];
    }
    int getPointVoxelsIndexPublic(const float &px, const float &py, const float &pz, int &index) {  // :1574
        return dspmap_voxel_index(map_, px, py, pz, &index);
    }
    dspmap *handle() { return map_; }

private:
    static std::vector<dspmap *> &instances() { static std::vector<dspmap *> v; return v; }
    static float &filter_resolution() { static float r = 0.15f; return r; }  // voxel_filtered_resolution (:132)
    void read(int &obstacles_num, pcl::PointCloud<pcl::PointXYZ> &cloud, float *future_status, float threshold) {
        if (xyz_.empty()) xyz_.resize((size_t)3 * VOXEL_NUM);
        if (future_status && future_status != pinned_) {  // the application's array is static (ex:371): pin it once
            dspmap_pin_host_buffer(map_, future_status, sizeof(float) * (size_t)VOXEL_NUM * PREDICTION_TIMES);
            pinned_ = future_status;
        }
        int n = 0;
        dspmap_get_occupancy(map_, threshold, xyz_.data(), VOXEL_NUM, &n, future_status);
        obstacles_num = n;
        for (int i = 0; i < n; ++i) {  // appended, not cleared (:391,:411)
            pcl::PointXYZ p;
            p.x = xyz_[3 * i]; p.y = xyz_[3 * i + 1]; p.z = xyz_[3 * i + 2];
            cloud.push_back(p);
        }
    }
    dspmap *map_ = nullptr;
    int record_flag_ = 0;
    float record_time_ = 1.f;
    std::vector<float> xyz_;
    float *pinned_ = nullptr;
};
