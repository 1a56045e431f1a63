// dspmap_replay — ROS-free stand-in for the reference's example node (src/map_sim_example.cpp): reads a recorded
// stream of (cloud, pose, time stamp) frames, runs DSPMap::update + getOccupancyMapWithFutureStatus through this
// repository's drop-in header, and writes per-frame results.  SURVEY.md §8(f) row 1.
//
//   dspmap_replay <stream.bin> [--out prefix] [--threshold 0.2] [--csv-at-frame k]
//
// stream.bin : int32 frames; per frame: int32 n, float32 pos[3], float32 quat[4] (w x y z), float64 t, float32 xyz[n*3]
//              (points in the sensor frame, as the application hands them to update(), map_sim_example.cpp:320-349)
// outputs    : <prefix>_frame<k>.occ  (float32 xyz triples of occupied voxel centres, ascending voxel index)
//              <prefix>_frame<k>.fut  (float32 [VOXEL_NUM][PREDICTION_TIMES], only with --future)
//              particle CSV in the reference's column order (flag,vx,vy,vz,px,py,pz,weight,voxel; dsp_dynamic.h:339-344)
//              when --csv-at-frame is given (setParticleRecordFlag)
#ifndef DSPMAP_HEADER
#define DSPMAP_HEADER "dsp_dynamic.h"
#endif
#include DSPMAP_HEADER
#include <chrono>
#include <cstdio>
#include <cstring>

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s stream.bin [--out prefix] [--threshold t] [--future] [--csv-at-frame k]\n", argv[0]);
        return 2;
    }
    string prefix = "replay";
    float threshold = 0.2f;  // map_sim_example.cpp:378
    bool write_future = false;
    int csv_frame = -1;
    for (int i = 2; i < argc; ++i) {
        if (!strcmp(argv[i], "--out") && i + 1 < argc) prefix = argv[++i];
        else if (!strcmp(argv[i], "--threshold") && i + 1 < argc) threshold = (float)atof(argv[++i]);
        else if (!strcmp(argv[i], "--future")) write_future = true;
        else if (!strcmp(argv[i], "--csv-at-frame") && i + 1 < argc) csv_frame = atoi(argv[++i]);
    }
    DSPMap map;
    map.setPredictionVariance(0.05, 0.05);  // map_sim_example.cpp:522-526
    map.setObservationStdDev(0.1);
    map.setNewBornParticleNumberofEachPoint(20);
    map.setNewBornParticleWeight(0.0001);
    DSPMap::setOriginalVoxelFilterResolution(0.1f);
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 3; }
    int frames = 0;
    if (fread(&frames, 4, 1, f) != 1) return 4;
    vector<float> pts, future((size_t)VOXEL_NUM * PREDICTION_TIMES);
    double total = 0;
    int done = 0;
    for (int k = 0; k < frames; ++k) {
        int n;
        float pose[7];
        double t;
        if (fread(&n, 4, 1, f) != 1 || fread(pose, 4, 7, f) != 7 || fread(&t, 8, 1, f) != 1) return 5;
        pts.resize((size_t)3 * n);
        if (n && fread(pts.data(), 4, (size_t)3 * n, f) != (size_t)3 * n) return 6;
        if (k == csv_frame) map.setParticleRecordFlag(-1);
        auto t0 = std::chrono::steady_clock::now();
        int ok = map.update(n, 3, pts.data(), pose[0], pose[1], pose[2], t, pose[3], pose[4], pose[5], pose[6]);
        int occupied = 0;
        pcl::PointCloud<pcl::PointXYZ> cloud;
        if (ok) map.getOccupancyMapWithFutureStatus(occupied, cloud, future.data(), threshold);
        auto t1 = std::chrono::steady_clock::now();
        if (k == csv_frame) map.setParticleRecordFlag(0);
        if (!ok) { printf("frame %d rejected\n", k); continue; }
        double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        if (k >= 5) { total += ms; ++done; }
        char name[512];
        snprintf(name, sizeof(name), "%s_frame%04d.occ", prefix.c_str(), k);
        if (FILE *o = fopen(name, "wb")) {
            for (int i = 0; i < occupied; ++i) {
                const float xyz[3] = {cloud.points[i].x, cloud.points[i].y, cloud.points[i].z};
                fwrite(xyz, 4, 3, o);
            }
            fclose(o);
        }
        if (write_future) {
            snprintf(name, sizeof(name), "%s_frame%04d.fut", prefix.c_str(), k);
            if (FILE *o = fopen(name, "wb")) { fwrite(future.data(), 4, future.size(), o); fclose(o); }
        }
        printf("frame %d: %d points, %d occupied voxels, %.3f ms\n", k, n, occupied, ms);
    }
    fclose(f);
    if (done) printf("****** Map avg time %f seconds over %d frames\n", total / done / 1e3, done);  // map_sim_example.cpp:361
    return 0;
}
