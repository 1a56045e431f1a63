// A ROS-free stand-in for the reference application's per-frame code (src/map_sim_example.cpp:39-57, 345-427, 522-528):
// the same calls, macros and types, compiled against this repository's drop-in header.  Reads a binary stream
// (tests write it), prints one summary line per frame.  Built and run by tests/test_dropin.py.
#include DSPMAP_HEADER
#include <cstdio>

DSPMap my_map;  // global object, like map_sim_example.cpp:39
const float res = 0.1;
float x_min = -MAP_LENGTH_VOXEL_NUM * VOXEL_RESOLUTION / 2;  // map_sim_example.cpp:52-57
float x_max = MAP_LENGTH_VOXEL_NUM * VOXEL_RESOLUTION / 2;

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    my_map.setPredictionVariance(0.05, 0.05);
    my_map.setObservationStdDev(0.1);
    my_map.setNewBornParticleNumberofEachPoint(20);
    my_map.setNewBornParticleWeight(0.0001);
    DSPMap::setOriginalVoxelFilterResolution(res);
    my_map.setParticleRecordFlag(0, 19.0);
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    int frames = 0;
    if (fread(&frames, 4, 1, f) != 1) return 4;
    static float future_status[VOXEL_NUM][PREDICTION_TIMES];
    vector<float> pts;
    for (int k = 0; k < frames; ++k) {
        int n;
        float pose[7];
        double t;
        if (fread(&n, 4, 1, f) != 1 || fread(pose, 4, 7, f) != 7 || fread(&t, 8, 1, f) != 1) return 5;
        pts.resize((size_t)3 * n);
        if (n && fread(pts.data(), 4, (size_t)3 * n, f) != (size_t)3 * n) return 6;
        if (!my_map.update(n, 3, pts.data(), pose[0], pose[1], pose[2], t, pose[3], pose[4], pose[5], pose[6])) {
            printf("frame %d rejected\n", k);
            continue;
        }
        int occupied_num = 0;
        pcl::PointCloud<pcl::PointXYZ> cloud_to_publish;
        my_map.getOccupancyMapWithFutureStatus(occupied_num, cloud_to_publish, &future_status[0][0], 0.2);
        double fsum = 0;
        for (int i = 0; i < VOXEL_NUM; ++i)
            for (int j = 0; j < PREDICTION_TIMES; ++j) fsum += future_status[i][j];
        float px, py, pz;
        my_map.getVoxelPositionFromIndexPublic(VOXEL_NUM / 2, px, py, pz);
        pcl::PointCloud<pcl::PointXYZINormal> km;
        my_map.getKMClusterResult(km);
        printf("frame %d occupied %d cloud %zu future_sum %.6f tagged %zu first %.4f %.4f %.4f\n", k, occupied_num,
               cloud_to_publish.size(), fsum, km.size(), occupied_num ? cloud_to_publish.points.at(0).x : 0.f,
               occupied_num ? cloud_to_publish.points.at(0).y : 0.f, occupied_num ? cloud_to_publish.points.at(0).z : 0.f);
    }
    fclose(f);
    return 0;
}
