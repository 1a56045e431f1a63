/* dspmap_b200.h — C-ABI of the B200-native DSP map (drop-in boundary for the per-frame particle loop).
 *
 * Every entry point names the reference interface it replaces (g-ch/DSP-map, include/dsp_dynamic.h unless
 * stated otherwise).  Plain C types only: no PCL / Eigen / torch types cross this boundary.  The C++ class
 * `DSPMap` in include/dsp_dynamic.h (this repository's drop-in header) is a thin wrapper over these calls.
 *
 * Conventions
 *   - All functions returning `int` return DSPMAP_OK (1) / a frame-rejected 0 where the reference returns 0,
 *     or a negative DSPMAP_E_* code for conditions the reference cannot have (CUDA failure, bad handle).
 *     Nothing throws.  dspmap_last_error() gives a human-readable string for the last negative code.
 *   - The library is not thread-safe per handle (neither is the reference: file-static state, dsp_dynamic.h:112-140).
 *   - There is NO CPU fallback: if no sm_100 device is usable, dspmap_create fails with DSPMAP_E_NO_DEVICE.
 */
#ifndef DSPMAP_B200_H
#define DSPMAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSPMAP_OK 1
#define DSPMAP_REJECTED 0        /* update(): bad quaternion / pose jump / time jump (dsp_dynamic.h:193-208) */
#define DSPMAP_E_NO_DEVICE (-1)
#define DSPMAP_E_CUDA (-2)
#define DSPMAP_E_BAD_ARG (-3)
#define DSPMAP_E_CAPACITY (-4)

#define DSPMAP_MAX_PREDICTION_TIMES 8
#define DSPMAP_TAGGED_STRIDE 7 /* x y z vx vy vz intensity: one point of input_cloud_with_velocity (dsp_dynamic.h:134) */

typedef struct dspmap dspmap; /* opaque handle */

/* Compile-time parameters of the reference (dsp_dynamic.h:38-70, mn:38-51, st:38-63) as run-time fields. */
typedef struct dspmap_config {
    int32_t nx, ny, nz;                 /* MAP_LENGTH/WIDTH/HEIGHT_VOXEL_NUM                 dsp_dynamic.h:38-40 */
    float resolution;                   /* VOXEL_RESOLUTION                                  :41 */
    int32_t angle_resolution;           /* ANGLE_RESOLUTION (degrees)                        :42 */
    int32_t half_fov_h, half_fov_v;     /* half_fov_h / half_fov_v (degrees)                 :49-50 */
    int32_t max_particles_per_voxel;    /* MAX_PARTICLE_NUM_VOXEL                            :43 */
    int32_t safe_particles_per_voxel;   /* SAFE_PARTICLE_NUM_VOXEL; 0 = derive (2x, static 5x) :65, st:63 */
    int32_t safe_particles_per_pyramid; /* SAFE_PARTICLE_NUM_PYRAMID; 0 = derive             :66 */
    int32_t pyramid_neighbor_n;         /* PYRAMID_NEIGHBOR_N; 1 = the 3x3 block of dsp_dynamic.h:1135, mn:43 */
    int32_t model;                      /* 0 = constant velocity (dsp_dynamic.h), 1 = static (dsp_static.h) */
    int32_t prediction_times;           /* PREDICTION_TIMES                                  :46 */
    float prediction_future_time[DSPMAP_MAX_PREDICTION_TIMES]; /*                           :47 */
    float occlusion_margin;             /* obstacle_thickness_for_occlusion 0.3 (:70,:761); voxel_resolution in mn:761 / st */
    int32_t init_particle_num;          /* DSPMap(init_particle_num, init_weight)            :145 */
    float init_weight;
    uint64_t table_seed;                /* seed of the two Gaussian tables; the reference uses time(NULL) (:1151) */
    uint64_t uniform_seed;              /* seed of the counter-based uniform stream that stands where rand() is (:1552) */
    int32_t gaussian_table_size;        /* GAUSSIAN_RANDOMS_NUM 10000000                     :72 */
    int32_t max_observations_per_pyramid; /* observation_max_points_num_one_pyramid 100      :69 */
    int32_t device;                     /* CUDA device ordinal */
    int32_t max_points;                 /* capacity for points per update() (0 = 65536) */
    /* voxel-subspace shard owned by this handle (multi-GPU): z-layers [z_begin, z_end); 0,0 = whole map */
    int32_t shard_z_begin, shard_z_end;
    /* How `(float)ANGLE_RESOLUTION / 180.f * M_PIf32` (:543) rounds.  dsp_dynamic.h keeps glibc's M_PIf32, a FLOAT literal
     * (its own double fallback at :77-79 is skipped because <cmath> defines the macro under _GNU_SOURCE): fp32 product, 0.
     * dsp_dynamic_multiple_neighbors.h:78 and dsp_static.h:74 redefine M_PIf32 as a DOUBLE literal: fp64 product rounded to
     * fp32 once, 1.  The boundary-plane normals, hence pyramid ids of particles within an ulp of a plane, depend on it. */
    int32_t pi_is_double;
} dspmap_config;

/* Fills `c` with the values of the reference tree as shipped (dsp_dynamic.h:38-50,145-168). */
void dspmap_default_config(dspmap_config *c);

/* DSPMap::DSPMap (:145-175) + setInitParameters (:525-591) + addRandomParticles (:594-624). */
int dspmap_create(const dspmap_config *cfg, dspmap **out);
/* DSPMap::~DSPMap (:177). */
void dspmap_destroy(dspmap *m);
const char *dspmap_last_error(void);

/* DSPMap::update (:181-353).  `pts` is HOST memory, n points of `stride` floats, first three used (:247,:289).
 * Runs the whole frame: binning, prediction + voxel/pyramid reassignment, observation weight update, the host
 * velocity-estimation step, newborn particles, occupancy + resampling + future status.  Returns 1 / 0 like the
 * reference. */
int dspmap_update(dspmap *m, int n, int stride, const float *pts, float px, float py, float pz, double t,
                  float qw, float qx, float qy, float qz);

/* Same frame, but the newborn input (the side thread's output input_cloud_with_velocity, :134,:815) is given
 * explicitly: n_tagged points of DSPMAP_TAGGED_STRIDE floats in the WORLD frame, as getKMClusterResult (:441)
 * returns them.  tagged == NULL keeps the previous frame's cloud (what the reference does when no point is in
 * view, :1379). */
int dspmap_update_tagged(dspmap *m, int n, int stride, const float *pts, float px, float py, float pz, double t,
                         float qw, float qx, float qy, float qz, const float *tagged, int n_tagged);

/* Device-resident variant of dspmap_update_tagged for callers whose clouds already live in HBM (no host copies,
 * no synchronisation; the frame is only enqueued on the handle's stream). d_pts: n*3 floats, d_tagged: n_tagged*7. */
int dspmap_update_device(dspmap *m, int n, const float *d_pts, float px, float py, float pz, double t, float qw,
                         float qx, float qy, float qz, const float *d_tagged, int n_tagged);

/* setPredictionVariance (:355-360; regenerates both Gaussian tables, cursors are kept),
 * setObservationStdDev (:362), setNewBornParticleWeight (:366), setNewBornParticleNumberofEachPoint (:370),
 * setParticleRecordFlag (:375), setOriginalVoxelFilterResolution (:380). */
int dspmap_set_prediction_variance(dspmap *m, float p_stddev, float v_stddev);
int dspmap_set_observation_stddev(dspmap *m, float ob_stddev);
int dspmap_set_newborn_weight(dspmap *m, float weight);
int dspmap_set_newborn_number(dspmap *m, int num);
/* The particle CSV (:325-350: one line flag,vx,vy,vz,px,py,pz,weight,voxel per live particle, sweep order) is written to
 * <prefix>particles_update_t_<update counter>_<ms>.csv; the reference's headers differ in the prefix (dsp_dynamic.h:333
 * particle_save_folder + "/", the other two without the separator), so the drop-in headers pass it complete. */
int dspmap_set_particle_record_flag(dspmap *m, int flag, float record_time, const char *prefix);
int dspmap_set_voxel_filter_resolution(dspmap *m, float res);

/* getOccupancyMap (:385-402) when future == NULL, getOccupancyMapWithFutureStatus (:405-426) otherwise.
 * Writes up to `cap` voxel centres (xyz triples, ascending voxel index) to xyz_out, the total count to *n_out,
 * V*T floats ([voxel][horizon]) to `future`, and zeroes the future columns (the reference's side effect,
 * :397-400,:421-424).  HOST pointers. */
int dspmap_get_occupancy(dspmap *m, float threshold, float *xyz_out, int cap, int *n_out, float *future);
/* Device-resident variant: d_xyz (cap*3 floats), d_count (1 int), d_future (V*T floats or NULL) are device
 * pointers; only enqueued. */
int dspmap_get_occupancy_device(dspmap *m, float threshold, float *d_xyz, int cap, int *d_count, float *d_future);
/* Pipelined reader (zero-copy readers): same results and the same side effect as dspmap_get_occupancy with
 * xyz capacity V, but the call only enqueues — the reader kernels on the map's stream, the device-to-host copies on a second
 * stream into page-locked buffers owned by the library — and returns a ticket (0 or 1; two result slots alternate).
 * dspmap_wait_occupancy blocks until that ticket's copies have landed and hands out HOST pointers into the slot: *n voxel
 * centres in xyz, V*T floats in future (NULL if with_future was 0).  They stay valid until the slot is reused, i.e. until the
 * second dspmap_get_occupancy_async call after this one.  Typical loop: update(k); get_async(k); wait(k-1); consume(k-1) — the
 * 4 B * V * T copy of frame k-1 overlaps update(k). */
int dspmap_get_occupancy_async(dspmap *m, float threshold, int with_future, int *ticket);
int dspmap_wait_occupancy(dspmap *m, int ticket, const float **xyz, int *n, const float **future);
/* Optional: page-locks a caller-owned output buffer (e.g. the application's static future_status array) so that
 * dspmap_get_occupancy can DMA straight into it instead of staging + memcpy. The buffer must outlive the handle or be
 * released with bytes == 0.
 * With DSPMAP_SPARSE_FUTURE=1 in the environment a registered future_status buffer is updated incrementally: only the
 * voxel rows with a non-zero future value cross PCIe (2-3 % of the grid) and the rows of the previous call are cleared on
 * the host, which yields the same V x T array as the dense copy.  While registered, the application must then treat the
 * buffer as read-only between calls (the reference application only reads it, map_sim_example.cpp:371-437). */
int dspmap_pin_host_buffer(dspmap *m, void *ptr, size_t bytes);
/* clearOccupancyMapPrediction (:431-438). */
/* Bytes the last dspmap_get_occupancy call moved from the device to the host (count, occupied list incl. its speculative
 * prefix, and the future grid: dense, or only its non-zero rows when the caller's array is registered). */
long long dspmap_last_reader_bytes(dspmap *m);
/* Bytes the last dspmap_update / dspmap_update_tagged call moved over PCIe: the cloud plus either the velocity-tagged newborn
 * input (host estimator, explicit input) or the matched clusters' velocities (device estimator); the cluster features the
 * device hands to the host matching. */
void dspmap_last_update_bytes(dspmap *m, long long *h2d, long long *d2h);
/* Counters of the last frame's velocity-estimation front end when it ran on the device (dspmap_estimator.cuh): points in view,
 * clusters of admissible size, ground points, dynamic clusters, their points, points of static clusters, occupied grid cells,
 * size of the tagged cloud.  Returns 1 if the map estimates on the device (DSPMAP_EST_GPU=1), 0 if on the host (default). */
int dspmap_estimator_stats(dspmap *m, int32_t *out8);
/* Diagnosis (maps created with DSPMAP_TIMELINE=1): milliseconds from the start of the last frame (behind k_frame_setup) to
 * the cluster features reaching the host, the arrival pass, the end of the observation binning, the early newborn placement,
 * the start of the weight pass, the newborn normaliser and the end of the frame; -1 for events the frame did not record. */
int dspmap_timeline(dspmap *m, float *out8);
int dspmap_clear_prediction(dspmap *m);
/* getKMClusterResult (:441-445): copies the last newborn input; returns the number of points. */
int dspmap_get_tagged_cloud(dspmap *m, float *out, int cap);

/* getVoxelPositionFromIndexPublic (:1556-1572) / getPointVoxelsIndexPublic (:1574-1584): pure host arithmetic. */
void dspmap_voxel_center(const dspmap *m, int index, float *xyz);
int dspmap_voxel_index(const dspmap *m, float x, float y, float z, int *index);
/* generateRandomFloat (:1551-1553) on the handle's uniform stream. */
float dspmap_uniform(dspmap *m, float lo, float hi);

/* Sizes: V, S, P, L, T, Nh, Nv, neighbour-table width, MAX ppv, nx, ny, nz, obs max, model, then two flags: exact fast
 * division verified for the voxel size / for sigma (16 ints). */
void dspmap_dims(const dspmap *m, int32_t *out);

/* State dump / load (tests, checkpointing).  Particle record = the reference's CSV columns (:339-344):
 * ids[n][2] = {voxel, slot}; vals[n][8] = {flag, vx, vy, vz, px, py, pz, weight}; sweep order on dump. */
int dspmap_dump_particles(dspmap *m, int32_t *ids, float *vals, int cap);
int dspmap_load_particles(dspmap *m, const int32_t *ids, const float *vals, int n);
/* voxels_objects_number[V][4+T] (:120). */
int dspmap_dump_voxel_objects(dspmap *m, float *out);
/* Last frame's binned observations (:497-515): counts[P], maxlen[P], pts[P][obs_max][5] (x y z Cz range). */
int dspmap_dump_observations(dspmap *m, int32_t *counts, float *maxlen, float *pts);
/* Last frame's pyramid lists (:124): offsets[P+1], entries[n][2] = {voxel, slot} in list order. */
int dspmap_dump_pyramid_lists(dspmap *m, int32_t *offsets, int32_t *entries, int cap);
/* The boundary-plane normals of the last frame, rotated to the sensor attitude (:226-232): h[(Nh+1)*3], v[(Nv+1)*3]. */
int dspmap_dump_plane_normals(dspmap *m, float *h, float *v);
/* c[0] position-noise cursor, c[1] velocity-noise cursor, c[2] uniform-stream counter (:483-484). */
int dspmap_cursors(dspmap *m, int64_t *c);
int dspmap_set_cursors(dspmap *m, int64_t p, int64_t v, int64_t u);

/* Per-frame counters of the last update: n_in, n_left_map, n_voxel_full, n_pyramid_full, n_moved,
 * n_fov, n_candidates, n_born, n_low_weight, n_pre, n_old, n_out, n_valid_points, capacity-overrun code (0 = none; dspmap.cu: report_overflow), kernel launches of
 * the last frame, kernel launches since create (16 int64). */
int dspmap_counters(dspmap *m, int64_t *out);

/* Debug / replay: overrides the remembered previous pose and time stamp (the function-local statics of update(),
 * dsp_dynamic.h:187-190) so that a frame can be replayed from an injected state. */
int dspmap_set_last_pose(dspmap *m, float px, float py, float pz, double t);
/* Debug: run only the first k stages of the next update()s (1 predict, 2 +observe, 3 +newborn, >=4 all). */
int dspmap_set_stage_limit(dspmap *m, int k);
/* Work is enqueued on this CUDA stream (a cudaStream_t; 0 = the handle's own stream). */
int dspmap_set_stream(dspmap *m, void *cuda_stream);
int dspmap_synchronize(dspmap *m);
/* Average device time (ms) of each kernel family since the last reset, measured with CUDA events when profiling is
 * enabled: names[i] / ms[i] / launches[i]; returns the number of families. */
int dspmap_profile_enable(dspmap *m, int on);
int dspmap_profile_read(dspmap *m, const char **names, float *ms, int32_t *launches, int cap);
/* The same event measurement per kernel (name as written at the launch site); returns the number of kernels. */
int dspmap_profile_read_kernels(dspmap *m, const char **names, float *ms, int32_t *launches, int cap);

/* ---- voxel-subspace sharding over the GPUs of one box (one handle per GPU / rank) -----------------------------------
 * The map's z layers are cut into `nranks` equal slabs; handle `rank` owns the particles of its slab.  A frame is
 * enqueued in six phases with five collectives between them, issued by the caller on the same stream
 * (dsp-map_b200/dspmap_b200/sharded.py does it with torch.distributed / NCCL):
 *   phase 0  binning, prediction; movers that leave the slab are written to xsend[dest rank]
 *            -> all-to-all of the fixed-size slabs xsend -> xrecv                       (the boundary exchange)
 *   phase 1  imported movers join the ordered arrival replay; registered particles are packed into gsend
 *            -> all-gather of the slab headers (counts), dspmap_shard_gather_records(max count), all-gather gsend -> grecv
 *   phase 2  identical global pyramid lists on every rank; C_z of the point pyramids i with i % nranks == rank
 *            -> all-reduce(sum) of czinv (zero-initialised, one writer per element: exact)
 *   phase 3  weights of the particle chunks c with c % nranks == rank (whoever owns the particles)
 *            -> all-reduce(sum) of the weight part of `shared`
 *   phase 4  owners apply the new weights; newborn split of the points whose voxel this rank owns
 *            -> all-reduce(sum) of the split part of `shared`
 *   phase 5  newborn candidates that land in the slab; occupancy, resampling, future status
 * Buffers are device memory owned by the caller: xsend/xrecv hold nranks slabs of (4 + cap_x*12) floats, gsend one and
 * grecv nranks slabs of (4 + cap_g*8) floats, czinv P*obs_max + max_points floats, shared max_points + nranks*cap_g
 * floats.  Results are bit-identical to a single handle. */
int dspmap_shard_config(dspmap *m, int rank, int nranks, float *xsend, float *xrecv, int cap_x, float *gsend, float *grecv,
                        int cap_g, float *czinv, float *shared);
/* Records per rank actually moved by this frame's all-gather (>= the largest count, <= cap_g): sets the slab stride of
 * grecv for phases 2..4. */
int dspmap_shard_gather_records(dspmap *m, int records);
int dspmap_shard_phase(dspmap *m, int phase, int n, const float *d_pts, float px, float py, float pz, double t, float qw,
                       float qx, float qy, float qz, const float *d_tagged, int n_tagged);

/* ---- the same, orchestrated by the library (C++ host): one call per frame, collectives inside -------------------------------
 * One process + one handle per GPU; the six phases and their collectives are issued from C++ on the handle's stream with
 * NCCL (libnccl.so.2 is opened with dlopen by dspmap_shard_init: no link-time dependency).  One host synchronisation per
 * frame, with device work queued behind it: the headers of the boundary exchange let every rank compute the same upper
 * bound of any rank's registered particles, which sizes the all-gather exactly.  Exchange buffers belong to the
 * library.  cap_x / cap_g <= 0 pick defaults (4096 crossers per rank pair; live-list capacity / nranks).
 *   rank 0:      dspmap_shard_unique_id(id)  ->  hand the DSPMAP_NCCL_ID_BYTES bytes to every rank (MPI, a file, a socket, ...)
 *   every rank:  dspmap_create(...); dspmap_shard_init(m, rank, nranks, id, 0, 0);
 *   per frame:   dspmap_shard_update(m, ...)       cloud, pose and newborn input replicated, device memory
 *                dspmap_shard_get_occupancy(m, ...) every rank receives the whole map's occupied list and future grid
 * The *_local variants drive all handles of one map in ONE process on ONE device (shared stream; the collectives are
 * device-to-device copies): the same orchestration code, testable on a single GPU. */
#define DSPMAP_NCCL_ID_BYTES 128
int dspmap_shard_unique_id(void *id128);
int dspmap_shard_init(dspmap *m, int rank, int nranks, const void *id128, int cap_x, int cap_g);
int dspmap_shard_init_local(dspmap **handles, int n, int cap_x, int cap_g);
int dspmap_shard_update(dspmap *m, int n, const float *d_pts, float px, float py, float pz, double t, float qw, float qx,
                        float qy, float qz, const float *d_tagged, int n_tagged);
int dspmap_shard_update_local(dspmap **handles, int n_handles, int n, const float *d_pts, float px, float py, float pz, double t,
                              float qw, float qx, float qy, float qz, const float *d_tagged, int n_tagged);
int dspmap_shard_get_occupancy(dspmap *m, float threshold, float *d_xyz, int cap, int *d_count, float *d_future);
int dspmap_shard_get_occupancy_local(dspmap **handles, int n_handles, float threshold, float *const *d_xyz, int cap,
                                     int *const *d_count, float *const *d_future);
/* out4: cap_x, cap_g, records per rank moved by the last frame's all-gather, frames so far. */
int dspmap_shard_info(dspmap *m, int32_t *out4);

/* Host-only access to the velocity-estimation step (the reference's side thread, dsp_dynamic.h:1377-1544; static variant
 * dsp_static.h:1285-1309) without a map or a GPU: used to pre-compute newborn inputs for device-resident streams and
 * by the CPU tests.  `estimate` consumes one frame (n x 3 points, sensor frame) and writes the tagged cloud (7 floats
 * per point, world frame); it returns the number of points, or -1 when nothing is in view (the previous cloud stays
 * valid, :1379). */
typedef struct dspmap_estimator dspmap_estimator;
dspmap_estimator *dspmap_estimator_create(const dspmap_config *cfg, float voxel_filter_resolution);
void dspmap_estimator_destroy(dspmap_estimator *e);
int dspmap_estimator_estimate(dspmap_estimator *e, int n, const float *pts, float px, float py, float pz, float dt,
                              float qw, float qx, float qy, float qz, float *out, int cap);
/* Runs `estimate` on a persistent helper thread (the hand-over dspmap_update uses when DSPMAP_EST_THREAD=1 moves the
 * estimation off the enqueueing thread, like the reference's std::thread, dsp_dynamic.h:297-311); results are identical. */
int dspmap_estimator_set_threaded(dspmap_estimator *e, int on);

/* Application-side preprocessing on the GPU: what src/map_sim_example.cpp:305-336 does per depth frame
 * before DSPMap::update — pcl::VoxelGrid down-sampling with leaf size `leaf` (ex:312-316), the camera -> map axis swap
 * x = z, y = -x, z = -y (ex:320-322), the open-interval crop to (range_min, range_max) (ex:325) and the cut at `cap`
 * points (ex:332-334) — so that a raw depth cloud crosses PCIe once.  Leaf membership and output order (ascending PCL leaf
 * index) are exact; a leaf's centroid is accumulated exactly in 2^-24 m fixed point and rounded once (PCL's own fp32 running
 * sum follows an unstable sort, i.e. is only defined up to its rounding error; see prefilter.cu).
 *   max_raw_floats: capacity of the raw staging buffer (points x stride); max_out_points: capacity of the result;
 *   max_leaves: capacity of the leaf grid over the cloud's bounding box (0 = 2^22; 28 B of HBM each).
 * dspmap_prefilter_run: HOST pointers; `pts` = n points `stride` floats apart (camera frame, NaN / inf allowed and skipped);
 * writes *n_out <= cap points to `out`.  DSPMAP_E_CAPACITY when the bounding box needs more than max_leaves leaves.
 * dspmap_prefilter_run_device: DEVICE pointers, enqueue only; *d_n_out = -1 signals the capacity error.
 * dspmap_update_raw = dspmap_prefilter_run + dspmap_update on the result (no intermediate copy to caller memory). */
typedef struct dspmap_prefilter dspmap_prefilter;
int dspmap_prefilter_create(int device, int max_raw_floats, int max_out_points, long long max_leaves, dspmap_prefilter **out);
void dspmap_prefilter_destroy(dspmap_prefilter *p);
int dspmap_prefilter_set_stream(dspmap_prefilter *p, void *cuda_stream);
const char *dspmap_prefilter_last_error(void);
long long dspmap_prefilter_launches(const dspmap_prefilter *p);
int dspmap_prefilter_run(dspmap_prefilter *p, int n, int stride, const float *pts, float leaf, const float *range_min,
                         const float *range_max, float *out, int cap, int *n_out);
int dspmap_prefilter_run_device(dspmap_prefilter *p, int n, int stride, const float *d_pts, float leaf, const float *range_min,
                                const float *range_max, float *d_out, int cap, int *d_n_out);
int dspmap_update_raw(dspmap *m, dspmap_prefilter *p, int n, int stride, const float *raw, float leaf, const float *range_min,
                      const float *range_max, float px, float py, float pz, double t, float qw, float qx, float qy, float qz,
                      int *n_filtered);

/* The estimator's Euclidean clustering on its own (stand-in for pcl::EuclideanClusterExtraction as the side thread uses
 * it, dsp_dynamic.h:1407-1417): labels[i] = index of point i's cluster in the output order (size descending, ties by
 * smallest member index), or -1 when its component is outside [min_size, max_size].  path: 0 = automatic, 1 = hash
 * grid, 2 = dense grid (fails with DSPMAP_E_CAPACITY when the bounding box is too large).  Returns the cluster count. */
int dspmap_euclidean_clusters(const float *xyz, int n, float tolerance, int min_size, int max_size, int path, int *labels);

#ifdef __cplusplus
}
#endif
#endif /* DSPMAP_B200_H */
