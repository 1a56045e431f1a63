// dspmap_types.h — shared host/device types of the B200 DSP map (product code).
#pragma once
#include <stdint.h>

typedef unsigned long long u64;

#define DSP_MAX_T 8
#define DSP_MAX_SLOTS 128   // S <= 128: one voxel's occupancy is a 128-bit mask
#define DSP_KEY_SHIFT 7     // sweep key = voxel * 128 + slot  (monotone in the reference's (voxel, slot) sweep order)
#define DSP_LUT_HALF 10001  // the PDF table is symmetric about index 10000 (dsp_dynamic.h:1282-1292)
#define DSP_MAX_NB_NUM 64   // newborn candidates per point are tracked in a 64-bit mask

// Constant for the life of a map. Passed to kernels by value.
struct MapConst {
    int V, S, P, L, T, Nh, Nv, NB, NBW, nx, ny, nz, OBS, max_ppv, model, G;
    int z_begin, z_end;  // voxel-subspace shard (z layers) owned by this handle
    float hx, hy, hz, res, occl;
    float ft[DSP_MAX_T];
    u64 vlo, vhi;  // valid slot bits (slot < S)
    float res_r;   // RN(1/res); used only when fast_res (exhaustively verified at create time, see k_verify_div)
    int fast_res;
    long long cap_pairs;  // capacity of the pair buffer G
    // voxel-subspace sharding (multi-GPU): this handle owns voxels [v_lo, v_hi) = z layers [rank * z_per_rank, ...)
    int sharded, rank, nranks, z_per_rank, v_lo, v_hi;
    int cap_x, cap_g;  // records per exchange slab (boundary crossers / registered particles)
};

// Per-frame scalars. Passed to kernels by value.
struct FrameConst {
    float q[4];    // sensor attitude (w x y z)
    float qi[4];   // its inverse (conjugate / squared norm), host-computed with the reference's arithmetic
    float sx, sy, sz;  // particle shift = -(odometry delta)   (dsp_dynamic.h:300)
    float dt;
    float cur[3];  // current sensor position                 (dsp_dynamic.h:213-215)
    float sigma, Pd, one_minus_Pd, kappa;
    float nb_weight;
    int nb_num, nb_min_static, nb_model_gen;
    int n_points, n_tagged;
    int stage_limit;
    float sigma_r;  // RN(1/sigma), used only when fast_sigma (exhaustively verified, see k_verify_div)
    int fast_sigma;
    int vz_mode;  // some particle may still carry vz != 0 (constructor-seeded): ordered prediction noise is active
    int tagged_padded;  // the tagged cloud ends in padding entries (x >= 1e29) that are not points (dspmap_estimator.cuh)
};

// Device-resident counters and scalars of one frame (one instance in global memory).
struct DevState {
    // list lengths
    int n_live, n_fov, n_mov, n_mov_owner, mov_top, n_cand, n_cand_owner, cand_top;
    // statistics (SURVEY.md §8d counters)
    int n_left_map, n_voxel_full, n_pyramid_full, n_moved, n_born, n_low_weight, n_pre, n_old, n_out, n_valid;
    int n_inmap_points, n_vdraw, n_rdraw;
    int n_vz, n_skipped;  // prediction-noise draws / particles skipped by the flag test (dsp_dynamic.h:649,653)
    int n_inexact;   // events whose exact serial semantics are not reproduced
    int overflow;    // a device list ran out of capacity
    int work_k4, work_k5;  // dynamic work queues
    int n_occ_voxels;      // occupied-voxel work list of the resampling kernel
    int ticket;            // "last block done" counter of k_arrive's serial replay
    int ticket_prep;       // "last block done" counter of k_pair_prep (maps with many pyramids)
    int n_rel;             // events the replay has to walk (movers + registered stayers of the pyramids that can overflow)
    int gather_max;        // sharded maps: upper bound of any rank's registered particles this frame (sizes the all-gather)
    int n_mov_fov;         // local movers heading for a pyramid (part of that bound)
    int work_eval, work_eval2, work_w2;  // work queues of the pair-buffer kernels
    unsigned long long total_pairs;  // (particle, point) pairs of this frame
    int use_store;         // total_pairs fits the pair buffer: the pair-buffer kernels run, else the recompute kernels
    float norm;      // sum of 1/C_z               (dsp_dynamic.h:799-805)
    float w_new;     // newborn particle weight    (dsp_dynamic.h:805)
    long long p_cur, v_cur, u_cur;  // noise-table cursors and uniform-stream counter (dsp_dynamic.h:483-484)
};

// One dynamic cluster of the device-side velocity estimation front end (dspmap_estimator.cuh), as handed to the host.
struct EstFeature { float cx, cy, cz; int size; int sorted_pos; };
