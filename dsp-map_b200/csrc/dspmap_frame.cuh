// dspmap_frame.cuh — the kernels of one map update, in pipeline order (product code, sm_100a).
#pragma once
#include <cuda_pipeline.h>
#include <cuda/ptx>
#include "dspmap_kernels.cuh"

// ------------------------------------------------------------------------------------------------------------
// (helper) the pair buffer is used when the frame's pairs fit it; otherwise the recompute kernels (k_ck / k_weight) take the frame
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool use_pair_buffer(const MapConst &mc, const DevPtrs &dp) {
    return dp.st->total_pairs <= (unsigned long long)mc.cap_pairs;
}
// ------------------------------------------------------------------------------------------------------------
// K0  frame setup: rotate the boundary-plane normals (dsp_dynamic.h:226-232), reset per-frame state (:235-238)
// ------------------------------------------------------------------------------------------------------------
__global__ void k_frame_setup(MapConst mc, FrameConst fc, DevPtrs dp) {
    pdl_enter();
    int np = mc.Nh + 1 + mc.Nv + 1;
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
        float o[3];
        dsp_rotate(dp.planes0 + 3 * i, fc.q, fc.qi, o);
        dp.planes[3 * i] = o[0];
        dp.planes[3 * i + 1] = o[1];
        dp.planes[3 * i + 2] = o[2];
    }
    if (threadIdx.x == 0) {
        DevState *s = dp.st;
        s->n_live = s->n_fov = s->n_mov = s->n_mov_owner = s->mov_top = 0;
        s->n_cand = s->n_cand_owner = s->cand_top = 0;
        s->n_left_map = s->n_voxel_full = s->n_pyramid_full = s->n_moved = s->n_born = 0;
        s->n_low_weight = s->n_pre = s->n_old = s->n_out = s->n_valid = 0;
        s->n_inmap_points = s->n_vdraw = s->n_rdraw = 0;
        s->n_inexact = 0;
        s->n_vz = s->n_skipped = 0;
        s->n_occ_voxels = 0;
        s->ticket = 0;
        s->ticket_prep = 0;
        s->n_rel = 0;
        s->n_mov_fov = 0;
        s->work_eval = s->work_eval2 = s->work_w2 = 0;
        s->total_pairs = 0ull;
        s->use_store = 0;
        s->work_k4 = s->work_k5 = 0;
        s->norm = 0.f;
        s->w_new = 0.f;
    }
    if (mc.sharded) {
        for (int r = threadIdx.x; r < mc.nranks; r += blockDim.x) reinterpret_cast<int *>(dp.xsend + (size_t)r * (SLAB_HDR + mc.cap_x * XREC))[0] = 0;
        if (threadIdx.x == 0) reinterpret_cast<int *>(dp.gsend)[0] = 0;
    }
    for (int i = threadIdx.x; i < mc.P; i += blockDim.x) {
        dp.obs_cnt[i] = 0;
        dp.obs_fill[i] = 0;
        dp.obs_maxbits[i] = __float_as_int(-1.f);
        dp.pcount[i] = 0;
        dp.pfill[i] = 0;
        dp.pub[i] = 0;
    }
}

// ------------------------------------------------------------------------------------------------------------
// K1  observation binning (dsp_dynamic.h:244-290): rotate, FOV test, pyramid id, count, max range
//     The reference appends points to their pyramid in input order and keeps the first OBS-1; here points are
//     counted, scattered and then ranked by input index inside their pyramid (k_obs_rank).
// ------------------------------------------------------------------------------------------------------------
__global__ void k_obs_classify(MapConst mc, FrameConst fc, DevPtrs dp) {
    pdl_enter();
    __shared__ float sp[3 * DSP_MAX_PLANES];
    load_planes(sp, dp, mc);
    const float *ph = sp, *pv = sp + 3 * (mc.Nh + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < fc.n_points; i += gridDim.x * blockDim.x) {
        float v[3] = {dp.pts[3 * i], dp.pts[3 * i + 1], dp.pts[3 * i + 2]}, r[3];
        dsp_rotate(v, fc.q, fc.qi, r);
        int pid = -1;
        float len = 0.f;
        if (dsp_in_fov(ph, pv, mc.Nh, mc.Nv, r[0], r[1], r[2])) {
            int h = dsp_pyr_scan(ph, mc.Nh, 1.f, r[0], r[1], r[2]);
            int w = dsp_pyr_scan(pv, mc.Nv, -1.f, r[0], r[1], r[2]);
            pid = h * mc.Nv + w;
            len = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
            if (h < 0 || w < 0) pid = -1;  // cannot happen for finite input inside the FOV planes
        }
        dp.OR[i] = make_float4(r[0], r[1], r[2], len);
        dp.OPID[i] = pid;
        if (pid >= 0) {
            atomicAdd(&dp.obs_cnt[pid], 1);
            atomicMax(&dp.obs_maxbits[pid], __float_as_int(len));  // ranges are positive: int order == float order
            agg_inc(&dp.st->n_valid);
        }
    }
}

// single-block exclusive scans of small int arrays (one array per block; blockIdx.x selects the job); optionally a
// second scan of min(x, cap) of the same input
struct ScanJob { const int *in; int *out; int *out_capped; int cap; int n; };
struct ScanJobs { ScanJob j[3]; };
__global__ void __launch_bounds__(1024) k_scan_small(ScanJobs jobs) {
    pdl_enter();
    __shared__ int wsum[32], wcap[32];
    const ScanJob J = jobs.j[blockIdx.x];
    const int n = J.n, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // one pass: every thread owns `per` consecutive elements (the arrays are a few tens of KB and sit in L2 / L1, so the
    // strided access costs nothing next to the block-wide synchronisations a round-per-1024-elements loop pays)
    const int per = (n + 1023) >> 10;
    const int b0 = min(n, threadIdx.x * per), b1 = min(n, b0 + per);
    int s = 0, sc = 0;
    for (int i = b0; i < b1; ++i) { const int x = J.in[i]; s += x; sc += min(x, J.cap); }
    int incl = s, inclc = sc;
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULLMASK, incl, d), tc = __shfl_up_sync(FULLMASK, inclc, d);
        if (lane >= d) { incl += t; inclc += tc; }
    }
    if (lane == 31) { wsum[wid] = incl; wcap[wid] = inclc; }
    __syncthreads();
    if (wid == 0) {
        int y0 = wsum[lane], yc0 = wcap[lane], y = y0, yc = yc0;
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(FULLMASK, y, d), tc = __shfl_up_sync(FULLMASK, yc, d);
            if (lane >= d) { y += t; yc += tc; }
        }
        wsum[lane] = y - y0;  // exclusive prefix of the warp totals
        wcap[lane] = yc - yc0;
    }
    __syncthreads();
    int run = wsum[wid] + incl - s, runc = wcap[wid] + inclc - sc;
    for (int i = b0; i < b1; ++i) {
        const int x = J.in[i];
        J.out[i] = run;
        run += x;
        if (J.out_capped) { J.out_capped[i] = runc; runc += min(x, J.cap); }
    }
    if (threadIdx.x == 1023) { J.out[n] = run; if (J.out_capped) J.out_capped[n] = runc; }
}


// exclusive scan of n ints by one block (any block size that is a multiple of 32); out may be shared or global memory,
// out[n] receives the total; wsum: 32 ints of shared memory
__device__ __forceinline__ void block_exclusive_scan(const int *in, int *out, int n, int *wsum) {
    const int T = blockDim.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = T >> 5;
    const int per = (n + T - 1) / T;
    const int b0 = min(n, (int)threadIdx.x * per), b1 = min(n, b0 + per);
    int s = 0;
    for (int i = b0; i < b1; ++i) s += in[i];
    int incl = s;
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(FULLMASK, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const int y0 = lane < nw ? wsum[lane] : 0;
        int y = y0;
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(FULLMASK, y, d);
            if (lane >= d) y += t;
        }
        wsum[lane] = y - y0;  // exclusive prefix of the warp totals
    }
    __syncthreads();
    int run = wsum[wid] + incl - s;
    for (int i = b0; i < b1; ++i) {
        const int x = in[i];
        out[i] = run;
        run += x;
    }
    if ((int)threadIdx.x == T - 1) out[n] = run;
    __syncthreads();  // the warp totals may be reused by a following scan; out[] is visible to the whole block
}
__global__ void k_obs_scatter(MapConst mc, FrameConst fc, DevPtrs dp) {
    pdl_enter();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < fc.n_points; i += gridDim.x * blockDim.x) {
        int pid = dp.OPID[i];
        if (pid < 0) continue;
        int pos = dp.obs_off[pid] + atomicAdd(&dp.obs_fill[pid], 1);
        dp.OSEG[pos] = i;
    }
}
__global__ void k_obs_rank(MapConst mc, FrameConst fc, DevPtrs dp) {
    pdl_enter();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < fc.n_points; i += gridDim.x * blockDim.x) {
        int pid = dp.OPID[i];
        if (pid < 0) continue;
        int b = dp.obs_off[pid], n = dp.obs_cnt[pid], rank = 0;
        for (int j = 0; j < n; ++j) rank += (dp.OSEG[b + j] < i);
        // the first OBS-1 points in input order are kept (:281-284: the OBS-th slot is overwritten, never read)
        if (rank < mc.OBS - 1) dp.OBSP[pid * mc.OBS + rank] = dp.OR[i];
    }
}

// ------------------------------------------------------------------------------------------------------------
// K2a  enumerate live particles from the occupancy masks; snapshot the frame-start masks
// ------------------------------------------------------------------------------------------------------------
__global__ void k_enumerate(MapConst mc, DevPtrs dp, int snapshot) {
    pdl_enter();
    int lane = threadIdx.x & 31;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int base = warp * 32; base < mc.V; base += nwarps * 32) {
        int v = base + lane;
        ulonglong2 m = make_ulonglong2(0ull, 0ull);
        if (v < mc.V) {
            m = dp.M[v];
            if (snapshot) dp.M0[v] = m;
        }
        int c = mask_popc(m), incl = c;
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(FULLMASK, incl, d);
            if (lane >= d) incl += t;
        }
        int total = __shfl_sync(FULLMASK, incl, 31);
        if (total == 0) continue;
        int wbase = 0;
        if (lane == 0) wbase = atomicAdd(&dp.st->n_live, total);
        wbase = __shfl_sync(FULLMASK, wbase, 0);
        int off = wbase + incl - c;
        // past the capacity the particles are not enumerated (flagged); what fits is written, so E has no unwritten gaps
        if (c && off + c > dp.cap_live) atomicOr(&dp.st->overflow, 1);
        u64 z = m.x;
        while (z) { int s = __ffsll((long long)z) - 1; z &= z - 1; if (off < dp.cap_live) dp.E[off] = (v << DSP_KEY_SHIFT) | s; ++off; }
        z = m.y;
        while (z) { int s = __ffsll((long long)z) - 1; z &= z - 1; if (off < dp.cap_live) dp.E[off] = (v << DSP_KEY_SHIFT) | (64 + s); ++off; }
    }
}

// prediction noise predicate (dsp_dynamic.h:649,653): live, not newborn-flagged, |vx*vy*vz| >= 1e-6
__device__ __forceinline__ bool vz_noisy(float4 B) {
    return (B.w > 0.1f && B.w < 6.f) && !((double)fabsf(B.x * B.y * B.z) < 1e-6);
}
// vz mode only: per-voxel count of particles that will draw prediction noise (then scanned over voxels)
__global__ void k_vz_count(MapConst mc, DevPtrs dp) {
    pdl_enter();
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < mc.V; v += gridDim.x * blockDim.x) {
        ulonglong2 m = dp.M[v], bits = make_ulonglong2(0ull, 0ull);
        for (int half = 0; half < 2; ++half) {
            u64 z = half ? m.y : m.x;
            while (z) {
                int sb = __ffsll((long long)z) - 1;
                z &= z - 1;
                if (vz_noisy(dp.PB[v * mc.S + sb + 64 * half])) { if (half) bits.y |= 1ull << sb; else bits.x |= 1ull << sb; }
            }
        }
        dp.MS[v] = bits;  // MS is free until the arrival grouping; it carries the predicate bits to k_predict
        int c = mask_popc(bits);
        dp.vzcnt[v] = c;
    }
}
#define SCAN_BLOCK 2048
__global__ void __launch_bounds__(256) k_scan_blocksum(const int *in, int n, int *blocksum) {
    pdl_enter();
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    int b = blockIdx.x * SCAN_BLOCK, c = 0;
    for (int i = threadIdx.x; i < SCAN_BLOCK; i += blockDim.x) if (b + i < n) c += in[b + i];
    for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(FULLMASK, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s, c);
    __syncthreads();
    if (threadIdx.x == 0) blocksum[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_scan_apply(const int *in, int n, const int *blockoff, int *out, int nblocks) {
    pdl_enter();
    __shared__ int tsum[256];
    int b = blockIdx.x * SCAN_BLOCK + threadIdx.x * (SCAN_BLOCK / 256);
    int loc[SCAN_BLOCK / 256], c = 0;
    for (int k = 0; k < SCAN_BLOCK / 256; ++k) { loc[k] = (b + k < n) ? in[b + k] : 0; c += loc[k]; }
    tsum[threadIdx.x] = c;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
        int a = threadIdx.x >= d ? tsum[threadIdx.x - d] : 0;
        __syncthreads();
        tsum[threadIdx.x] += a;
        __syncthreads();
    }
    int run = blockoff[blockIdx.x] + tsum[threadIdx.x] - c;
    for (int k = 0; k < SCAN_BLOCK / 256; ++k) { if (b + k < n) out[b + k] = run; run += loc[k]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = blockoff[nblocks];
}
__global__ void k_vz_advance(MapConst mc, DevPtrs dp) {
    pdl_enter();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int n = dp.vzoff[mc.V];
        dp.st->n_vz = n;
        dp.st->v_cur = (dp.st->v_cur + 3ll * n) % mc.G;
    }
}

// ------------------------------------------------------------------------------------------------------------
// K2b  prediction (dsp_dynamic.h:645-694; dsp_static.h:640-646) + reindexing (:669-671) + FOV / pyramid id of
//      the new position (:1233-1243).  Stayers are updated in place; movers leave their slot and go to the mover
//      buffer, to be placed by k_arrive in the reference's sweep order.
//      Velocity process noise (:653-659, :1262-1269) cannot fire here: with LIMIT_MOVEMENT_IN_XY_PLANE = 1 every
//      particle has vz == 0 after its first prediction or at birth, so |vx*vy*vz| < 1e-6 always holds; the one
//      reachable case (constructor-seeded particles in their first prediction) takes the ordered path below
//      (fc.vz_mode: ranks from k_vz_count + scan).
// ------------------------------------------------------------------------------------------------------------
__global__ void k_predict(MapConst mc, FrameConst fc, DevPtrs dp) {
    pdl_enter();
    __shared__ float sp[3 * DSP_MAX_PLANES];
    load_planes(sp, dp, mc);
    const float *ph = sp, *pv = sp + 3 * (mc.Nh + 1);
    const int n = min(dp.st->n_live, dp.cap_live);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int key = dp.E[i];
        int v = key >> DSP_KEY_SHIFT, s = key & (DSP_MAX_SLOTS - 1);
        int a = v * mc.S + s;
        float4 A = dp.PA[a], B = dp.PB[a];
        if (!(B.w > 0.1f && B.w < 6.f)) {  // newborn-flagged (constructor-seeded) particles are not predicted (:649)
            agg_inc(&dp.st->n_skipped);
            continue;
        }
        if (mc.model == 1) {
            B.x = 0.f; B.y = 0.f; B.z = 0.f;
            A.x += fc.sx; A.y += fc.sy; A.z += fc.sz;
        } else {
            // (:653-659) three table draws, cursor in sweep order; the predicate bits were frozen by k_vz_count
            if (fc.vz_mode && mask_test(dp.MS[v], s)) {
                ulonglong2 vzm = dp.MS[v];
                int rank = dp.vzoff[v];
                if (s < 64) rank += __popcll(vzm.x & ((1ull << s) - 1ull));
                else rank += __popcll(vzm.x) + __popcll(vzm.y & ((1ull << (s - 64)) - 1ull));
                long long vc = (dp.st->v_cur + 3ll * rank) % mc.G;
                B.x += dp.vtab[vc];
                B.y += dp.vtab[(vc + 1) % mc.G];
                B.z += dp.vtab[(vc + 2) % mc.G];
            }
            B.z = 0.f;
            A.x += fc.dt * B.x + fc.sx;
            A.y += fc.dt * B.y + fc.sy;
            A.z += fc.dt * B.z + fc.sz;
        }
        int d = dsp_voxel_index(mc, A.x, A.y, A.z);
        if (d < 0) {  // left the map (:686-690)
            mask_atomic_clear(dp.M, v, s);
            agg_inc(&dp.st->n_left_map);
            continue;
        }
        int q = -1;
        if (dsp_in_fov(ph, pv, mc.Nh, mc.Nv, A.x, A.y, A.z)) {
            int h = dsp_pyr_scan(ph, mc.Nh, 1.f, A.x, A.y, A.z);
            int w = dsp_pyr_scan(pv, mc.Nv, -1.f, A.x, A.y, A.z);
            if (h >= 0 && w >= 0) q = h * mc.Nv + w;
        }
        if (d == v) {
            B.w = 1.f;
            dp.PA[a] = A;
            dp.PB[a] = B;
            if (q >= 0) {
                int k = agg_inc(&dp.st->n_fov);
                dp.Fkey[k] = key;
                dp.Faddr[k] = a;
                dp.Fq[k] = q;
                dp.FP[k] = A;
                if (!mc.sharded) atomicAdd(&dp.pcount[q], 1);
            }
        } else {
            mask_atomic_clear(dp.M, v, s);  // "remove from ori voxel first" (:1210)
            B.w = 7.f;
            if (mc.sharded && (d < mc.v_lo || d >= mc.v_hi)) {  // crosses into another rank's voxel subspace
                float *slab = dp.xsend + (size_t)dsp_owner(mc, d) * (SLAB_HDR + mc.cap_x * XREC);
                int k = atomicAdd(reinterpret_cast<int *>(slab), 1);
                if (k >= mc.cap_x) { atomicOr(&dp.st->overflow, 8); continue; }
                float *rec = slab + SLAB_HDR + (size_t)k * XREC;
                *reinterpret_cast<float4 *>(rec) = A;
                *reinterpret_cast<float4 *>(rec + 4) = B;
                reinterpret_cast<int *>(rec)[8] = key;
                reinterpret_cast<int *>(rec)[9] = d;
                reinterpret_cast<int *>(rec)[10] = q;
                continue;
            }
            if (q >= 0 && !mc.sharded) atomicAdd(&dp.pub[q], 1);  // upper bound of the pyramid's list length (arrive_needs_replay)
            if (q >= 0 && mc.sharded) agg_inc(&dp.st->n_mov_fov);
            int k = agg_inc(&dp.st->n_mov);
            dp.MBA[k] = A;
            dp.MBB[k] = B;
            dp.MBkey[k] = key;
            dp.MBdst[k] = d;
            dp.MBq[k] = q;
            if (atomicAdd(&dp.mcnt[d], 1) == 0) dp.mowner[agg_inc(&dp.st->n_mov_owner)] = d;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// arrival grouping: arrivals to the same destination voxel are gathered into one contiguous segment so that each
// arrival can rank itself among them.  owner pass: one thread per distinct destination allocates the segment and
// snapshots the destination's mask; scatter pass: arrivals write their order key into the segment.
// ------------------------------------------------------------------------------------------------------------
__global__ void k_group_owner(DevPtrs dp, const int *n_owner, const int *owner, const int *cnt, int *base, int *top) {
    pdl_enter();
    int lane = threadIdx.x & 31;
    int n = *n_owner;
    int nround = (n + 31) & ~31;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += gridDim.x * blockDim.x) {
        int d = i < n ? owner[i] : -1;
        int c = d >= 0 ? cnt[d] : 0, incl = c;
        for (int k = 1; k < 32; k <<= 1) {
            int t = __shfl_up_sync(FULLMASK, incl, k);
            if (lane >= k) incl += t;
        }
        int total = __shfl_sync(FULLMASK, incl, 31), wb = 0;
        if (lane == 0) wb = atomicAdd(top, total);
        wb = __shfl_sync(FULLMASK, wb, 0);
        if (d >= 0) {
            base[d] = wb + incl - c;
            dp.MS[d] = dp.M[d];
        }
    }
}
__global__ void k_group_scatter(const int *n_items, const int *dst, const int *key, const int *base, int *fill, int *seg, int *segi) {
    pdl_enter();
    int n = *n_items;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int d = dst[i];
        const int pos = base[d] + atomicAdd(&fill[d], 1);
        seg[pos] = key[i];
        if (segi) segi[pos] = i;
    }
}

// ------------------------------------------------------------------------------------------------------------
// K3  moveParticle's voxel half (dsp_dynamic.h:1206-1230), replayed in the reference's sweep order:
//     arrivals from lower-index voxels come before the destination's own particles are processed and see the
//     frame-start mask M0; the destination's own leavers then free their slots; arrivals from higher-index voxels
//     come last.  The k-th arrival of a phase takes the k-th free slot; no free slot => the particle vanishes.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void arrive_body(const MapConst &mc, const FrameConst &fc, const DevPtrs &dp) {
    const int n = dp.st->n_mov;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int d = dp.MBdst[i], key = dp.MBkey[i];
        int b = dp.mbase[d], c = dp.mcnt[d];
        int dkey = d << DSP_KEY_SHIFT;
        bool early = key < dkey;
        int n_early = 0, rank = 0;
        for (int j = 0; j < c; ++j) {
            int kj = dp.mseg[b + j];
            bool ej = kj < dkey;
            n_early += ej;
            rank += (ej == early) && (kj < key);
        }
        ulonglong2 m0 = dp.M0[d], mid = dp.MS[d];  // mid = M0 minus the destination's own leavers
        int slot;
        if (early) {
            slot = mask_nth_free(mc, m0, rank);
        } else {
            ulonglong2 t = mask_take_free(mc, m0, n_early);
            ulonglong2 m2 = make_ulonglong2(mid.x | t.x, mid.y | t.y);
            slot = mask_nth_free(mc, m2, rank);
        }
        if (slot < 0) {  // voxel full: the particle vanishes (:1227-1229)
            agg_inc(&dp.st->n_voxel_full);
            continue;
        }
        int a = d * mc.S + slot;
        dp.PA[a] = dp.MBA[i];
        dp.PB[a] = dp.MBB[i];
        mask_atomic_set(dp.M, d, slot);
        agg_inc(&dp.st->n_moved);
        int q = dp.MBq[i];
        if (q >= 0) {
            int k = agg_inc(&dp.st->n_fov);
            dp.Fkey[k] = key;  // list order is the order of processing, i.e. the SOURCE sweep key
            dp.Faddr[k] = a;
            dp.Fq[k] = q;
            dp.FP[k] = dp.MBA[i];
            if (!mc.sharded) atomicAdd(&dp.pcount[q], 1);
        }
    }
}
// ------------------------------------------------------------------------------------------------------------
// Pyramid-list overflow (dsp_dynamic.h:1243-1259).  When a pyramid's list is full the particle being processed vanishes and
// frees its voxel slot AT THAT MOMENT of the sweep, so a later arrival of the same frame may take the slot: slot assignment
// and list membership then depend on each other along the whole sweep, and the parallel ranking above no longer applies.
// Lists overflow only when the reference's size formula yields a tiny L (a handful of entries for small maps at 1 degree);
// no BASELINE configuration comes near.  k_arrive therefore first bounds every list from above — particles that stayed in
// their voxel inside the field of view (counted by k_predict) plus ALL movers heading for the pyramid — and only if some
// bound exceeds L it replays moveParticle serially, in sweep order, exactly as the reference executes it:
//   phase A (whole grid): rank the frame's events (registered stayers and movers; keys are unique) by sweep key;
//   phase B (the block that finishes last, one thread): walk them in order with working masks W (= MS, seeded from M0):
//     arrivals from lower-index voxels see the destination before its own sweep, later ones after its own leavers
//     (M0 & ~M, removed lazily when the sweep passes the voxel) and its own vanished stayers freed their slots.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool arrive_needs_replay(const MapConst &mc, const DevPtrs &dp) {
    int over = 0;
    if (!mc.sharded)
        for (int q = threadIdx.x; q < mc.P; q += blockDim.x) over |= dp.pcount[q] + dp.pub[q] > mc.L;
    return __syncthreads_or(over) != 0;
}
__device__ __forceinline__ void replay_apply_leavers(const DevPtrs &dp, int v) {  // the sweep has passed voxel v
    if (dp.vzcnt[v]) return;
    dp.vzcnt[v] = 1;
    const ulonglong2 m0 = dp.M0[v], m = dp.M[v];  // M = M0 minus everything that left v (k_predict)
    ulonglong2 w = dp.MS[v];
    w.x &= ~(m0.x & ~m.x);
    w.y &= ~(m0.y & ~m.y);
    dp.MS[v] = w;
}
__device__ __forceinline__ void replay_clear(const DevPtrs &dp, int v, int s) {
    ulonglong2 w = dp.MS[v];
    if (s < 64) w.x &= ~(1ull << s); else w.y &= ~(1ull << (s - 64));
    dp.MS[v] = w;
}
__device__ __forceinline__ void replay_touch(const void *p) {  // bring the line into this SM's L1 (the walk below is one thread: latency is everything)
#ifdef __CUDA_ARCH__
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}
__device__ void arrive_replay(const MapConst &mc, const DevPtrs &dp) {
    __shared__ bool s_last;
    __shared__ int s_pos;  // how far the walk has come (a hint for the prefetching warps: exchanged with shared-memory atomics)
    DevState *st = dp.st;
    const int n_stay = st->n_fov, n_mov = st->n_mov, n_ev = n_stay + n_mov;  // (nobody changes them before phase B)
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    int *ev = dp.PSkey, *evl = dp.PSaddr, *evk = dp.rkey;  // free until k_pyr_scatter
    // phase A (whole grid): working masks; the events the walk has to see.  A registered stayer only matters if ITS pyramid can
    // overflow (k_predict counted it into pcount already, and a list that cannot fill up never turns anyone away): the walk
    // is as long as the movers plus the stayers of the few pyramids over the bound, not as long as the live list.
    for (int v = tid; v < mc.V; v += nth) { dp.MS[v] = dp.M0[v]; dp.vzcnt[v] = 0; }
    for (int e = tid; e < n_ev; e += nth) {
        int key;
        if (e < n_stay) {
            const int q = dp.Fq[e];
            if (!(dp.pcount[q] + dp.pub[q] > mc.L)) continue;
            key = dp.Fkey[e];
        } else {
            key = dp.MBkey[e - n_stay];
        }
        const int k = atomicAdd(&st->n_rel, 1);
        evl[k] = e;
        evk[k] = key;
        dp.rpos[k] = atomicAdd(&dp.rcount[key >> DSP_KEY_SHIFT], 1);  // its place among the events of its voxel (any order; sorted below)
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&st->ticket, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // phase B (the block that finishes last): order the events by sweep key — a counting sort by voxel (the key's upper bits;
    // rcount holds the events per voxel), then the handful of events of a voxel by slot — and walk them.  (Ranking the events
    // by counting smaller keys, 20 k x 20 k comparisons on one SM, took 3.6 ms of the 15 ms a cfg3 replay frame cost.)
    const int n_rel = *(volatile int *)&st->n_rel;
    {
        __shared__ int wsum[32];
        int *skey = dp.rkey2;
        block_exclusive_scan(dp.rcount, dp.rcount, mc.V, wsum);  // in place: every thread rewrites only the stretch it has just summed
        for (int k = threadIdx.x; k < n_rel; k += blockDim.x) {
            const int key = __ldcg(evk + k);
            const int at = dp.rcount[key >> DSP_KEY_SHIFT] + __ldcg(dp.rpos + k);
            ev[at] = __ldcg(evl + k);
            skey[at] = key;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < n_rel; k += blockDim.x) {
            if (__ldcg(dp.rpos + k) != 0) continue;  // one thread per voxel that has events: insertion sort of its stretch by key
            const int v = __ldcg(evk + k) >> DSP_KEY_SHIFT;
            const int lo = dp.rcount[v], hi = dp.rcount[v + 1];
            for (int i = lo + 1; i < hi; ++i) {
                const int ki = skey[i], ei = ev[i];
                int j = i - 1;
                while (j >= lo && skey[j] > ki) { skey[j + 1] = skey[j]; ev[j + 1] = ev[j]; --j; }
                skey[j + 1] = ki;
                ev[j + 1] = ei;
            }
        }
        __syncthreads();
        for (int v = threadIdx.x; v <= mc.V; v += blockDim.x) dp.rcount[v] = 0;  // (the scan left offsets everywhere) zero between uses
    }
    for (int q = threadIdx.x; q < mc.P; q += blockDim.x)
        if (dp.pcount[q] + dp.pub[q] > mc.L) dp.pcount[q] = 0;  // recounted by the walk; the other lists keep k_predict's count of their stayers
    if (threadIdx.x == 0) s_pos = 0;
    __syncthreads();
    if (threadIdx.x >= 32) {
        // the other warps run a little ahead of the walk and pull what it is going to read into L1
        const int ahead = blockDim.x - 32;
        for (int r = threadIdx.x - 32; r < n_rel; r += ahead) {
            while (r > atomicOr(&s_pos, 0) + 96) {  // (close ahead: an event touches ~1 KB of lines, and they have to still be in L1 when the walk arrives)
#ifdef __CUDA_ARCH__
                __nanosleep(100);
#endif
            }
            const int e = ev[r];
            int v;
            if (e < n_stay) {
                replay_touch(dp.Fq + e);
                v = dp.Fkey[e] >> DSP_KEY_SHIFT;
            } else {
                const int i = e - n_stay;
                replay_touch(dp.MBkey + i); replay_touch(dp.MBq + i); replay_touch(dp.MBA + i); replay_touch(dp.MBB + i);
                v = dp.MBdst[i];
            }
            replay_touch(dp.MS + v); replay_touch(dp.M0 + v); replay_touch(dp.M + v); replay_touch(dp.vzcnt + v);
        }
    }
    if (threadIdx.x == 0) {
        int n_fov = n_stay, n_moved = 0, n_vfull = 0, n_pfull = 0;
        // the walk reads one event ahead: the next event's record is requested before the current one is applied
        int e_nx = n_rel > 0 ? ev[0] : 0, key_nx = 0, q_nx = 0, d_nx = 0;
        if (n_rel > 0) {
            if (e_nx < n_stay) { key_nx = dp.Fkey[e_nx]; q_nx = dp.Fq[e_nx]; }
            else { key_nx = dp.MBkey[e_nx - n_stay]; d_nx = dp.MBdst[e_nx - n_stay]; q_nx = dp.MBq[e_nx - n_stay]; }
        }
        for (int r = 0; r < n_rel; ++r) {
            if ((r & 15) == 0) atomicExch(&s_pos, r);
            const int e = e_nx, key = key_nx, q = q_nx, d = d_nx;
            if (r + 1 < n_rel) {
                e_nx = ev[r + 1];
                if (e_nx < n_stay) { key_nx = dp.Fkey[e_nx]; q_nx = dp.Fq[e_nx]; }
                else { key_nx = dp.MBkey[e_nx - n_stay]; d_nx = dp.MBdst[e_nx - n_stay]; q_nx = dp.MBq[e_nx - n_stay]; }
            }
            if (e < n_stay) {  // stayed in its voxel, inside the field of view: joins its pyramid's list unless that is full
                if (dp.pcount[q] < mc.L) {
                    ++dp.pcount[q];
                } else {  // vanishes (:1256-1259)
                    const int v = key >> DSP_KEY_SHIFT;
                    replay_apply_leavers(dp, v);
                    replay_clear(dp, v, key & (DSP_MAX_SLOTS - 1));
                    dp.Fq[e] = -1;
                    ++n_pfull;
                }
                continue;
            }
            const int i = e - n_stay;
            if ((key >> DSP_KEY_SHIFT) > d) replay_apply_leavers(dp, d);
            ulonglong2 w = dp.MS[d];
            const int slot = mask_nth_free(mc, w, 0);
            if (slot < 0) { ++n_vfull; continue; }  // voxel full: the particle vanishes (:1227-1229)
            const int a = d * mc.S + slot;
            dp.PA[a] = dp.MBA[i];
            dp.PB[a] = dp.MBB[i];
            ++n_moved;
            if (q >= 0 && dp.pcount[q] >= mc.L) { ++n_pfull; continue; }  // pyramid full: the slot is free again at once
            if (slot < 64) w.x |= 1ull << slot; else w.y |= 1ull << (slot - 64);
            dp.MS[d] = w;
            if (q >= 0) {
                ++dp.pcount[q];
                dp.Fkey[n_fov] = key;
                dp.Faddr[n_fov] = a;
                dp.Fq[n_fov] = q;
                dp.FP[n_fov] = dp.MBA[i];
                ++n_fov;
            }
        }
        atomicExch(&s_pos, n_rel);
        st->n_fov = n_fov;
        st->n_moved = n_moved;
        st->n_voxel_full = n_vfull;
        st->n_pyramid_full = n_pfull;
    }
    __syncthreads();
    // final masks: the working masks, with the leavers of the voxels the walk never passed removed as well
    for (int v = threadIdx.x; v < mc.V; v += blockDim.x) {
        const ulonglong2 m0 = dp.M0[v], m = dp.M[v];
        ulonglong2 w = dp.MS[v];
        if (!dp.vzcnt[v]) { w.x &= ~(m0.x & ~m.x); w.y &= ~(m0.y & ~m.y); }
        dp.M[v] = w;
    }
}
__global__ void k_arrive(MapConst mc, FrameConst fc, DevPtrs dp) {
    pdl_enter();
    if (arrive_needs_replay(mc, dp)) arrive_replay(mc, dp);
    else arrive_body(mc, fc, dp);
}

// ------------------------------------------------------------------------------------------------------------
// Sharded mode (multi-GPU, voxel subspaces = z slabs).  After the all-to-all, movers that crossed into this rank's slab
// join the local mover list; their sweep keys are global, so k_arrive replays them in the reference's order.  After the
// arrival pass every rank contributes its registered particles to an all-gather, and all ranks build identical global
// pyramid lists (slot addresses are only meaningful on the owner; -1 elsewhere).
// ------------------------------------------------------------------------------------------------------------
__global__ void k_shard_import(MapConst mc, DevPtrs dp) {
    pdl_enter();
    const int slab = SLAB_HDR + mc.cap_x * XREC;
    const int total = mc.nranks * mc.cap_x;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int src = t / mc.cap_x, j = t - src * mc.cap_x;
        if (src == mc.rank) continue;
        const float *sl = dp.xrecv + (size_t)src * slab;
        if (j >= min(reinterpret_cast<const int *>(sl)[0], mc.cap_x)) continue;
        const float *rec = sl + SLAB_HDR + (size_t)j * XREC;
        const int d = reinterpret_cast<const int *>(rec)[9];
        int k = agg_inc(&dp.st->n_mov);
        dp.MBA[k] = *reinterpret_cast<const float4 *>(rec);
        dp.MBB[k] = *reinterpret_cast<const float4 *>(rec + 4);
        dp.MBkey[k] = reinterpret_cast<const int *>(rec)[8];
        dp.MBdst[k] = d;
        dp.MBq[k] = reinterpret_cast<const int *>(rec)[10];
        if (atomicAdd(&dp.mcnt[d], 1) == 0) dp.mowner[agg_inc(&dp.st->n_mov_owner)] = d;
    }
}
// end of phase 0: every exchange slab's header carries, besides its own count of boundary crossers, what the receiver needs
// to bound ANY rank's number of registered particles without another exchange: this rank's own candidates (particles that
// stayed in the field of view + local movers heading for a pyramid) and the crossers it sends in total
__global__ void k_shard_headers(MapConst mc, DevPtrs dp) {
    pdl_enter();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int slab = SLAB_HDR + mc.cap_x * XREC;
        int total = 0;
        for (int r = 0; r < mc.nranks; ++r) total += min(reinterpret_cast<int *>(dp.xsend + (size_t)r * slab)[0], mc.cap_x);
        for (int r = 0; r < mc.nranks; ++r) {
            int *h = reinterpret_cast<int *>(dp.xsend + (size_t)r * slab);
            h[1] = dp.st->n_fov + dp.st->n_mov_fov;
            h[2] = total;
        }
    }
}
// after the all-to-all: max over the ranks of (own candidates) + all crossers of the frame >= any rank's registered particles;
// the same number on every rank (every rank holds every sender's header)
__global__ void k_shard_bound(MapConst mc, DevPtrs dp, int *out) {
    pdl_enter();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int slab = SLAB_HDR + mc.cap_x * XREC;
        int own = 0, cross = 0;
        for (int r = 0; r < mc.nranks; ++r) {
            const int *h = reinterpret_cast<const int *>((r == mc.rank ? dp.xsend : dp.xrecv) + (size_t)r * slab);
            own = max(own, h[1]);
            cross += h[2];
        }
        dp.st->gather_max = own + cross;
        *out = own + cross;
    }
}
__global__ void k_shard_pack_fov(MapConst mc, DevPtrs dp) {
    pdl_enter();
    const int n = dp.st->n_fov;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        reinterpret_cast<int *>(dp.gsend)[0] = min(n, mc.cap_g);
        if (n > mc.cap_g) atomicOr(&dp.st->overflow, 16);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < min(n, mc.cap_g); i += gridDim.x * blockDim.x) {
        float *rec = dp.gsend + SLAB_HDR + (size_t)i * GREC;
        reinterpret_cast<int *>(rec)[0] = dp.Fkey[i];
        reinterpret_cast<int *>(rec)[1] = dp.Fq[i];
        reinterpret_cast<int *>(rec)[2] = dp.Faddr[i];
        *reinterpret_cast<float4 *>(rec + 4) = dp.FP[i];
    }
}
// pass 0 counts per pyramid, pass 1 scatters (after the scan of the counts)
__global__ void k_shard_fov_gathered(MapConst mc, DevPtrs dp, int pass) {
    pdl_enter();
    const int slab = SLAB_HDR + mc.cap_g * GREC;
    const int total = mc.nranks * mc.cap_g;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int src = t / mc.cap_g, j = t - src * mc.cap_g;
        const float *sl = dp.grecv + (size_t)src * slab;
        if (j >= reinterpret_cast<const int *>(sl)[0]) continue;
        const float *rec = sl + SLAB_HDR + (size_t)j * GREC;
        const int q = reinterpret_cast<const int *>(rec)[1];
        if (pass == 0) {
            atomicAdd(&dp.pcount[q], 1);
        } else {
            const int pos = dp.poff[q] + atomicAdd(&dp.pfill[q], 1);
            dp.PSkey[pos] = reinterpret_cast<const int *>(rec)[0];
            dp.PSaddr[pos] = src == mc.rank ? reinterpret_cast<const int *>(rec)[2] : -1;
            dp.PSpay[pos] = *reinterpret_cast<const float4 *>(rec + 4);
        }
    }
}

// zero the buffers that are merged with all-reduce(sum): exactly one rank writes each element
__global__ void k_shard_zero(MapConst mc, FrameConst fc, DevPtrs dp, int which) {
    pdl_enter();
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    if (which == 0) {  // C_z [P][OBS] and, right behind it, 1/C_z in bin order
        const size_t n = (size_t)mc.P * mc.OBS + fc.n_points;
        for (size_t i = tid; i < n; i += nth) dp.CZ[i] = 0.f;
    } else if (which == 1) {  // new weights by global list index
        const size_t n = (size_t)dp.poff[mc.P];
        for (size_t i = tid; i < n; i += nth) dp.NW[i] = 0.f;
    } else {                  // newborn split per tagged point
        for (size_t i = tid; i < (size_t)fc.n_tagged; i += nth) dp.nst_shared[i] = 0.f;
    }
}
__global__ void k_shard_apply_weights(MapConst mc, DevPtrs dp) {
    pdl_enter();
    const int n = dp.poff[mc.P];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int a = dp.LA[i];
        if (a >= 0) dp.PA[a].w = dp.NW[i];
    }
}
// ------------------------------------------------------------------------------------------------------------
// K3b  pyramid lists (dsp_dynamic.h:1243-1259): scatter registered particles to their pyramid's segment, sort each
//      segment by sweep key, keep the first L (the rest vanish, :1256-1259), and materialise a compact per-pyramid
//      copy of (px, py, pz, w) for the two observation passes.
// ------------------------------------------------------------------------------------------------------------
// fused != 0: every block first scans the pyramid counts into shared memory itself (P + 1 ints; block 0 also publishes the
// offsets for the kernels that follow) instead of waiting for a scan launch of its own on the critical path
__global__ void __launch_bounds__(256) k_pyr_scatter(MapConst mc, DevPtrs dp, int fused) {
    pdl_enter();
    extern __shared__ int s_poff[];
    __shared__ int wsum[32];
    const int *poff = dp.poff;
    if (fused) {
        block_exclusive_scan(dp.pcount, s_poff, mc.P, wsum);
        if (blockIdx.x == 0)
            for (int q = threadIdx.x; q <= mc.P; q += blockDim.x) dp.poff[q] = s_poff[q];
        poff = s_poff;
    }
    const int n = dp.st->n_fov;
    // neighbouring particles mostly fall in the same pyramid: lanes that share one are counted together and their
    // leader reserves the block of slots with a single atomic (the order inside a segment is fixed by k_pyr_sort)
    const int lane = threadIdx.x & 31;
    const int stride = gridDim.x * blockDim.x;
    for (int i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < n; i0 += stride) {
        const int i = i0 + lane;
        const int q = i < n ? dp.Fq[i] : -1;
        const unsigned peers = __match_any_sync(FULLMASK, q);
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (lane == leader && q >= 0) base = atomicAdd(&dp.pfill[q], __popc(peers));
        base = __shfl_sync(FULLMASK, base, leader);
        if (q >= 0) {
            const int pos = poff[q] + base + __popc(peers & ((1u << lane) - 1u));
            dp.PSkey[pos] = dp.Fkey[i];
            dp.PSaddr[pos] = dp.Faddr[i];
        }
    }
}

#define PYR_SORT_CAP 8192
#ifndef PYR_RANK_MAX
#define PYR_RANK_MAX 1024  // segments up to this length are ranked by counting instead of sorted
#endif
// Bitonic sort in shared memory, one thread per compare-exchange pair: the stages that exchange keys less than 32 apart (40 of
// the 55 stages of a 1024-key network) stay inside one warp's 64-element block and need a warp barrier only.
// (Measured on B200, cfg2: 27.3 -> 20.2 us against one thread per key with block-wide barriers; profiles/r02_ab_switches.jsonl.)
__global__ void __launch_bounds__(512) k_pyr_sort(MapConst mc, DevPtrs dp, float Pd) {
    pdl_enter();
    extern __shared__ u64 skey[];  // PYR_SORT_CAP entries: (sweep key << 32) | slot address
    for (int q = blockIdx.x; q < mc.P; q += gridDim.x) {
        const int n = dp.pcount[q], b = dp.poff[q];
        const int keep = min(n, mc.L);
        if (threadIdx.x == 0) dp.plen[q] = keep;
        if (n == 0) continue;
        if (n <= PYR_RANK_MAX) {
            // short segment (every segment of the BASELINE configurations): no sorting network — each thread counts the keys
            // below its own (keys are unique; the segment sits in shared memory and is read by broadcast) and writes its
            // particle straight to that rank.  One barrier instead of ~45; the gather of the particle is issued before the count.
            int *ikey = reinterpret_cast<int *>(skey);
            for (int i = threadIdx.x; i < n; i += blockDim.x) ikey[i] = dp.PSkey[b + i];
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int a = dp.PSaddr[b + i];
                const float4 pa = mc.sharded ? dp.PSpay[b + i] : dp.PA[a];
                const int ki = ikey[i];
                int r = 0;
#pragma unroll 8
                for (int j = 0; j < n; ++j) r += ikey[j] < ki;
                if (r < keep) {
                    dp.LA[b + r] = a;
                    dp.LP[b + r] = pa;
                    dp.PW[b + r] = Pd * pa.w;
                } else if (a >= 0) {  // pyramid full (sharded maps only, see below)
                    mask_atomic_clear(dp.M, a / mc.S, a % mc.S);
                    atomicAdd(&dp.st->n_pyramid_full, 1);
                    atomicOr(&dp.st->overflow, 32);
                }
            }
            __syncthreads();
        } else if (n <= PYR_SORT_CAP) {
            int m = 1;
            while (m < n) m <<= 1;
            for (int i = threadIdx.x; i < m; i += blockDim.x)
                skey[i] = i < n ? ((u64)(unsigned)dp.PSkey[b + i] << 32) | (unsigned)(mc.sharded ? i : dp.PSaddr[b + i]) : ~0ull;
            __syncthreads();
            // one thread per compare-exchange PAIR; a warp's 32 consecutive pairs of a stage with j <= 16 lie inside one
            // 64-element block that no other warp touches in that stage, so those stages need a warp barrier only
            for (int k = 2; k <= m; k <<= 1) {
                if (k >= 64) __syncthreads();  // the previous round ended in warp-local stages: publish them block-wide
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
                        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), p = i | j;
                        const u64 x = skey[i], y = skey[p];
                        const bool up = (i & k) == 0;
                        if ((x > y) == up) { skey[i] = y; skey[p] = x; }
                    }
                    if (j >= 32) __syncthreads(); else __syncwarp();
                }
            }
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int a = (int)(unsigned)(skey[i] & 0xffffffffull);
                float4 pa;
                if (mc.sharded) {  // the low word is the position inside the segment: address and payload come from there
                    pa = dp.PSpay[b + a];
                    a = dp.PSaddr[b + a];
                } else {
                    pa = dp.PA[a];
                }
                if (i < keep) {
                    dp.LA[b + i] = a;
                    dp.LP[b + i] = pa;
                    dp.PW[b + i] = Pd * pa.w;
                } else if (a >= 0) {  // pyramid full: the particle vanishes and frees its voxel slot (:1256-1259)
                    // single GPU: unreachable, k_arrive replays such frames serially.  Sharded maps build their lists after the
                    // gather, so the slot is freed after all arrivals were placed: flagged (code 32), see DESIGN.md
                    mask_atomic_clear(dp.M, a / mc.S, a % mc.S);
                    atomicAdd(&dp.st->n_pyramid_full, 1);
                    atomicOr(&dp.st->overflow, 32);
                }
            }
            __syncthreads();
        } else {  // oversized segment: rank by counting straight from global memory (slow, rare)
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int ki = dp.PSkey[b + i], r = 0;
                for (int j = 0; j < n; ++j) r += dp.PSkey[b + j] < ki;
                int a = dp.PSaddr[b + i];
                if (r < keep) {
                    const float4 pa = mc.sharded ? dp.PSpay[b + i] : dp.PA[a];
                    dp.LA[b + r] = a;
                    dp.LP[b + r] = pa;
                    dp.PW[b + r] = Pd * pa.w;
                } else if (a >= 0) {
                    mask_atomic_clear(dp.M, a / mc.S, a % mc.S);
                    atomicAdd(&dp.st->n_pyramid_full, 1);
                    atomicOr(&dp.st->overflow, 32);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// K4  C_z pass (dsp_dynamic.h:709-739).  For every observation point z of pyramid i:
//         C_z = sum over neighbour pyramids n (table order), particles of n (list order) of  P_d * w * g(p; z)
//               + (expected_new_born + kappa)
//     as ONE fp32 chain in exactly that order.  A CTA owns a pyramid: all threads evaluate a tile of
//     (particle, point) terms into shared memory, then one thread per point adds the tile's terms to its running
//     sum in list order.
// ------------------------------------------------------------------------------------------------------------
#define K4_THREADS 256
#define K4_TERMS 8192  // term tile capacity (floats)
__global__ void __launch_bounds__(K4_THREADS) k_ck(MapConst mc, FrameConst fc, DevPtrs dp) {
    pdl_enter();
    extern __shared__ float smem[];
    float *lut = smem;                          // DSP_LUT_HALF
    float *term = lut + DSP_LUT_HALF + 3;       // K4_TERMS
    float4 *ptl = (float4 *)(term + K4_TERMS);  // particle tile, <= 256
    float4 *zs = ptl + 256;                     // points of this pyramid, <= OBS
    __shared__ int s_item;
    if (use_pair_buffer(mc, dp)) return;  // the pair-buffer kernels handle this frame
    for (int i = threadIdx.x; i < DSP_LUT_HALF; i += blockDim.x) lut[i] = dp.lut[i];
    const float enb = fc.nb_weight * (float)dp.st->n_valid * (float)fc.nb_num;  // :292
    const float add_k = enb + fc.kappa;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(&dp.st->work_k4, 1);
        __syncthreads();
        const int i = s_item;
        if (i >= mc.P) break;
        const int npts = min(dp.obs_cnt[i], mc.OBS - 1);
        if (npts == 0) continue;
        for (int t = threadIdx.x; t < npts; t += blockDim.x) zs[t] = dp.OBSP[i * mc.OBS + t];
        const int tp = min(256, K4_TERMS / npts);
        float acc = 0.f;
        const unsigned magic = npts > 1 ? 0xffffffffu / (unsigned)npts + 1u : 0u;  // exact t / npts for t < 65536
        const int nn = dp.nbr[i * mc.NBW];
        for (int ns = 0; ns < nn; ++ns) {
            const int pc = dp.nbr[i * mc.NBW + 1 + ns];
            const int lb = dp.poff[pc], ln = dp.plen[pc];
            for (int t0 = 0; t0 < ln; t0 += tp) {
                const int cur = min(tp, ln - t0);
                __syncthreads();
                for (int t = threadIdx.x; t < cur; t += blockDim.x) ptl[t] = dp.LP[lb + t0 + t];
                __syncthreads();
                const int pairs = cur * npts;
                for (int t = threadIdx.x; t < pairs; t += blockDim.x) {
                    int k = npts > 1 ? (int)__umulhi((unsigned)t, magic) : t;
                    int z = t - k * npts;
                    float4 p = ptl[k], o = zs[z];
                    float gk = dsp_pdf(lut, p.x, o.x, fc.sigma) * dsp_pdf(lut, p.y, o.y, fc.sigma) * dsp_pdf(lut, p.z, o.z, fc.sigma);
                    term[t] = fc.Pd * p.w * gk;
                }
                __syncthreads();
                if (threadIdx.x < npts)
                    for (int k = 0; k < cur; ++k) acc += term[k * npts + threadIdx.x];
            }
        }
        if (threadIdx.x < npts) {
            acc += add_k;
            dp.CZ[i * mc.OBS + threadIdx.x] = acc;
            dp.INV[dp.obs_capoff[i] + threadIdx.x] = 1.f / acc;  // for the newborn normaliser (:802)
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// K5  weight pass (dsp_dynamic.h:743-790).  One thread per registered particle; the observation points of the
//     particle's neighbour pyramids are staged in shared memory in (table order, bin order) and summed in that
//     order:  w *= (1 - P_d) + sum_z P_d * g(p; z) / C_z.
// ------------------------------------------------------------------------------------------------------------
#define K5_THREADS 256
__global__ void __launch_bounds__(K5_THREADS) k_weight(MapConst mc, FrameConst fc, DevPtrs dp, int chunks_per_pyr) {
    pdl_enter();
    extern __shared__ float smem[];
    float *lut = smem;
    float4 *zs = (float4 *)(lut + DSP_LUT_HALF + 3);  // NB * (OBS-1) staged points: x y z C_z
    __shared__ int s_item;
    if (use_pair_buffer(mc, dp)) return;  // the pair-buffer kernels handle this frame
    for (int i = threadIdx.x; i < DSP_LUT_HALF; i += blockDim.x) lut[i] = dp.lut[i];
    int staged = -1, nz = 0;
    const int items = mc.P * chunks_per_pyr;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(&dp.st->work_k5, 1);
        __syncthreads();
        const int it = s_item;
        if (it >= items) break;
        const int i = it / chunks_per_pyr, c = it - i * chunks_per_pyr;
        const int ln = dp.plen[i];
        if (c * K5_THREADS >= ln) continue;
        if (staged != i) {
            __syncthreads();
            nz = 0;
            const int nn = dp.nbr[i * mc.NBW];
            for (int ns = 0; ns < nn; ++ns) {
                const int ni = dp.nbr[i * mc.NBW + 1 + ns];
                const int cnt = min(dp.obs_cnt[ni], mc.OBS - 1);
                for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
                    float4 o = dp.OBSP[ni * mc.OBS + t];
                    o.w = dp.CZ[ni * mc.OBS + t];
                    zs[nz + t] = o;
                }
                nz += cnt;
            }
            staged = i;
            __syncthreads();
        }
        const int j = c * K5_THREADS + threadIdx.x;
        if (j >= ln) continue;
        const int lb = dp.poff[i];
        float4 p = dp.LP[lb + j];
        float dist = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
        float maxlen = __int_as_float(dp.obs_maxbits[i]);
        if (maxlen > 0.f && dist > maxlen + mc.occl) continue;  // occluded (:761)
        float sum = 0.f;
        for (int z = 0; z < nz; ++z) {
            float4 o = zs[z];
            float gk = dsp_pdf(lut, p.x, o.x, fc.sigma) * dsp_pdf(lut, p.y, o.y, fc.sigma) * dsp_pdf(lut, p.z, o.z, fc.sigma);
            sum += fc.Pd * gk / o.w;
        }
        int a = dp.LA[lb + j];
        if (a >= 0) dp.PA[a].w = p.w * (fc.one_minus_Pd + sum);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Pair-buffer observation passes.  Every (particle, point) pair with the point's pyramid among the particle's pyramid's
// neighbours is evaluated ONCE (k_pair_eval) into G; the C_z pass (k_cz_chain) and the weight pass (k_weight2) then add
// their terms in exactly the reference's order (dsp_dynamic.h:709-739, 743-790).
//   layout  G[rowbase[i] + z * totlen[i] + j],  i = pyramid of the point, z = its bin index, j = position of the
//           particle in the concatenation of i's neighbour lists (neighbour-table order, list order)
//   k_cz_chain reads a row sequentially (j); k_weight2 reads it with consecutive lanes = consecutive particles.
// ------------------------------------------------------------------------------------------------------------
// One block: per point pyramid i the offsets of its neighbours' lists in the concatenation (cum), the block's size in G and
// the number of 32-particle chunks of its own list; then the two exclusive scans (rowbase, chunk_off) and the chunk -> pyramid
// table the pair-buffer kernels index with their queue tickets.  (Round 1 ran this as k_pair_prep + a scan launch and let
// every work item find its pyramid by binary search over chunk_off.)
// (With many pyramids — cfg3: 21 600 at 1 degree, 25 neighbours each — the per-pyramid part is spread over several blocks and
// the block that finishes last does the scans: 0.11 ms of a 0.76 ms frame as a single block.)
__global__ void __launch_bounds__(1024) k_pair_prep(MapConst mc, DevPtrs dp) {
    pdl_enter();
    __shared__ int wsum[32];
    __shared__ int s_last;
    unsigned long long local = 0ull;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < mc.P; i += gridDim.x * blockDim.x) {
        const int np = min(dp.obs_cnt[i], mc.OBS - 1), nn = dp.nbr[i * mc.NBW];
        int c = 0;
        for (int ns = 0; ns < nn; ++ns) {
            dp.cum[i * mc.NBW + ns] = c;
            c += dp.plen[dp.nbr[i * mc.NBW + 1 + ns]];
        }
        dp.totlen[i] = c;
        unsigned long long pr = (unsigned long long)np * (unsigned long long)c;
        dp.pairs[i] = pr > 0x3fffffffull ? 0x3fffffff : (int)pr;
        local += pr;
        dp.chunks[i] = (dp.plen[i] + 31) >> 5;
    }
    if (local) atomicAdd(&dp.st->total_pairs, local);
    if (gridDim.x > 1) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(&dp.st->ticket_prep, 1) == (int)gridDim.x - 1;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
    }
    __syncthreads();
    block_exclusive_scan(dp.pairs, dp.rowbase, mc.P, wsum);
    block_exclusive_scan(dp.chunks, dp.chunk_off, mc.P, wsum);
    for (int i = threadIdx.x; i < mc.P; i += blockDim.x) {
        const int c0 = dp.chunk_off[i], c1 = c0 + dp.chunks[i];
        for (int c = c0; c < c1; ++c) dp.chunk_pyr[c] = i;
    }
    // the newborn normaliser (k_norm) runs BESIDE the C_z pass and takes each 1 / C_z as soon as it is there: "not there yet" is 0
    if (!mc.sharded) {
        const int n = dp.obs_capoff[mc.P];
        for (int t = threadIdx.x; t < n; t += blockDim.x) dp.INV[t] = 0.f;
    }
}
// position of pyramid a inside pyramid b's neighbour list (the relation is symmetric)
__device__ __forceinline__ int nb_index_of(const MapConst &mc, const DevPtrs &dp, int b, int a) {
    const int nn = dp.nbr[b * mc.NBW];
    for (int k = 0; k < nn; ++k)
        if (dp.nbr[b * mc.NBW + 1 + k] == a) return k;
    return -1;
}
__device__ __forceinline__ int chunk_to_pyramid(const int *chunk_off, int P, int c) {  // last a with chunk_off[a] <= c
    int lo = 0, hi = P;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunk_off[mid] <= c) lo = mid; else hi = mid;
    }
    return lo;
}
#define EVAL_THREADS 512
#define TILE_LD 33  // per-warp 32 x 32 staging tile of k_weight2w, padded
#define EVAL_GROUP 3     // neighbour pyramids per work item
#define EVAL_PTS 112     // float4 slots for a pyramid's points (OBS <= 128 is checked at create time; the default is 100)
#define EVAL_WARP_F4 (128 + 32)  // shared memory of one warp in float4: points, then the chunk's 32 particles
#define EVAL_SMEM_BYTES (((DSP_LUT_HALF + 3 + 31) & ~31) * 4 + (EVAL_THREADS / 32) * EVAL_WARP_F4 * 16 + (EVAL_THREADS / 32) * 8)
// queryNormalPDF's table index (dsp_dynamic.h:1294-1300) with the clamp moved behind the conversion:
//     |trunc(clamp(cx, -9.9, 9.9) * 1000 + 10000) - 10000|  ==  min(|trunc(cx * 1000 + 10000) - 10000|, 9900)
// for EVERY finite float cx (the conversion saturates; checked exhaustively over all 2^32 bit patterns by
// tests/test_host.py::test_pdf_index_clamp_can_move_behind_the_conversion), which turns three float compare / select
// instructions and a three-instruction absolute value into IABS + IMNMX.
__device__ __forceinline__ float dsp_pdf_i(const float *lut, float x, float mu, const FrameConst &fc) {
    const float cx = fc.fast_sigma ? dsp_div_known(x - mu, fc.sigma, fc.sigma_r) : (x - mu) / fc.sigma;
    const int i = __float2int_rz(cx * 1000 + 10000);
    const int j = (int)((unsigned)i - 10000u);
    return lut[min(abs(j), 9900)];
}
// The pair evaluation.  A work item is one chunk of 32 particles of pyramid a x up to EVAL_GROUP of the point pyramids i that
// see it; the (particle, point) tile of one (chunk, i) is nrows x np values that lie CONTIGUOUSLY in G (row-major rows of
// np), so the warp walks it flat: lane l takes pairs l, l + 32, ... and every store is a full coalesced line — no
// transposing tile.  The pyramid's observation points and the chunk's particles are staged in the warp's shared memory by
// 1-D bulk copies (cp.async.bulk, completion on an mbarrier): "per-pyramid observation points staged in shared memory via
// TMA" (BASELINE.json north_star).  Round 1's kernel read every point with a global load in front of its first use (23 % of
// all stall samples) and spent 20 % of its instructions on the transposed store.
// mode 0: all items; sharded: mode 1 = rows of the point pyramids this rank computes C_z for (i % nranks == rank),
// mode 2 = rows of the particle chunks this rank computes weights for (chunk % nranks == rank)
__global__ void __launch_bounds__(EVAL_THREADS) k_pair_eval(MapConst mc, FrameConst fc, DevPtrs dp, int mode) {
    extern __shared__ __align__(128) float sm[];
    float *lut = sm;
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    float4 *s_pts = reinterpret_cast<float4 *>(sm + ((DSP_LUT_HALF + 3 + 31) & ~31)) + wl * EVAL_WARP_F4;
    float4 *s_par = s_pts + 128;
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<float4 *>(sm + ((DSP_LUT_HALF + 3 + 31) & ~31)) + (EVAL_THREADS / 32) * EVAL_WARP_F4);
    uint64_t *bar = bars + wl;
    pdl_trigger();
    for (int i = threadIdx.x; i < DSP_LUT_HALF; i += blockDim.x) lut[i] = dp.lut[i];  // constant after create: staged before the wait
    if (lane == 0) {
        cuda::ptx::mbarrier_init(bar, 1);  // one arrival (the issuing lane's expect_tx) + the bytes
        cuda::ptx::fence_mbarrier_init(cuda::ptx::sem_release, cuda::ptx::scope_cluster);
    }
    pdl_wait();
    if (!use_pair_buffer(mc, dp)) return;
    __syncthreads();
    unsigned phase = 0u;
    const int nchunks = dp.chunk_off[mc.P];
    // (one queue ticket per (chunk, neighbour) would be ~21 000 same-address atomics per frame at cfg2)
    const int groups = (mc.NB + EVAL_GROUP - 1) / EVAL_GROUP;
    const int items = nchunks * groups;
    for (;;) {
        int it = 0;
        if (lane == 0) it = atomicAdd(mode == 2 ? &dp.st->work_eval2 : &dp.st->work_eval, 1);
        it = __shfl_sync(FULLMASK, it, 0);
        if (it >= items) break;
        const int c = it / groups, g0 = (it - c * groups) * EVAL_GROUP;
        if (mode == 2 && c % mc.nranks != mc.rank) continue;
        const int a = dp.chunk_pyr[c];
        const int nn = dp.nbr[a * mc.NBW];
        if (g0 >= nn) continue;
        const int k0 = (c - dp.chunk_off[a]) << 5;
        const int nrows = min(32, dp.plen[a] - k0);
        bool have_particles = false;
        for (int ns = g0; ns < min(nn, g0 + EVAL_GROUP); ++ns) {
            const int i = dp.nbr[a * mc.NBW + 1 + ns];  // a point pyramid that sees pyramid a
            if (mode == 1 && i % mc.nranks != mc.rank) continue;
            const int np = min(dp.obs_cnt[i], mc.OBS - 1);
            if (np == 0) continue;
            __syncwarp();  // every lane is done with the previous tile's points
            if (lane == 0) {
                cuda::ptx::fence_proxy_async(cuda::ptx::space_shared);  // earlier generic reads of the buffers before the async writes
                const unsigned bytes = (unsigned)np * 16u + (have_particles ? 0u : (unsigned)nrows * 16u);
                cuda::ptx::mbarrier_arrive_expect_tx(cuda::ptx::sem_release, cuda::ptx::scope_cta, cuda::ptx::space_shared, bar, bytes);
                cuda::ptx::cp_async_bulk(cuda::ptx::space_cluster, cuda::ptx::space_global, s_pts, dp.OBSP + (size_t)i * mc.OBS, (unsigned)np * 16u, bar);
                if (!have_particles)
                    cuda::ptx::cp_async_bulk(cuda::ptx::space_cluster, cuda::ptx::space_global, s_par, dp.LP + dp.poff[a] + k0, (unsigned)nrows * 16u, bar);
            }
            have_particles = true;
            // the tile's place in G while the copies fly
            float *gb = dp.G + (size_t)dp.rowbase[i] + (size_t)(dp.cum[i * mc.NBW + dp.nbrev[a * mc.NBW + 1 + ns]] + k0) * np;
            const int total = nrows * np;
            const unsigned magic = np > 1 ? 0xffffffffu / (unsigned)np + 1u : 0u;  // exact f / np for f < 65536
            while (!cuda::ptx::mbarrier_try_wait_parity(bar, phase)) {}
            phase ^= 1u;
#pragma unroll 2
            for (int f = lane; f < total; f += 32) {
                const int r = np > 1 ? (int)__umulhi((unsigned)f, magic) : f;
                const int z = f - r * np;
                const float4 p = s_par[r], o = s_pts[z];
                gb[f] = dsp_pdf_i(lut, p.x, o.x, fc) * dsp_pdf_i(lut, p.y, o.y, fc) * dsp_pdf_i(lut, p.z, o.z, fc);
            }
        }
    }
}
// One point's chain over `cur` rows of a staged tile: acc += w[j] * t[j * np] for j = 0 .. cur-1, IN THAT ORDER (one dependent
// fp32 add per row, 4 cycles each: the floor of the whole C_z pass is the longest chain).  The products of the next eight rows
// are formed while the current eight are added, so the shared-memory latency (~30 cycles) never sits in front of the adds.
__device__ __forceinline__ float cz_chain_rows(float acc, const float *t, const float *w, int np, int cur) {
    int jj = 0;
    if (cur >= 8) {
        float a0 = w[0] * t[0], a1 = w[1] * t[np], a2 = w[2] * t[2 * np], a3 = w[3] * t[3 * np];
        float a4 = w[4] * t[4 * np], a5 = w[5] * t[5 * np], a6 = w[6] * t[6 * np], a7 = w[7] * t[7 * np];
        for (jj = 8; jj + 8 <= cur; jj += 8) {
            const float *tn = t + jj * np;
            const float *wn = w + jj;
            const float b0 = wn[0] * tn[0], b1 = wn[1] * tn[np], b2 = wn[2] * tn[2 * np], b3 = wn[3] * tn[3 * np];
            const float b4 = wn[4] * tn[4 * np], b5 = wn[5] * tn[5 * np], b6 = wn[6] * tn[6 * np], b7 = wn[7] * tn[7 * np];
            acc += a0; acc += a1; acc += a2; acc += a3; acc += a4; acc += a5; acc += a6; acc += a7;
            a0 = b0; a1 = b1; a2 = b2; a3 = b3; a4 = b4; a5 = b5; a6 = b6; a7 = b7;
        }
        acc += a0; acc += a1; acc += a2; acc += a3; acc += a4; acc += a5; acc += a6; acc += a7;
    }
    for (; jj < cur; ++jj) acc += w[jj] * t[jj * np];
    return acc;
}
// C_z (dsp_dynamic.h:709-739): a CTA per point pyramid streams the pyramid's contiguous block of G through shared
// memory (double-buffered cp.async); thread z < np adds its column in list order: one fp32 chain per point.
// Measured on B200 (cfg2): 256 threads with 2 x 32 KB tiles (3 CTAs / SM) 54 us, 128 threads with 2 x 16 KB tiles
// (6 CTAs / SM, every pyramid resident at once) 64 us — the longer tiles amortise the per-tile barrier better.
// A three-stage variant in which the warps without a column multiplied the landed tile by its row weights in place, so that
// the chain was one load and one add per term, was bit-identical and slower: 81 vs 66 us (profiles/r02_variants.jsonl).
template <int CZ_THREADS, int CZ_TILE, int CZ_JT>
__global__ void __launch_bounds__(CZ_THREADS) k_cz_chain(MapConst mc, FrameConst fc, DevPtrs dp) {
    pdl_enter();
    extern __shared__ float czsm[];
    float *tile0 = czsm, *tile1 = czsm + CZ_TILE + 8;  // + room for the alignment phase
    float *pws0 = czsm + 2 * (CZ_TILE + 8), *pws1 = pws0 + CZ_JT;
    __shared__ int s_item;
    if (!use_pair_buffer(mc, dp)) return;
    const float enb = fc.nb_weight * (float)dp.st->n_valid * (float)fc.nb_num;  // :292
    const float add_k = enb + fc.kappa;
    const int tid = threadIdx.x;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(&dp.st->work_k4, 1);
        __syncthreads();
        const int i = s_item;
        if (i >= mc.P) break;
        const int np = min(dp.obs_cnt[i], mc.OBS - 1);
        if (np == 0) continue;
        if (mc.sharded && i % mc.nranks != mc.rank) continue;  // another rank computes this pyramid's C_z
        const int nn = dp.nbr[i * mc.NBW];
        const int JT = min(CZ_JT, CZ_TILE / np);
        const float *gsrc = dp.G + (size_t)dp.rowbase[i];
        auto len_of = [&](int k) { return dp.plen[dp.nbr[i * mc.NBW + 1 + k]]; };
        auto off_of = [&](int k) { return dp.poff[dp.nbr[i * mc.NBW + 1 + k]]; };
        // tile iterator over (neighbour ns, particle offset k0); the "issue" state runs one tile ahead of the consumer
        int ins = 0, ik0 = 0, iln = nn > 0 ? len_of(0) : 0;
        int ph0 = 0, ph1 = 0;
        const float *ig = gsrc;
        auto issue = [&](int buf) -> int {  // returns the number of particle rows of the issued tile, 0 when exhausted
            while (ins < nn && ik0 >= iln) {
                ++ins;
                ik0 = 0;
                iln = ins < nn ? len_of(ins) : 0;
            }
            if (ins >= nn) return 0;
            const int cur = min(JT, iln - ik0), nfl = cur * np;
            float *t = buf ? tile1 : tile0;
            // 16-byte copies: start at the 16 B boundary below the tile and keep the source's phase inside the buffer
            const int ph = (int)((reinterpret_cast<size_t>(ig) >> 2) & 3);
            const float *src = ig - ph;
            const int nq = (ph + nfl + 3) >> 2;
            for (int q = tid; q < nq; q += CZ_THREADS) __pipeline_memcpy_async(t + 4 * q, src + 4 * q, 16);
            if (buf) ph1 = ph; else ph0 = ph;
            if (tid < cur) __pipeline_memcpy_async((buf ? pws1 : pws0) + tid, dp.PW + off_of(ins) + ik0 + tid, 4);
            ig += nfl;
            ik0 += cur;
            return cur;
        };
        float acc = 0.f;
        int cur = issue(0);
        __pipeline_commit();
        int buf = 0;
        while (cur > 0) {
            const int nxt = issue(buf ^ 1);
            __pipeline_commit();
            __pipeline_wait_prior(1);
            __syncthreads();
            if (tid < np) {  // the chain: (P_d * w) * g added in list order, one fp32 add per term
                acc = cz_chain_rows(acc, (buf ? tile1 : tile0) + (buf ? ph1 : ph0) + tid, buf ? pws1 : pws0, np, cur);
            }
            __syncthreads();
            cur = nxt;
            buf ^= 1;
        }
        __pipeline_wait_prior(0);
        if (tid < np) {
            acc += add_k;
            dp.CZ[i * mc.OBS + tid] = acc;
            dp.INV[dp.obs_capoff[i] + tid] = 1.f / acc;  // for the newborn normaliser (:802)
        }
    }
}
// (A bulk-copy ring for this pass — a producer thread streaming stages with cp.async.bulk on mbarriers, chain warps releasing
// them — was measured twice: 92.5 us with 16 KB stages in round 1, 56.8 - 63.5 us with 33 KB stages in round 2, against 54.6 -
// 59.6 us for the kernel above; deleted.  profiles/r02_variants.jsonl.)
// weights (dsp_dynamic.h:743-790): a CTA per 32 particles of a pyramid.  For one neighbour pyramid at a time, ALL threads
// turn the chunk's contiguous 32 x np tile of G into quotient terms (P_d * g) / C_z (flat, coalesced loads); then warp 0,
// lane = particle, adds its row in bin order.  Neighbours are visited in table order, so each particle's sum is one fp32
// chain in the reference's order.  Two term buffers let the next neighbour's divisions overlap the current chain.
#ifndef W2_SWITCH
#define W2_SWITCH 4096  // chunks of 32 particles: below, CTA-per-chunk (k_weight2); from here on, warp-per-chunk (k_weight2w)
#endif
#ifndef W2_THREADS
#define W2_THREADS 128   // warp 0 adds the chains, the other warps turn tiles into quotient terms
#endif
#define W2_NP 100  // padded row length (np <= 99; odd stride: no bank conflicts in the chain)
__global__ void __launch_bounds__(W2_THREADS) k_weight2(MapConst mc, FrameConst fc, DevPtrs dp) {
    pdl_enter();
    __shared__ float terms[2][32 * (W2_NP + 1)];
    __shared__ float czs[2][W2_NP];
    __shared__ int s_item;
    if (!use_pair_buffer(mc, dp)) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nchunks = dp.chunk_off[mc.P];
    if (nchunks >= W2_SWITCH) return;  // many chunks: k_weight2w takes the frame
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(&dp.st->work_w2, 1);
        __syncthreads();
        const int c = s_item;
        if (c >= nchunks) break;
        const int a = chunk_to_pyramid(dp.chunk_off, mc.P, c);
        const int k0 = (c - dp.chunk_off[a]) << 5;
        const int ln = dp.plen[a], lb = dp.poff[a];
        const int nrows = min(32, ln - k0);
        const int nn = dp.nbr[a * mc.NBW];
        if (mc.sharded && c % mc.nranks != mc.rank) continue;  // chunks are dealt round-robin, whoever owns the particles
        // chain state lives in warp 0 (lane = particle)
        bool act = false;
        float pw = 0.f, sum = 0.f;
        if (wid == 0 && lane < nrows) {
            const float4 p = dp.LP[lb + k0 + lane];
            pw = p.w;
            const float dist = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
            const float maxlen = __int_as_float(dp.obs_maxbits[a]);
            act = !(maxlen > 0.f && dist > maxlen + mc.occl);  // occluded particles keep their weight (:761)
        }
        int buf = 0, prev_np = 0, prev_ld = 1;
        // Producer state: the first eight G values and the (at most two) C_z values of the NEXT neighbour are requested
        // before the current phase's barrier, so their L2 / HBM latency passes while warp 0 adds the previous chain.
        // (Measured alternatives on B200, cfg2: no prefetch 93 us; lane = bin / warp = particle rows, which needs a third
        // of the instructions per division but leaves np / 32 of the lanes idle, 104 us; this version 91 us.)
        const int t96 = tid - 32;
        const int PSTR = W2_THREADS - 32;
        float gpre[8], czpre0 = 1.f, czpre1 = 1.f;
        const float *gb_n = nullptr;
        int nfl_n = 0;
        auto prefetch = [&](int ns) {
            nfl_n = 0;
            if (ns >= nn) return;
            const int b = dp.nbr[a * mc.NBW + 1 + ns];
            const int np = min(dp.obs_cnt[b], mc.OBS - 1);
            if (np == 0) return;
            gb_n = dp.G + (size_t)dp.rowbase[b] + (size_t)(dp.cum[b * mc.NBW + nb_index_of(mc, dp, b, a)] + k0) * np;
            nfl_n = nrows * np;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int f = t96 + u * PSTR;
                gpre[u] = f < nfl_n ? __ldg(gb_n + f) : 0.f;
            }
            const float *cz = dp.CZ + (size_t)b * mc.OBS;
            czpre0 = t96 < np ? cz[t96] : 1.f;
            czpre1 = t96 + PSTR < np ? cz[t96 + PSTR] : 1.f;
        };
        if (wid > 0) prefetch(0);
        for (int ns = 0; ns <= nn; ++ns) {
            // phase 1 (warps 1..3): quotient terms of neighbour ns into terms[buf], while warp 0 runs the previous chain
            int np = 0, ld = 1;
            if (ns < nn) {
                const int b = dp.nbr[a * mc.NBW + 1 + ns];
                np = min(dp.obs_cnt[b], mc.OBS - 1);
                ld = np | 1;
            }
            if (wid > 0 && ns < nn) {
                if (np > 0) {
                    if (t96 < np) czs[buf][t96] = czpre0;
                    if (t96 + PSTR < np) czs[buf][t96 + PSTR] = czpre1;
                    const float *gb = gb_n;
                    const int nfl = nfl_n;
                    asm volatile("bar.sync 1, %0;" ::"n"(W2_THREADS - 32) : "memory");  // only the producer warps
                    const unsigned magic = np > 1 ? 0xffffffffu / (unsigned)np + 1u : 0u;  // exact f / np for f < 65536
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int f = t96 + u * PSTR;
                        if (f < nfl) {
                            const int r = np > 1 ? (int)__umulhi((unsigned)f, magic) : f;
                            const int z = f - r * np;
                            terms[buf][r * ld + z] = (fc.Pd * gpre[u]) / czs[buf][z];
                        }
                    }
                    // tiles of more than 768 terms (np > 24): further batches of eight loads in flight per thread
                    for (int f0 = t96 + 8 * PSTR; f0 < nfl; f0 += 8 * PSTR) {
                        float g[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int f = f0 + u * PSTR;
                            g[u] = f < nfl ? __ldg(gb + f) : 0.f;
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int f = f0 + u * PSTR;
                            if (f < nfl) {
                                const int r = np > 1 ? (int)__umulhi((unsigned)f, magic) : f;
                                const int z = f - r * np;
                                terms[buf][r * ld + z] = (fc.Pd * g[u]) / czs[buf][z];
                            }
                        }
                    }
                }
                prefetch(ns + 1);
            }
            // phase 2 (warp 0): add the previous neighbour's terms, particle rows in bin order
            if (wid == 0 && act && prev_np > 0) {
                const float *t = terms[buf ^ 1] + lane * prev_ld;
#pragma unroll 8
                for (int z = 0; z < prev_np; ++z) sum += t[z];
            }
            __syncthreads();
            prev_np = np;
            prev_ld = ld;
            buf ^= 1;
        }
        if (wid == 0 && lane < nrows) {
            const float w_new = act ? pw * (fc.one_minus_Pd + sum) : pw;
            if (mc.sharded) dp.NW[lb + k0 + lane] = w_new;  // merged over ranks, applied by the particle's owner
            else if (act) dp.PA[dp.LA[lb + k0 + lane]].w = w_new;
        }
    }
}
// weights (dsp_dynamic.h:743-790), variant for MANY chunks (>= W2_SWITCH): one warp per 32 particles, lanes = particles.  The 32 x 32 sub-tiles of G
// are read as coalesced rows (lane = point) into registers one sub-tile AHEAD of the one being consumed, staged through a
// per-warp shared tile, and each lane adds its particle's row in (neighbour-table, bin) order.
#define W2W_THREADS 256
__global__ void __launch_bounds__(W2W_THREADS, 3) k_weight2w(MapConst mc, FrameConst fc, DevPtrs dp) {
    pdl_enter();
    __shared__ float tiles[(W2W_THREADS / 32) * 32 * TILE_LD];
    __shared__ float czall[(W2W_THREADS / 32) * 32];
    if (!use_pair_buffer(mc, dp)) return;
    if (dp.chunk_off[mc.P] < W2_SWITCH) return;  // few chunks: the CTA-per-chunk kernel has more parallelism
    float *tile = tiles + (threadIdx.x >> 5) * (32 * TILE_LD);
    float *czs = czall + (threadIdx.x >> 5) * 32;
    const int lane = threadIdx.x & 31;
    const int nchunks = dp.chunk_off[mc.P];
    for (;;) {
        int c = 0;
        if (lane == 0) c = atomicAdd(&dp.st->work_w2, 1);
        c = __shfl_sync(FULLMASK, c, 0);
        if (c >= nchunks) break;
        if (mc.sharded && c % mc.nranks != mc.rank) continue;  // chunks are dealt round-robin over the ranks
        const int a = chunk_to_pyramid(dp.chunk_off, mc.P, c);
        const int k0 = (c - dp.chunk_off[a]) << 5;
        const int ln = dp.plen[a], lb = dp.poff[a];
        const int nrows = min(32, ln - k0);
        bool act = lane < nrows;
        const float4 p = act ? dp.LP[lb + k0 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (act) {
            const float dist = sqrtf(p.x * p.x + p.y * p.y + p.z * p.z);
            const float maxlen = __int_as_float(dp.obs_maxbits[a]);
            if (maxlen > 0.f && dist > maxlen + mc.occl) act = false;  // occluded (:761): weight unchanged
        }
        const int nn = dp.nbr[a * mc.NBW];
        // sub-tile iterator over (neighbour ns, point block z0)
        int ns = -1, z0 = 0, np = 0;
        const float *gb = nullptr, *cz = nullptr;
        auto advance = [&]() -> int {  // moves to the next sub-tile; returns its number of points, 0 when done
            z0 += 32;
            while (ns < 0 || z0 >= np) {
                if (++ns >= nn) return 0;
                const int b = dp.nbr[a * mc.NBW + 1 + ns];
                np = min(dp.obs_cnt[b], mc.OBS - 1);
                z0 = 0;
                if (np == 0) continue;
                gb = dp.G + (size_t)dp.rowbase[b] + (size_t)(dp.cum[b * mc.NBW + nb_index_of(mc, dp, b, a)] + k0) * np;
                cz = dp.CZ + (size_t)b * mc.OBS;
            }
            return min(32, np - z0);
        };
        float v[32], czv = 1.f;
        auto prefetch = [&](int nsub) {
            if (lane < nsub) {
                const float *src = gb + z0 + lane;
#pragma unroll
                for (int r = 0; r < 32; ++r) v[r] = r < nrows ? __ldg(src + (size_t)r * np) : 0.f;
                czv = cz[z0 + lane];
            }
        };
        float sum = 0.f;
        int nsub = advance();
        if (nsub) prefetch(nsub);
        while (nsub) {
            if (lane < nsub) {
#pragma unroll
                for (int r = 0; r < 32; ++r) tile[r * TILE_LD + lane] = v[r];
                czs[lane] = czv;
            }
            __syncwarp();
            const int cur = nsub;
            nsub = advance();
            if (nsub) prefetch(nsub);  // in flight while the current sub-tile is consumed
            if (act) {
#pragma unroll 4
                for (int zl = 0; zl < cur; ++zl) sum += (fc.Pd * tile[lane * TILE_LD + zl]) / czs[zl];
            }
            __syncwarp();
        }
        if (lane < nrows) {
            const float w_new = act ? p.w * (fc.one_minus_Pd + sum) : p.w;
            if (mc.sharded) dp.NW[lb + k0 + lane] = w_new;  // merged over ranks, applied by the particle's owner
            else if (act) dp.PA[dp.LA[lb + k0 + lane]].w = w_new;
        }
    }
}

// (A warp-per-chunk weight kernel without block-wide barriers — the chunk's tile walked flat by the 32 lanes, quotients through
// an exact fast path for zero / subnormal dividends — took 207 us against 86 us for k_weight2: nine tiles in sequence per warp
// is too long a chain.  Deleted; profiles/r02_variants.jsonl.)

// Exhaustive check of dsp_div_known for ONE divisor over every float a with |a| <= max (both signs).  What has to be
// identical to IEEE division is the integer the quotient is turned into, so that is what is compared:
//   mode 0: (int)(a / b)                                   voxel coordinates (dsp_dynamic.h:1078-1080)
//   mode 1: (int)(clamp(a / b, +-9.9) * 1000 + 10000)      PDF table index   (dsp_dynamic.h:1296-1300)
__device__ __forceinline__ int verify_index(float q, int mode) {
    if (mode == 0) return (int)q;
    if (q > 9.9f) q = 9.9f;
    else if (q < -9.9f) q = -9.9f;
    return (int)(q * 1000 + 10000);
}
__global__ void k_verify_div(float b, float r, unsigned max_bits, int mode, int *bad) {
    pdl_enter();
    int local = 0;
    for (unsigned long long u = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; u <= max_bits; u += (unsigned long long)gridDim.x * blockDim.x) {
        const float a = __uint_as_float((unsigned)u);
        if (verify_index(a / b, mode) != verify_index(dsp_div_known(a, b, r), mode)) ++local;
        if (verify_index((-a) / b, mode) != verify_index(dsp_div_known(-a, b, r), mode)) ++local;
    }
    if (local) atomicAdd(bad, local);
}


// K6a  newborn normaliser (dsp_dynamic.h:799-805): sum of 1/C_z over (pyramid, bin) order, one fp32 chain.
// poll: the kernel was launched beside the C_z pass (k_pair_prep zeroed INV): a value of 0 has not been written yet (1 / C_z is
// never 0), so every load waits for its value; the serial chain of thread 0 then ends a few microseconds behind the C_z pass
// instead of starting there (beside the weight pass it ran three times slower than alone and ended 30 - 80 us after it).
// (The wait is bounded, about 1 ms per value: if the C_z pass never runs beside this kernel — a profiler that serialises
// kernels, DSPMAP_NORM_POLL=0 is for that — the frame ends with capacity code 64 instead of a hung device.)
__device__ __forceinline__ float norm_take(const float *p, int poll, int *overflow) {
    if (!poll) return *p;
    float v = 0.f;
    for (int it = 0; it < 5000; ++it) {
        v = *(const volatile float *)p;
        if (v != 0.f) return v;
#ifdef __CUDA_ARCH__
        __nanosleep(200);
#endif
    }
    atomicOr(overflow, 64);
    return v;
}
__global__ void k_norm(MapConst mc, FrameConst fc, DevPtrs dp, int poll) {
    pdl_enter();
    __shared__ __align__(16) float buf[2][1024];
    const int n = dp.obs_capoff[mc.P];
    float acc = 0.f;
    int nchunk = (n + 1023) / 1024;
    if (nchunk > 0)
        for (int t = threadIdx.x; t < 1024; t += blockDim.x) buf[0][t] = t < n ? norm_take(dp.INV + t, poll, &dp.st->overflow) : 0.f;
    __syncthreads();
    for (int c = 0; c < nchunk; ++c) {
        int nb = (c + 1) & 1, b0 = (c + 1) * 1024;
        if (threadIdx.x >= 32) {  // warps 1.. prefetch the next chunk while warp 0 adds the current one
            if (c + 1 < nchunk)
                for (int t = threadIdx.x - 32; t < 1024; t += blockDim.x - 32) buf[nb][t] = b0 + t < n ? norm_take(dp.INV + b0 + t, poll, &dp.st->overflow) : 0.f;
        } else if (threadIdx.x == 0) {
            int cnt = min(1024, n - c * 1024);
            const float4 *s4 = reinterpret_cast<const float4 *>(buf[c & 1]);
            const int c4 = cnt >> 2;
#pragma unroll 4
            for (int k = 0; k < c4; ++k) {
                float4 x = s4[k];
                acc += x.x; acc += x.y; acc += x.z; acc += x.w;
            }
            for (int k = c4 << 2; k < cnt; ++k) acc += buf[c & 1][k];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        dp.st->norm = acc;
        dp.st->w_new = fc.nb_weight * acc;
    }
}

// ------------------------------------------------------------------------------------------------------------
// K6  newborn particles (dsp_dynamic.h:796-921; dsp_static.h:779-829)
// ------------------------------------------------------------------------------------------------------------
// point pass 0: corrected point, its voxel, in-map predicate (:817-827,:846-848; the static variant has no test)
__global__ void k_nb_point0(MapConst mc, FrameConst fc, DevPtrs dp) {
    pdl_enter();
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < fc.n_tagged; m += gridDim.x * blockDim.x) {
        const float *pt = dp.tagged + 7 * m;
        float cx = pt[0] - fc.cur[0], cy = pt[1] - fc.cur[1], cz = pt[2] - fc.cur[2];
        int pv = dsp_voxel_index(mc, cx, cy, cz);
        dp.NPC[m] = make_float4(cx, cy, cz, __int_as_float(pv));
        dp.ninmap[m] = (mc.model == 1) ? !(fc.tagged_padded && !(pt[0] < 1e29f)) : (pv >= 0);
        dp.nimask[m] = 0ull;
    }
}
__device__ __forceinline__ u64 bits_below(int p) { return p >= 64 ? ~0ull : ((1ull << p) - 1ull); }
// which of a point's nb_num candidates land inside the map (:871-875): one thread per candidate
__global__ void k_nb_mask(MapConst mc, FrameConst fc, DevPtrs dp) {
    pdl_enter();
    const int total = fc.n_tagged * fc.nb_num;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        int m = t / fc.nb_num, p = t - m * fc.nb_num;
        if (!dp.ninmap[m]) continue;
        float4 pc = dp.NPC[m];
        long long c = (dp.st->p_cur + 3ll * ((long long)dp.nrank[m] * fc.nb_num + p)) % mc.G;
        float px = pc.x + dp.ptab[c];
        float py = pc.y + dp.ptab[(c + 1) % mc.G];
        float pz = pc.z + dp.ptab[(c + 2) % mc.G];
        if (dsp_voxel_index(mc, px, py, pz) >= 0) atomicOr(&dp.nimask[m], 1ull << p);
    }
}
// point pass 1: Dempster-Shafer split from the resident particles of the point's voxel (:829-866) — one warp per point,
// lanes = slots, the three weight sums added in slot order — and how many table / uniform draws the point consumes.
// phase 0: single GPU (split + counts); phase 1: sharded, split by the owner of the point's voxel only; phase 2: sharded, counts
__device__ __forceinline__ void nb_point1_body(const MapConst &mc, const FrameConst &fc, const DevPtrs &dp, int phase) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int R = (mc.S + 31) >> 5;
    for (int m = warp; m < fc.n_tagged; m += nwarps) {
        if (!dp.ninmap[m]) {
            if (lane == 0) {
                if (phase != 1) { dp.nvcnt[m] = 0; dp.nrcnt[m] = 0; }
            }
            continue;
        }
        int n_static = 0;
        if (phase == 2) n_static = (int)dp.nst_shared[m];  // summed over ranks: exactly one owner contributed
        const int pv0 = __float_as_int(dp.NPC[m].w);
        if (phase == 1 && (mc.model != 0 || pv0 < mc.v_lo || pv0 >= mc.v_hi)) {  // not this rank's voxel
            continue;  // nst_shared was zeroed: another rank (or nobody) contributes
        }
        if (mc.model == 0 && phase != 2) {
            const int pv = pv0;
            const ulonglong2 msk = dp.M[pv];
            float ws = 0.f, wd = 0.f, wsd = 0.f;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (r >= R) break;
                const unsigned live = r < 2 ? (unsigned)(msk.x >> (32 * r)) : (unsigned)(msk.y >> (32 * (r - 2)));
                if (live == 0u) continue;
                float w = 0.f;
                int cls = 3;
                if ((live >> lane) & 1u) {
                    const int a = pv * mc.S + 32 * r + lane;
                    const float4 B = dp.PB[a];
                    if (B.w > 0.9f && B.w < 14.f) {  // not newborn (:830)
                        w = dp.PA[a].w;
                        const float va = fabsf(B.x) + fabsf(B.y) + fabsf(B.z);
                        cls = va < 0.1f ? 0 : (va < 0.5f ? 1 : 2);
                    }
                }
                unsigned k = __ballot_sync(FULLMASK, cls != 3);
                while (k) {
                    const int l = __ffs(k) - 1;
                    k &= k - 1;
                    const float wv = __shfl_sync(FULLMASK, w, l);
                    const int cl = __shfl_sync(FULLMASK, cls, l);
                    if (cl == 0) ws += wv;
                    else if (cl == 1) wsd += wv;
                    else wd += wv;
                }
            }
            const float tot = ws + wd + wsd;
            const float m_s = ws / tot, m_d = wd / tot, m_sd = wsd / tot;
            const float p_s = (m_s + m_s + m_sd) * 0.5f, p_d = (m_d + m_d + m_sd) * 0.5f;
            const float np_ = p_s + p_d;
            const float ps_n = p_s / np_;
            const float prod = (float)fc.nb_model_gen * ps_n;
            n_static = (prod != prod) ? INT_MIN : (int)prod;
            n_static = max(fc.nb_min_static, n_static);
        }
        if (phase == 1) {
            if (lane == 0) dp.nst_shared[m] = (float)n_static;
            continue;
        }
        if (lane == 0) {
            const u64 im = dp.nimask[m];
            const float *pt = dp.tagged + 7 * m;
            u64 vm = 0ull, rm = 0ull;
            if (mc.model == 0 && pt[6] > 0.01f) {  // only points tagged dynamic draw velocities (:883,:894)
                u64 nonstatic = im & ~bits_below(max(n_static, 0));
                u64 est = (pt[3] > -100.f) ? bits_below(fc.nb_model_gen) : 0ull;  // (:881)
                vm = nonstatic & est;
                rm = nonstatic & ~est;
            }
            dp.nstatic[m] = n_static;
            dp.nvcnt[m] = __popcll(vm);
            dp.nrcnt[m] = __popcll(rm);
        }
    }
}
// (Fusing the two scans of the draw counts into this kernel's last block — 256 threads, 40 elements each, one scan after the
// other — cost 24 us per frame against the separate 2 x 1024-thread launch: profiles/r02_variants.jsonl.)
__global__ void __launch_bounds__(256) k_nb_point1(MapConst mc, FrameConst fc, DevPtrs dp, int phase) {
    pdl_enter();
    nb_point1_body(mc, fc, dp, phase);
}
// Newborn particles are born in two steps.  WHERE a candidate lands — its position (:871-873), its voxel, and which free slot
// of that voxel it takes in the reference's serial (point, candidate) order (addAParticle, :1183-1201) — depends only on
// the cloud, the position-noise table and the occupancy masks after the arrival pass, so k_nb_cand, the grouping kernels
// and k_nb_place run EARLY, beside the observation passes.  WHAT it carries — the velocity class (:877-907: it needs the
// Dempster-Shafer split, i.e. the weights the observation update has just produced) and the weight (:909: it needs
// sum 1/C_z) — is filled in by k_nb_fill after the weight pass.  Slots born early carry the newborn flag 15 from the start,
// so the split (which skips newborn particles, :830) and everything else that walks a voxel sees them as the reference does.
__global__ void k_nb_cand(MapConst mc, FrameConst fc, DevPtrs dp) {
    pdl_enter();
    const int total = fc.n_tagged * fc.nb_num;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        int m = t / fc.nb_num, p = t - m * fc.nb_num;
        u64 im = dp.nimask[m];
        if (!((im >> p) & 1ull)) continue;
        float4 pc = dp.NPC[m];
        long long c = (dp.st->p_cur + 3ll * ((long long)dp.nrank[m] * fc.nb_num + p)) % mc.G;
        float px = pc.x + dp.ptab[c];
        float py = pc.y + dp.ptab[(c + 1) % mc.G];
        float pz = pc.z + dp.ptab[(c + 2) % mc.G];
        int d = dsp_voxel_index(mc, px, py, pz);
        if (mc.sharded && (d < mc.v_lo || d >= mc.v_hi)) continue;  // another rank places this candidate
        int k = agg_inc(&dp.st->n_cand);
        if (k >= dp.cap_cand) { atomicOr(&dp.st->overflow, 2); continue; }
        dp.CA[k] = make_float4(px, py, pz, 0.f);
        dp.Caddr[k] = -1;  // not born (yet)
        dp.Ckey[k] = m * DSP_MAX_NB_NUM + p;
        dp.Cdst[k] = d;
        if (atomicAdd(&dp.ccnt[d], 1) == 0) dp.cowner[agg_inc(&dp.st->n_cand_owner)] = d;
    }
}
// velocity (:877-907) and weight (:909) of the candidates that were born
__global__ void k_nb_fill(MapConst mc, FrameConst fc, DevPtrs dp, u64 useed) {
    pdl_enter();
    const int n = min(dp.st->n_cand, dp.cap_cand);
    const float w_new = dp.st->w_new;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int a = dp.Caddr[k];
        if (a < 0) continue;
        const int key = dp.Ckey[k], m = key / DSP_MAX_NB_NUM, p = key - m * DSP_MAX_NB_NUM;
        float vx = 0.f, vy = 0.f, vz = 0.f;
        const float *pt = dp.tagged + 7 * m;
        if (mc.model == 0 && pt[6] > 0.01f) {
            const u64 im = dp.nimask[m];
            int n_static = dp.nstatic[m];
            if (p >= n_static) {
                u64 nonstatic = im & ~bits_below(max(n_static, 0));
                u64 est = (pt[3] > -100.f) ? bits_below(fc.nb_model_gen) : 0ull;
                if ((est >> p) & 1ull) {
                    int j = __popcll(nonstatic & est & bits_below(p));
                    long long vc = (dp.st->v_cur + 3ll * (dp.nvoff[m] + j)) % mc.G;
                    vx = pt[3] + 4 * dp.vtab[vc];
                    vy = pt[4] + 4 * dp.vtab[(vc + 1) % mc.G];
                    vz = pt[5] + 4 * dp.vtab[(vc + 2) % mc.G];
                } else {
                    int j = __popcll(nonstatic & ~est & bits_below(p));
                    u64 uc = (u64)dp.st->u_cur + 3ull * (u64)(dp.nroff[m] + j);
                    vx = dsp_uniform(useed, uc, -1.5f, 1.5f);
                    vy = dsp_uniform(useed, uc + 1, -1.5f, 1.5f);
                    vz = dsp_uniform(useed, uc + 2, -0.5f, 0.5f);
                }
            }
            vz = 0.f;  // LIMIT_MOVEMENT_IN_XY_PLANE (:905-907)
        }
        dp.PB[a] = make_float4(vx, vy, vz, 15.f);
        dp.PA[a].w = w_new;
    }
}
// addAParticle (:1183-1201) in serial order: the k-th candidate of a voxel (in (point, candidate) order) takes its k-th free
// slot.  One thread per candidate, walking the grouped segments: it counts the keys of its voxel's segment below its own
// (keys are unique; neighbouring threads share the segment, so the reads are broadcasts from L1), and if that rank is below
// the number of free slots of the voxel's mask — the snapshot MS the grouping pass took, i.e. the mask after the arrival pass
// — it takes the rank-th free slot.  No rounds, no shuffles: every candidate decides for itself.
// (Round 1 gave a warp to each destination voxel and extracted the minimum key once per free slot: 38 us at cfg2.)
// Voxels with MANY candidates (a few hundred to a few thousand: voxels near the sensor, where the rays are dense) used to set
// the kernel's duration — every one of their candidates counted the whole segment, 85 us at cfg2 once the map is populated.
// Above NBP_HEAVY candidates a thread first counts only against the segment's first NBP_SAMPLE keys (the order inside a
// segment is arbitrary, so they are a sample): if that many of those alone are smaller than its key as the voxel has free
// slots, the candidate is not born — exactly, no estimate involved.  The few survivors of a block (those whose key is small
// enough to have a chance) are then ranked one after the other by whole warps, 32 keys of the segment per step.
#define NBP_HEAVY 96
#define NBP_SAMPLE 96
__device__ __forceinline__ void nb_place_born(const MapConst &mc, const DevPtrs &dp, int cand, int d, ulonglong2 msk, int rank) {
    const int slot = mask_nth_free(mc, msk, rank);
    const int a = d * mc.S + slot;
    dp.PA[a] = dp.CA[cand];                        // position; the weight follows in k_nb_fill
    dp.PB[a] = make_float4(0.f, 0.f, 0.f, 15.f);   // newborn flag now, velocity in k_nb_fill
    dp.Caddr[cand] = a;
    mask_atomic_set(dp.M, d, slot);
}
__global__ void __launch_bounds__(256) k_nb_place(MapConst mc, FrameConst fc, DevPtrs dp) {
    pdl_enter();
    __shared__ int s_surv[256];
    __shared__ int s_nsurv;
    const int n = min(dp.st->cand_top, dp.cap_cand);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int born = 0;
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {  // (uniform per block: barriers inside)
        if (threadIdx.x == 0) s_nsurv = 0;
        __syncthreads();
        const int pos = base + threadIdx.x;
        if (pos < n) {
            const int cand = dp.csegi[pos];
            const int d = dp.Cdst[cand];
            const ulonglong2 msk = dp.MS[d];
            const int nfree = mask_free(mc, msk);
            if (nfree > 0) {
                const int b = dp.cbase[d], c = dp.ccnt[d], key = dp.cseg[pos];
                const int first = c > NBP_HEAVY ? NBP_SAMPLE : c;
                int rank = 0;  // (a candidate whose rank reaches the number of free slots is not born: no need to finish the count)
                for (int j0 = 0; j0 < first && rank < nfree; j0 += 8) {
                    const int j1 = min(first, j0 + 8);
                    for (int j = j0; j < j1; ++j) rank += dp.cseg[b + j] < key;
                }
                if (rank < nfree) {
                    if (first == c) {
                        nb_place_born(mc, dp, cand, d, msk, rank);
                        ++born;
                    } else {
                        s_surv[atomicAdd(&s_nsurv, 1)] = pos;  // still in the race: ranked against the whole segment below
                    }
                }
            }
        }
        __syncthreads();
        const int ns = s_nsurv;
        for (int k = wid; k < ns; k += 8) {
            const int p2 = s_surv[k];
            const int cand = dp.csegi[p2];
            const int d = dp.Cdst[cand];
            const ulonglong2 msk = dp.MS[d];
            const int nfree = mask_free(mc, msk);
            const int b = dp.cbase[d], c = dp.ccnt[d], key = dp.cseg[p2];
            int rank = 0;
            for (int j0 = 0; j0 < c; j0 += 128) {  // four coalesced loads in flight per lane; the race is over once nfree keys are smaller
                int r4 = 0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 32 * u + lane;
                    r4 += j < c && dp.cseg[b + j] < key;
                }
                for (int sft = 16; sft > 0; sft >>= 1) r4 += __shfl_xor_sync(FULLMASK, r4, sft);
                rank += r4;
                if (rank >= nfree) break;
            }
            if (rank < nfree && lane == 0) {
                nb_place_born(mc, dp, cand, d, msk, rank);
                ++born;
            }
        }
        __syncthreads();
    }
    for (int sft = 16; sft > 0; sft >>= 1) born += __shfl_down_sync(FULLMASK, born, sft);
    if ((threadIdx.x & 31) == 0 && born) atomicAdd(&dp.st->n_born, born);
}
// ------------------------------------------------------------------------------------------------------------
// K7a list of occupied voxels (balances K7: occupied voxels are x-adjacent, so a strided sweep would hand one warp up
//     to 32 of them); empty voxels get their zero occupancy here.
// ------------------------------------------------------------------------------------------------------------
__global__ void k_voxel_list(MapConst mc, DevPtrs dp) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int base = warp * 32; base < mc.V; base += nwarps * 32) {
        const int v = base + lane;
        bool occ = false;
        if (v < mc.V) {
            const ulonglong2 m = dp.M[v];
            occ = (m.x | m.y) != 0ull;
            if (!occ) dp.OCCV[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const unsigned b = __ballot_sync(FULLMASK, occ);
        if (b == 0u) continue;
        int wb = 0;
        if (lane == 0) wb = atomicAdd(&dp.st->n_occ_voxels, __popc(b));
        wb = __shfl_sync(FULLMASK, wb, 0);
        if (occ) dp.E[wb + __popc(b & ((1u << lane) - 1u))] = v;
    }
}

// ------------------------------------------------------------------------------------------------------------
// K7  occupancy, future status and resampling (dsp_dynamic.h:924-1057).
//     A voxel's sums (fp32, slot order) and its systematic-resampling state machine are serial over ~20 particles, and
//     there are only a few thousand occupied voxels: the work is latency, not throughput.  A warp takes RS_VPW voxels:
//       stage  (all lanes, one voxel after the other) lane l owns slots l, l+32, ...: coalesced loads, the low-weight test
//              (:941), the future-status scatter of the old particles (:950-964); the kept particles go to the warp's
//              shared tile at their rank in slot order: (vx, vy, vz, w), slot, "old" and "flag is not 1" bits;
//       walk   (lane j = voxel j) both order-dependent loops run over the tile with plain shared-memory loads whose
//              addresses do not depend on the running sums; the verdicts (new weight, or removed) and the duplicates to make
//              (:1021-1044: the first free slot at that moment, else the weight piles up on the original) go back through
//              shared memory;
//       apply  (all lanes) survivors get their weight and flag 1, duplicates are copied from their source slot.
//     (Round 1's kernel gave a whole warp to one voxel and fetched every operand of the walk with find-first-set +
//     shuffle: 2 450 warp instructions per voxel, 35 us at cfg2.)
// ------------------------------------------------------------------------------------------------------------
#ifndef RS_VPW
#define RS_VPW 2     // voxels per warp (measured on B200, cfg2: 8 -> 83 us, 4 -> 53, 2 -> 38, 1 -> 40; profiles/r02_variants.jsonl)
#endif
#ifndef RS_WARPS
#define RS_WARPS 8   // warps per block
#endif
// shared memory of one warp: tile[RS_VPW][S + 1] float4 | verd[RS_VPW][S] float | tagb, dsrc, ddst [RS_VPW][S] u8 | wafter[RS_VPW] float
__host__ __device__ __forceinline__ size_t rs_warp_bytes(int S) {
    return ((size_t)RS_VPW * (S + 1) * 16 + (size_t)RS_VPW * S * 4 + (size_t)RS_VPW * S * 3 + (size_t)RS_VPW * 4 + 15) & ~(size_t)15;
}
#ifndef RS_MINB
#define RS_MINB 5   // blocks per SM the register allocation has to allow (92 -> 51 registers: 40 instead of 52 us at cfg2)
#endif
__global__ void __launch_bounds__(32 * RS_WARPS, RS_MINB) k_resample(MapConst mc, FrameConst fc, DevPtrs dp) {
    pdl_enter();
    extern __shared__ __align__(16) float4 rs_smem4[];
    unsigned char *rs_smem = reinterpret_cast<unsigned char *>(rs_smem4);
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int S = mc.S, R = (S + 31) >> 5, ROW = S + 1;  // (+1: the rows of the walking lanes start in different banks)
    unsigned char *base = rs_smem + (size_t)wl * rs_warp_bytes(S);
    float4 *tile = reinterpret_cast<float4 *>(base);
    float *verd = reinterpret_cast<float *>(base + (size_t)RS_VPW * ROW * 16);
    unsigned char *tagb = reinterpret_cast<unsigned char *>(verd + RS_VPW * S);  // slot | 0x80 if the flag is not 1 yet
    unsigned char *dsrc = tagb + RS_VPW * S, *ddst = dsrc + RS_VPW * S;
    float *wafter = reinterpret_cast<float *>(ddst + RS_VPW * S);
    const unsigned below = (1u << lane) - 1u;
    int c_pre = 0, c_old = 0, c_out = 0, c_low = 0;
    const int nocc = dp.st->n_occ_voxels;
    for (int item0 = warp * RS_VPW; item0 < nocc; item0 += nwarps * RS_VPW) {
        const int nv = min(RS_VPW, nocc - item0);
        // lane j < nv fetches voxel j's id and mask; the staging loop reads them by shuffle
        int my_v = 0;
        ulonglong2 my_m = make_ulonglong2(0ull, 0ull);
        if (lane < nv) {
            my_v = dp.E[item0 + lane];  // E carries the occupied-voxel list (k_voxel_list)
            my_m = dp.M[my_v];
        }
        int my_n = 0, my_ndup = 0;
        bool my_res = false;
        ulonglong2 my_occ = make_ulonglong2(0ull, 0ull), my_old = my_occ;  // kept slots; "old" bits by RANK
        __syncwarp();  // the previous group's readers are done with the tile
        // ---- stage
#pragma unroll 2
        for (int j = 0; j < nv; ++j) {
            const int v = __shfl_sync(FULLMASK, my_v, j);
            const u64 mx = __shfl_sync(FULLMASK, my_m.x, j), my = __shfl_sync(FULLMASK, my_m.y, j);
            int n = 0, n_low = 0;
            u64 kx = 0ull, ky = 0ull, ox = 0ull, oy = 0ull;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (r >= R) break;
                const unsigned live = r < 2 ? (unsigned)(mx >> (32 * r)) : (unsigned)(my >> (32 * (r - 2)));
                if (live == 0u) continue;
                float4 A = make_float4(0.f, 0.f, 0.f, 0.f), B = A;
                const bool mine = (live >> lane) & 1u;
                if (mine) {
                    const int a = v * S + 32 * r + lane;
                    A = dp.PA[a];
                    B = dp.PB[a];
                }
                const bool kp = mine && !((double)A.w < 1e-3);  // (:941) particles below 1e-3 are dropped
                const bool od = kp && B.w < 10.f;               // not newborn (:944)
                const unsigned keep = __ballot_sync(FULLMASK, kp), old = __ballot_sync(FULLMASK, od);
                if (kp) {
                    const int i = n + __popc(keep & below);
                    tile[j * ROW + i] = make_float4(B.x, B.y, B.z, A.w);
                    tagb[j * S + i] = (unsigned char)((32 * r + lane) | (B.w != 1.f ? 0x80 : 0));
                    if (od) {  // future status of the old particles (:950-964): each lane scatters its own
                        for (int t = 0; t < mc.T; ++t) {
                            const float ft = mc.ft[t];
                            const float fx = A.x + B.x * ft, fy = A.y + B.y * ft, fz = A.z + B.z * ft;
                            const int fi = dsp_voxel_index(mc, fx, fy, fz);
                            if (fi >= 0) atomicAdd(&dp.FUT[(size_t)fi * mc.T + t], A.w);
                        }
                        // this particle's "old" bit at its rank, collected by a second ballot below
                    }
                }
                // "old" bits in RANK order: lane i of the row's kept particles -> bit n + rank.  Every lane computes its own
                // bit position; an OR-reduction over the warp assembles the 128-bit mask.
                u64 bx = 0ull, by = 0ull;
                if (od) {
                    const int i = n + __popc(keep & below);
                    if (i < 64) bx = 1ull << i; else by = 1ull << (i - 64);
                }
                if (old) {  // (uniform)
                    bx |= __shfl_xor_sync(FULLMASK, bx, 16); by |= __shfl_xor_sync(FULLMASK, by, 16);
                    bx |= __shfl_xor_sync(FULLMASK, bx, 8);  by |= __shfl_xor_sync(FULLMASK, by, 8);
                    bx |= __shfl_xor_sync(FULLMASK, bx, 4);  by |= __shfl_xor_sync(FULLMASK, by, 4);
                    bx |= __shfl_xor_sync(FULLMASK, bx, 2);  by |= __shfl_xor_sync(FULLMASK, by, 2);
                    bx |= __shfl_xor_sync(FULLMASK, bx, 1);  by |= __shfl_xor_sync(FULLMASK, by, 1);
                    ox |= bx;
                    oy |= by;
                }
                if (r < 2) kx |= (u64)keep << (32 * r); else ky |= (u64)keep << (32 * (r - 2));
                n += __popc(keep);
                n_low += __popc(live) - __popc(keep);
            }
            if (lane == j) {
                my_n = n;
                my_occ = make_ulonglong2(kx, ky);
                my_old = make_ulonglong2(ox, oy);
                c_low += n_low;
            }
        }
        __syncwarp();
        // ---- walk: lane j < nv owns voxel j
        if (lane < nv) {
            const int n = my_n;
            const float4 *row = tile + lane * ROW;
            // sums in slot order (:938-973)
            float wsum = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
            for (int i = 0; i < n; ++i) {
                const float4 p = row[i];
                const bool od = i < 64 ? (my_old.x >> i) & 1ull : (my_old.y >> (i - 64)) & 1ull;
                if (od) { sx += p.x; sy += p.y; sz += p.z; }
                wsum += p.w;
            }
            const int n_old = __popcll(my_old.x) + __popcll(my_old.y);
            float4 o = make_float4(wsum, 0.f, 0.f, 0.f);
            if (n_old > 0) { o.y = sx / (float)n_old; o.z = sy / (float)n_old; o.w = sz / (float)n_old; }
            dp.OCCV[my_v] = o;
            ulonglong2 occ = my_occ;
            if (n >= 5) {  // (:986) systematic resampling to at most MAX particles, offset 0.5 * w_after, slot order
                my_res = true;
                const int n_after = n > mc.max_ppv ? mc.max_ppv : n;
                const float w_after = wsum / (float)n_after;
                wafter[lane] = w_after;
                float acc_ori = 0.f, acc_new = w_after * 0.5f;
                for (int i = 0; i < n; ++i) {
                    acc_ori += row[i].w;
                    if (acc_ori > acc_new) {
                        float wk = w_after;
                        acc_new += w_after;
                        bool full = false;
                        while (acc_ori > acc_new) {  // duplicate heavy particles into the first free slot (:1021-1044)
                            const int fs = full ? -1 : mask_nth_free(mc, occ, 0);
                            if (fs >= 0) {
                                dsrc[lane * S + my_ndup] = (unsigned char)i;
                                ddst[lane * S + my_ndup] = (unsigned char)fs;
                                ++my_ndup;
                                if (fs < 64) occ.x |= 1ull << fs; else occ.y |= 1ull << (fs - 64);
                            } else {
                                wk += w_after;
                                full = true;
                            }
                            acc_new += w_after;
                        }
                        verd[lane * S + i] = wk;
                    } else {  // removed (:1046-1050)
                        const int sl = tagb[lane * S + i] & 0x7f;
                        if (sl < 64) occ.x &= ~(1ull << sl); else occ.y &= ~(1ull << (sl - 64));
                        verd[lane * S + i] = -1.f;
                    }
                }
            }
            dp.M[my_v] = occ;
            c_pre += n;
            c_old += n_old;
            c_out += mask_popc(occ);
        }
        __syncwarp();
        // ---- apply
        for (int j = 0; j < nv; ++j) {
            const int v = __shfl_sync(FULLMASK, my_v, j), n = __shfl_sync(FULLMASK, my_n, j), nd = __shfl_sync(FULLMASK, my_ndup, j);
            const bool res = __shfl_sync(FULLMASK, (int)my_res, j) != 0;
            for (int i = lane; i < n; i += 32) {
                const float4 p = tile[j * ROW + i];
                const unsigned char tb = tagb[j * S + i];
                const float nw = res ? verd[j * S + i] : p.w;
                if (!(nw < 0.f)) {
                    const int a = v * S + (tb & 0x7f);
                    if (nw != p.w) dp.PA[a].w = nw;
                    if (tb & 0x80) dp.PB[a].w = 1.f;  // newborn / moved flags become "valid" (:968)
                }
            }
            for (int d = lane; d < nd; d += 32) {
                const int i = dsrc[j * S + d], fs = ddst[j * S + d];
                const float4 p = tile[j * ROW + i];
                const float4 A = dp.PA[v * S + (tagb[j * S + i] & 0x7f)];  // position of the original (its weight is being rewritten: not used)
                dp.PA[v * S + fs] = make_float4(A.x, A.y, A.z, wafter[j]);
                dp.PB[v * S + fs] = make_float4(p.x, p.y, p.z, 0.6f);
            }
        }
    }
    // the walking lanes hold the counts: one atomic per counter and BLOCK (same-address atomics serialise in L2; a few
    // thousand of them per counter were the longest part of this kernel)
    __shared__ int s_cnt[4];
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int d = 16; d > 0; d >>= 1) {
        c_pre += __shfl_down_sync(FULLMASK, c_pre, d);
        c_old += __shfl_down_sync(FULLMASK, c_old, d);
        c_out += __shfl_down_sync(FULLMASK, c_out, d);
        c_low += __shfl_down_sync(FULLMASK, c_low, d);
    }
    if (lane == 0 && (c_pre | c_low)) {
        atomicAdd(&s_cnt[0], c_pre);
        atomicAdd(&s_cnt[1], c_old);
        atomicAdd(&s_cnt[2], c_out);
        atomicAdd(&s_cnt[3], c_low);
    }
    __syncthreads();
    if (threadIdx.x == 0 && (s_cnt[0] | s_cnt[3])) {
        atomicAdd(&dp.st->n_pre, s_cnt[0]);
        atomicAdd(&dp.st->n_old, s_cnt[1]);
        atomicAdd(&dp.st->n_out, s_cnt[2]);
        if (s_cnt[3]) atomicAdd(&dp.st->n_low_weight, s_cnt[3]);
    }
}

// end of frame: reset the arrival-grouping tables touched this frame; advance the noise cursors by what the reference's
// serial newborn loop would have drawn (dsp_dynamic.h:1162-1178); flag a frame no observation kernel handled
__global__ void k_cleanup(MapConst mc, FrameConst fc, DevPtrs dp, int newborn_ran, int fallback_launched) {
    pdl_enter();
    int n1 = dp.st->n_mov_owner, n2 = dp.st->n_cand_owner;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += gridDim.x * blockDim.x) {
        if (i < n1) { int d = dp.mowner[i]; dp.mcnt[d] = 0; dp.mfill[d] = 0; }
        else { int d = dp.cowner[i - n1]; dp.ccnt[d] = 0; dp.cfill[d] = 0; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        DevState *s = dp.st;
        if (newborn_ran) {
            const int inmap_pts = dp.nrank[fc.n_tagged], vd = dp.nvoff[fc.n_tagged], rd = dp.nroff[fc.n_tagged];
            s->n_inmap_points = inmap_pts;
            s->n_vdraw = vd;
            s->n_rdraw = rd;
            s->p_cur = (s->p_cur + 3ll * inmap_pts * fc.nb_num) % mc.G;
            s->v_cur = (s->v_cur + 3ll * vd) % mc.G;
            s->u_cur += 3ll * rd;
        }
        s->use_store = use_pair_buffer(mc, dp) ? 1 : 0;
        if (!s->use_store && !fallback_launched && fc.stage_limit >= 2) atomicOr(&s->overflow, 4);
    }
}

// ------------------------------------------------------------------------------------------------------------
// K8  readers (dsp_dynamic.h:385-438): ordered compaction of occupied voxel centres, future copy-out + zeroing
// ------------------------------------------------------------------------------------------------------------
#define OCC_BLOCK 512  // voxels per block
// fidx != nullptr: the non-zero voxel rows of the future grid are packed into (fidx, fval) as they are read — rows in no
// particular order, *d_nf counts them (zeroed by the caller) — instead of the dense copy into d_future: the sparse copy-out
// of the host reader without its three extra kernels.
__global__ void __launch_bounds__(256) k_occ_count(MapConst mc, DevPtrs dp, float thr, int *blockcnt, float *d_future, int *fidx, float *fval, int *d_nf) {
    pdl_enter();
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    int b = blockIdx.x * OCC_BLOCK, c = 0;
    for (int i = threadIdx.x; i < OCC_BLOCK; i += blockDim.x) {
        int v = b + i;
        if (v < mc.V && dp.OCCV[v].x > thr) ++c;
    }
    for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(FULLMASK, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s, c);
    __syncthreads();
    if (threadIdx.x == 0) blockcnt[blockIdx.x] = s;
    // future status: copy out (if asked) and clear (:416-424)
    if (fidx) {
        const int lane = threadIdx.x & 31;
        for (int i0 = 0; i0 < OCC_BLOCK; i0 += blockDim.x) {  // (uniform trip count: warp collectives inside)
            const int v = b + i0 + threadIdx.x;
            float row[DSP_MAX_T];
            bool nz = false;
            if (v < mc.V)
                for (int t = 0; t < mc.T; ++t) {
                    row[t] = dp.FUT[(size_t)v * mc.T + t];
                    nz |= row[t] != 0.f;
                }
            const unsigned bal = __ballot_sync(FULLMASK, nz);
            if (bal == 0u) continue;
            int base = 0;
            if (lane == 0) base = atomicAdd(d_nf, __popc(bal));
            base = __shfl_sync(FULLMASK, base, 0);
            if (nz) {
                const int pos = base + __popc(bal & ((1u << lane) - 1u));
                fidx[pos] = v;
                for (int t = 0; t < mc.T; ++t) {
                    fval[(size_t)pos * mc.T + t] = row[t];
                    dp.FUT[(size_t)v * mc.T + t] = 0.f;
                }
            }
        }
        return;
    }
    size_t fb = (size_t)b * mc.T, fe = min((size_t)mc.V, (size_t)b + OCC_BLOCK) * mc.T;
    for (size_t i = fb + threadIdx.x; i < fe; i += blockDim.x) {
        if (d_future) d_future[i] = dp.FUT[i];
        dp.FUT[i] = 0.f;
    }
}
__global__ void __launch_bounds__(256) k_occ_write(MapConst mc, DevPtrs dp, float thr, const int *blockcnt, float *xyz, int cap, int *d_count, int nblocks) {
    pdl_enter();
    __shared__ int wsum[8], s_before, s_total;
    // this block's offset = the occupied voxels of all blocks before it (k_occ_count's per-block counts): summed here instead
    // of by a scan launch between the two reader kernels
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    {
        int before = 0, total = 0;
        for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
            const int c = blockcnt[b];
            total += c;
            if (b < (int)blockIdx.x) before += c;
        }
        for (int d = 16; d > 0; d >>= 1) {
            before += __shfl_down_sync(FULLMASK, before, d);
            total += __shfl_down_sync(FULLMASK, total, d);
        }
        if (threadIdx.x == 0) { s_before = 0; s_total = 0; }
        __syncthreads();
        if (lane == 0) { atomicAdd(&s_before, before); atomicAdd(&s_total, total); }
        __syncthreads();
    }
    int b = blockIdx.x * OCC_BLOCK;
    int run = s_before;
    if (blockIdx.x == 0 && threadIdx.x == 0 && d_count) *d_count = s_total;
    for (int i0 = 0; i0 < OCC_BLOCK; i0 += 256) {
        int v = b + i0 + threadIdx.x;
        bool occ = v < mc.V && dp.OCCV[v].x > thr;
        unsigned bal = __ballot_sync(FULLMASK, occ);
        if (lane == 0) wsum[w] = __popc(bal);
        __syncthreads();
        int before = 0, tot = 0;
        for (int k = 0; k < 8; ++k) { int x = wsum[k]; if (k < w) before += x; tot += x; }
        if (occ) {
            int pos = run + before + __popc(bal & ((1u << lane) - 1u));
            if (pos < cap) {
                float c[3];
                dsp_voxel_center(mc, v, c);
                xyz[3 * pos] = c[0]; xyz[3 * pos + 1] = c[1]; xyz[3 * pos + 2] = c[2];
            }
        }
        run += tot;
        __syncthreads();
    }
}
// Sparse copy-out of the future-status grid (experiment switch DSPMAP_SPARSE_FUTURE=1).  The reference hands the application
// V x T floats per call (dsp_dynamic.h:416-418), 4.2 MB at cfg2, of which 2.4 % of the voxel rows are non-zero (measured on
// the reference's state): over PCIe only those rows travel, as (voxel id, T values) records in ascending voxel order, and the
// host patches the application's array (dspmap_get_occupancy).  Both kernels read the dense device copy k_occ_count made.
__global__ void k_future_clear(MapConst mc, DevPtrs dp) {
    pdl_enter();
    size_t n = (size_t)mc.V * mc.T;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dp.FUT[i] = 0.f;
}

// ------------------------------------------------------------------------------------------------------------
// state dump / load
// ------------------------------------------------------------------------------------------------------------
__global__ void k_dump_gather(MapConst mc, DevPtrs dp, int *keys, float *vals) {
    pdl_enter();
    const int n = min(dp.st->n_live, dp.cap_live);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int key = dp.E[i];
        int a = (key >> DSP_KEY_SHIFT) * mc.S + (key & (DSP_MAX_SLOTS - 1));
        float4 A = dp.PA[a], B = dp.PB[a];
        keys[i] = key;
        float *o = vals + 8 * (size_t)i;
        o[0] = B.w; o[1] = B.x; o[2] = B.y; o[3] = B.z; o[4] = A.x; o[5] = A.y; o[6] = A.z; o[7] = A.w;
    }
}
__global__ void k_load_scatter(MapConst mc, DevPtrs dp, const int *ids, const float *vals, int n) {
    pdl_enter();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int v = ids[2 * i], s = ids[2 * i + 1];
        const float *o = vals + 8 * (size_t)i;
        int a = v * mc.S + s;
        dp.PA[a] = make_float4(o[4], o[5], o[6], o[7]);
        dp.PB[a] = make_float4(o[1], o[2], o[3], o[0]);
        mask_atomic_set(dp.M, v, s);
    }
}
__global__ void k_reset_counter(int *p) {
    pdl_enter();
    *p = 0;
}
