// dspmap_kernels.cuh — sm_100a kernels of the DSP map's per-frame particle loop (product code).
//
// Compiled with -fmad=false, IEEE division and square root, no flush-to-zero: every fp32 operation below
// rounds once, in the order the reference (g-ch/DSP-map include/dsp_dynamic.h) writes it, so particle
// positions, weights, voxel ids, pyramid ids and slot ids come out bit-identical to the reference CPU loop.
// The serial "first free slot" / "append to list" semantics of the reference are reproduced by ordering
// everything on the sweep key (voxel * 128 + slot), see DESIGN.md.
#pragma once
#include <cuda_runtime.h>
#include "dspmap_types.h"

#define FULLMASK 0xffffffffu

// Programmatic dependent launch (griddepcontrol, sm_90+).  When the host launches the frame's kernels with
// cudaLaunchAttributeProgrammaticStreamSerialization (dspmap.cu: launch_kernel, DSPMAP_PDL=1), kernel k+1 may be scheduled
// while kernel k still runs; nothing the predecessor writes may be touched before pdl_wait(), which returns once every
// prerequisite grid has completed and its memory is visible.  EVERY kernel executes the wait before it exits (also on its
// early-return paths): the guarantee is transitive only through kernels that waited.  Launched without the attribute both
// instructions are no-ops.
__device__ __forceinline__ void pdl_trigger() {
#ifdef __CUDA_ARCH__  // (the warp-level kernels are also compiled for the host by tests/simt: no PTX there)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() {
#ifdef __CUDA_ARCH__
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_enter() {
    pdl_trigger();
    pdl_wait();
}

struct DevPtrs {
    // particle store, dense direct-addressed: slot address a = voxel * S + slot
    float4 *PA;       // px py pz weight
    float4 *PB;       // vx vy vz flag
    ulonglong2 *M;    // occupancy mask per voxel (working)
    ulonglong2 *M0;   // occupancy mask at frame start
    ulonglong2 *MS;   // sparse snapshots taken by the arrival-grouping owner pass
    float4 *OCCV;     // per voxel: weight sum, mean vx, vy, vz   (voxels_objects_number[.][0..3])
    float *FUT;       // per voxel x horizon future weight sums    (voxels_objects_number[.][4..])
    int *E;           // live particle keys of this frame
    int *vzcnt, *vzoff, *vzblk, *vzblkoff;  // ordered prediction-noise ranks (vz mode only)
    // observations
    const float *pts;   // n x 3, sensor frame
    float4 *OR;         // rotated point + range
    int *OPID;          // pyramid id or -1
    int *obs_cnt, *obs_fill, *obs_maxbits, *obs_off, *obs_capoff, *OSEG;
    float4 *OBSP;       // [P][OBS] x y z range
    float *CZ;          // [P][OBS] C_z
    float *INV;         // 1 / C_z, dense in (pyramid, bin order)
    // movers (particles that change voxel in prediction)
    float4 *MBA, *MBB;
    int *MBkey, *MBdst, *MBq;
    int *mcnt, *mfill, *mbase, *mowner, *mseg;
    // particles registered in FOV pyramids
    int *Fkey, *Faddr, *Fq;
    float4 *FP;         // position + weight of the registered particle (payload for the sharded all-gather)
    float4 *PSpay;      // payload scattered with PSkey / PSaddr (sharded mode)
    // sharded mode exchange buffers (caller-owned device memory, see dspmap_shard_config)
    float *xsend, *xrecv;  // [nranks][4 + cap_x * 12]: header {count}, then boundary-crossing movers
    float *gsend, *grecv;  // gsend [4 + cap_g * 8], grecv [nranks][4 + cap_g * 8]: registered particles
    float *nst_shared;     // [max_points] per-point static newborn count (as float), summed over ranks
    float *NW;             // [nranks * cap_g] new weights by global list index, summed over ranks (one writer each)
    int *pcount, *pfill, *poff, *plen;
    int *pub;           // per pyramid: movers heading for it this frame (upper bound for the overflow test of k_arrive)
    int *rkey;          // replay scratch: sweep keys of the events to walk
    int *rkey2, *rpos;  // replay scratch: keys in sweep order; an event's place among the events of its voxel
    int *rcount;        // replay scratch: events per voxel (V + 1, zero between uses)
    int *PSkey, *PSaddr;
    int *LA;            // per-pyramid sorted list: slot address
    float4 *LP;         // per-pyramid sorted list: px py pz weight (post-prediction)
    float *PW;          // per-pyramid sorted list: P_d * weight
    // pair buffer: G[rowbase[i] + z * totlen[i] + j] = g(particle j of pyramid i's concatenated neighbour lists; point z of i)
    float *G;
    int *cum, *totlen, *pairs, *rowbase, *chunks, *chunk_off;
    int *chunk_pyr;     // chunk -> pyramid (k_pair_prep)
    const int *nbrev;   // [P][NBW]: nbrev[a][1 + ns] = position of a in the neighbour list of nbr[a][1 + ns] (host-built)
    // newborn
    const float *tagged;  // n_tagged x 7, world frame
    float4 *NPC;          // corrected point + voxel id
    int *ninmap, *nrank, *nstatic, *nvcnt, *nrcnt, *nvoff, *nroff;
    u64 *nimask;
    float4 *CA;           // candidate position (w unused)
    int *Ckey, *Cdst, *Caddr;  // (point, candidate) key, destination voxel, slot address once born (-1 otherwise)
    int *ccnt, *cfill, *cbase, *cowner, *cseg, *csegi;
    // tables
    const float *ptab, *vtab, *lut;
    const float *planes0;  // boundary-plane normals, sensor frame: (Nh+1) + (Nv+1) vectors
    float *planes;         // rotated
    const int *nbr;        // [P][NBW]
    DevState *st;
    int cap_live, cap_cand;
};

// ------------------------------------------------------------------------------------------------------------
// bit-exact scalar helpers
// ------------------------------------------------------------------------------------------------------------

// rotateVectorByQuaternion (dsp_dynamic.h:1303-1322): q * (0, v) * q^-1 as two scalar Hamilton products.
__host__ __device__ __forceinline__ void dsp_rotate(const float *v, const float *q, const float *qi, float *o) {
    const float aw = q[0], ax = q[1], ay = q[2], az = q[3];
    const float bw = 0.f, bx = v[0], by = v[1], bz = v[2];
    const float tw = aw * bw - ax * bx - ay * by - az * bz;
    const float tx = aw * bx + ax * bw + ay * bz - az * by;
    const float ty = aw * by + ay * bw + az * bx - ax * bz;
    const float tz = aw * bz + az * bw + ax * by - ay * bx;
    const float iw = qi[0], ix = qi[1], iy = qi[2], iz = qi[3];
    o[0] = tw * ix + tx * iw + ty * iz - tz * iy;
    o[1] = tw * iy + ty * iw + tz * ix - tx * iz;
    o[2] = tw * iz + tz * iw + tx * iy - ty * ix;
}
__host__ __device__ __forceinline__ void dsp_quat_inverse(const float *q, float *qi) {
    const float n2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3] + q[0] * q[0];
    if (n2 > 0.f) {
        qi[0] = q[0] / n2;
        qi[1] = -q[1] / n2;
        qi[2] = -q[2] / n2;
        qi[3] = -q[3] / n2;
    } else {
        qi[0] = qi[1] = qi[2] = qi[3] = 0.f;
    }
}
// vectorMultiply (dsp_dynamic.h:1324-1326)
__device__ __forceinline__ float dsp_dot(const float *n, float x, float y, float z) { return x * n[0] + y * n[1] + z * n[2]; }
// ifInPyramidsArea (:1329-1339); ph/pv = rotated plane normals
__device__ __forceinline__ bool dsp_in_fov(const float *ph, const float *pv, int Nh, int Nv, float x, float y, float z) {
    return dsp_dot(ph, x, y, z) >= 0.f && dsp_dot(ph + 3 * Nh, x, y, z) <= 0.f && dsp_dot(pv, x, y, z) <= 0.f &&
           dsp_dot(pv + 3 * Nv, x, y, z) >= 0.f;
}
// findPointPyramidHorizontalIndex / VerticalIndex (:1341-1367): first sign change, seeds +1 / -1
__device__ __forceinline__ int dsp_pyr_scan(const float *pl, int n, float seed, float x, float y, float z) {
    float last = seed;
    for (int i = 0; i < n; i++) {
        float d = dsp_dot(pl + 3 * (i + 1), x, y, z);
        if (last * d <= 0.f) return i;
        last = d;
    }
    return -1;
}
// Correctly rounded a / b for a divisor known in advance: r = RN(1/b), one product and two fused corrections.
// It is used ONLY after k_verify_div has compared it bit for bit with IEEE division over every float of the reachable
// argument range for that particular divisor (dspmap.cu: verify_fast_div); otherwise the callers divide.
__host__ __device__ __forceinline__ float dsp_div_known(float a, float b, float r) {
#ifdef __CUDA_ARCH__
    const float q0 = a * r;
    const float e = __fmaf_rn(-q0, b, a);
    return __fmaf_rn(e, r, q0);
#else
    (void)r;
    return a / b;
#endif
}
// getParticleVoxelsIndex (:1076-1088) with ifParticleIsOut (:1118-1125); -1 when outside
__host__ __device__ __forceinline__ int dsp_voxel_index(const MapConst &mc, float x, float y, float z) {
    if (x >= mc.hx || x <= -mc.hx || y >= mc.hy || y <= -mc.hy || z >= mc.hz || z <= -mc.hz) return -1;
    int ix, iy, iz;
    if (mc.fast_res) {
        ix = (int)dsp_div_known(x + mc.hx, mc.res, mc.res_r);
        iy = (int)dsp_div_known(y + mc.hy, mc.res, mc.res_r);
        iz = (int)dsp_div_known(z + mc.hz, mc.res, mc.res_r);
    } else {
        ix = (int)((x + mc.hx) / mc.res);
        iy = (int)((y + mc.hy) / mc.res);
        iz = (int)((z + mc.hz) / mc.res);
    }
    int idx = iz * mc.ny * mc.nx + iy * mc.nx + ix;
    if (idx < 0 || idx >= mc.V) return -1;
    return idx;
}
// getVoxelPositionFromIndex (:1090-1107)
__host__ __device__ __forceinline__ void dsp_voxel_center(const MapConst &mc, int idx, float *o) {
    int zs = mc.ny * mc.nx, iz = idx / zs, rem = idx - iz * zs, iy = rem / mc.nx, ix = rem - iy * mc.nx;
    float cx = -mc.hx + mc.res * 0.5f, cy = -mc.hy + mc.res * 0.5f, cz = -mc.hz + mc.res * 0.5f;
    o[0] = (float)ix * mc.res + cx;
    o[1] = (float)iy * mc.res + cy;
    o[2] = (float)iz * mc.res + cz;
}
// queryNormalPDF (:1294-1301) against the half table: lut[|i - 10000|] == standard_gaussian_pdf[i]
__device__ __forceinline__ float dsp_pdf(const float *lut, float x, float mu, float sigma) {
    float cx = (x - mu) / sigma;
    if (cx > 9.9f) cx = 9.9f;
    else if (cx < -9.9f) cx = -9.9f;
    int i = (int)(cx * 1000 + 10000) - 10000;
    return lut[i < 0 ? -i : i];
}
// same, with the exhaustively verified exact division by sigma
__device__ __forceinline__ float dsp_pdf_f(const float *lut, float x, float mu, const FrameConst &fc) {
    float cx = fc.fast_sigma ? dsp_div_known(x - mu, fc.sigma, fc.sigma_r) : (x - mu) / fc.sigma;
    if (cx > 9.9f) cx = 9.9f;
    else if (cx < -9.9f) cx = -9.9f;
    int i = (int)(cx * 1000 + 10000) - 10000;
    return lut[i < 0 ? -i : i];
}
// the counter-based uniform stream standing where rand() is (dsp_dynamic.h:1552)
__host__ __device__ __forceinline__ uint32_t dsp_u31(u64 seed, u64 k) {
    u64 z = seed + (k + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 33);
}
// generateRandomFloat (:1551-1553) with RAND_MAX = 2147483647
__host__ __device__ __forceinline__ float dsp_uniform(u64 seed, u64 k, float lo, float hi) {
    int r = (int)dsp_u31(seed, k);
    return lo + (float)r / ((float)(2147483647 / (hi - lo)));
}

// ------------------------------------------------------------------------------------------------------------
// mask helpers (bits >= S are never set in stored masks)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int mask_popc(ulonglong2 m) { return __popcll(m.x) + __popcll(m.y); }
__device__ __forceinline__ int mask_free(const MapConst &mc, ulonglong2 m) { return __popcll(~m.x & mc.vlo) + __popcll(~m.y & mc.vhi); }
// the first k free slots of m, as a mask
__device__ __forceinline__ ulonglong2 mask_take_free(const MapConst &mc, ulonglong2 m, int k) {
    ulonglong2 r = make_ulonglong2(0ull, 0ull);
    u64 z = ~m.x & mc.vlo;
    while (k > 0 && z) { u64 b = z & (0ull - z); r.x |= b; z ^= b; --k; }
    z = ~m.y & mc.vhi;
    while (k > 0 && z) { u64 b = z & (0ull - z); r.y |= b; z ^= b; --k; }
    return r;
}
// index of the n-th (0-based) free slot of m, or -1
__device__ __forceinline__ int mask_nth_free(const MapConst &mc, ulonglong2 m, int n) {
    u64 z = ~m.x & mc.vlo;
    int c = __popcll(z);
    if (n < c) {
        for (int i = 0; i < n; ++i) z &= z - 1;
        return __ffsll((long long)z) - 1;
    }
    n -= c;
    z = ~m.y & mc.vhi;
    if (n < __popcll(z)) {
        for (int i = 0; i < n; ++i) z &= z - 1;
        return 64 + __ffsll((long long)z) - 1;
    }
    return -1;
}
__device__ __forceinline__ bool mask_test(ulonglong2 m, int s) { return s < 64 ? (m.x >> s) & 1ull : (m.y >> (s - 64)) & 1ull; }
__device__ __forceinline__ void mask_atomic_clear(ulonglong2 *M, int v, int s) {
    u64 *w = reinterpret_cast<u64 *>(M + v) + (s >> 6);
    atomicAnd(w, ~(1ull << (s & 63)));
}
__device__ __forceinline__ void mask_atomic_set(ulonglong2 *M, int v, int s) {
    u64 *w = reinterpret_cast<u64 *>(M + v) + (s >> 6);
    atomicOr(w, 1ull << (s & 63));
}

// owner rank of a voxel under z-slab sharding
__host__ __device__ __forceinline__ int dsp_owner(const MapConst &mc, int voxel) {
    int r = (voxel / (mc.nx * mc.ny)) / mc.z_per_rank;
    return r < mc.nranks ? r : mc.nranks - 1;
}
#define XREC 12  // words per boundary-crosser record: A(4) B(4) key dst q pad
#define GREC 8   // words per registered-particle record: key q addr pad px py pz w
#define SLAB_HDR 4

// warp-aggregated counter increment; returns this thread's index
__device__ __forceinline__ int agg_inc(int *ctr) {
    unsigned m = __activemask();
    int lane = threadIdx.x & 31;
    int leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(ctr, __popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

__device__ __forceinline__ void load_planes(float *sp, const DevPtrs &dp, const MapConst &mc) {
    int n = 3 * (mc.Nh + 1 + mc.Nv + 1);
    for (int i = threadIdx.x; i < n; i += blockDim.x) sp[i] = dp.planes[i];
    __syncthreads();
}
#define DSP_MAX_PLANES 600  // (Nh + 1 + Nv + 1) vectors, shared-memory staging
