// dspmap_estimator.cuh — the reference's side thread on the device (product code; SURVEY.md §8f row 2, "GPU connected
// components version").  Replaces the per-frame front half of DSPMap::velocityEstimationThread (g-ch/DSP-map
// include/dsp_dynamic.h:1377-1447 and :1503-1540; static variant dsp_static.h:1285-1309): FOV filter of the rotated cloud,
// ground / non-ground split, Euclidean clustering of the non-ground points (pcl::EuclideanClusterExtraction: tolerance
// 2 x filter resolution, sizes 5..10000, member indices ascending, clusters by size descending, ties by first member),
// cluster centroids, static / dynamic classification, and the layout of the velocity-tagged cloud the newborn step reads
// (points of the dynamic clusters cluster by cluster, then the ground points, then the points of the static clusters).
// What stays on the host is the Hungarian matching of a few dozen cluster centroids against the previous frame's
// (:1449-1499, VelocityEstimator::match): the kernels hand it the centroids through page-locked memory and
// k_est_apply writes the velocities it returns into the tagged cloud.
//
// Results are those of velocity_estimator.cpp (the host implementation, which tests/test_host.py pins against the
// reference's own thread) bit for bit: components of the graph "d^2 <= tol^2" do not depend on how they are found, and every
// floating-point expression (rotation, world shift, squared distance, centroid sums in ascending index order) is evaluated
// in the order the host code writes it (-fmad=false).
//
// The tagged cloud is padded: the host does not know how many points survive (points of clusters below 5 or above 10000
// members are dropped) when it sizes the newborn kernels, so it launches them for n_pad >= n points and the entries behind
// the real ones are placed far outside the map (k_nb_point0 skips them: FrameConst::tagged_padded).
#pragma once
#include <climits>
#include "dspmap_kernels.cuh"

#define EST_FAR 1e30f          // coordinate of a padding entry
#define EST_MIN_CLUSTER 5      // :1413
#define EST_MAX_CLUSTER 10000  // :1414
#define EST_DYN_MAX_POINTS 200 // DYNAMIC_CLUSTER_MAX_POINT_NUM (:52)
#define EST_DYN_MAX_HEIGHT 1.5f  // DYNAMIC_CLUSTER_MAX_CENTER_HEIGHT (:53)

struct EstConst {
    float nrm[12];  // the four outer FOV planes, rotated (h first, h last, v first, v last)
    float q[4], qi[4], cur[3];
    float filter_res, tol2, inv_cell;
    int n, n_pad_prev, n_pad, model;
    int nt_override;  // >= 0: the cloud on the device was supplied by the caller and has this many entries (kept when nothing is in view)
    unsigned hash_mask;
};
// counters of one frame (device) and what the host reads (page-locked, mapped)
enum {
    EC_NV = 0, EC_NC, EC_NGROUND, EC_NDYN, EC_NDYNPTS, EC_NSTATICPTS, EC_NCELLS, EC_TICKET, EC_TICKET2,  // zeroed in front of every frame
    EC_NTAGGED,  // persists: the cloud is kept when nothing is in view
    EC_COUNT = 12
};
#define EC_PER_FRAME EC_NTAGGED
#define EC_CURSOR EC_NGROUND  // (scratch until k_est_clusters sets the ground count: allocation cursor of the cell-sorted point array)
struct EstPtrs {
    const float *pts;  // n x 3, sensor frame
    float4 *W;         // world-frame point, w = class (0 not in view, 1 ground / static, 2 non-ground)
    int *parent, *csize, *label, *pos, *grank, *mrank, *kidx, *flag_g, *flag_r;
    u64 *hkey;         // hash grid: packed cell -> slot
    int *hcnt;         // per slot: points of the cell - 1 (the frame's memset leaves -1)
    unsigned *cells;   // slots of the occupied cells, in order of creation (a cell's number is its position here)
    int *cellid;       // slot -> cell number
    int *cellof;       // point -> slot of its cell
    int *cmin, *ccount, *rootc;  // per cell: smallest point index / points of the component it is the root of; per point: root cell
    int *cbase, *ccells;  // per cell: start and length of its points in SW
    int *bbox;         // per cell: bounding box of its points, 6 order-preserving ints (min xyz, max xyz)
    float4 *SW;        // the non-ground points sorted by cell (x y z, w = point index)
    int *croot, *csz, *spos, *cdyn, *coff, *dseq, *order, *a_dyn, *a_dsz, *a_ssz, *p_dyn, *p_dsz, *p_ssz;  // per cluster
    float4 *cfeat;
    int *cnt;          // EC_* (device); EC_NTAGGED persists across frames
    int *h_hdr;        // EC_* (host, mapped)
    EstFeature *h_feat;  // dynamic clusters in sorted order (host, mapped)
    float *tagged;     // n_pad x 7
    int *tcid;         // per tagged entry: its dynamic cluster (index into h_feat / cvel) or -1
    const float4 *cvel;  // per dynamic cluster: velocity + intensity, written by the host after the matching
};

__device__ __forceinline__ int est_cell(float v, float inv_cell) {
    const float f = floorf(v * inv_cell);
    return (int)fminf(fmaxf(f, -1048575.f), 1048575.f);
}
__device__ __forceinline__ u64 est_pack(int a, int b, int c) {
    return ((u64)((a + (1 << 20)) & 0x1FFFFF) << 42) | ((u64)((b + (1 << 20)) & 0x1FFFFF) << 21) | (u64)((c + (1 << 20)) & 0x1FFFFF);
}
__device__ __forceinline__ unsigned est_hash(u64 k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    return (unsigned)k;
}

// rotation, FOV test, world shift, ground split (:226-257, :1387-1398); union-find and hash-grid set-up
__global__ void k_est_classify(EstConst ec, EstPtrs ep) {
    pdl_enter();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ec.n; i += gridDim.x * blockDim.x) {
        const float v[3] = {ep.pts[3 * i], ep.pts[3 * i + 1], ep.pts[3 * i + 2]};
        float r[3];
        dsp_rotate(v, ec.q, ec.qi, r);
        const float d0 = r[0] * ec.nrm[0] + r[1] * ec.nrm[1] + r[2] * ec.nrm[2];
        const float d1 = r[0] * ec.nrm[3] + r[1] * ec.nrm[4] + r[2] * ec.nrm[5];
        const float d2 = r[0] * ec.nrm[6] + r[1] * ec.nrm[7] + r[2] * ec.nrm[8];
        const float d3 = r[0] * ec.nrm[9] + r[1] * ec.nrm[10] + r[2] * ec.nrm[11];
        const bool in = d0 >= 0.f && d1 <= 0.f && d2 <= 0.f && d3 >= 0.f;
        const float wx = r[0] + ec.cur[0], wy = r[1] + ec.cur[1], wz = r[2] + ec.cur[2];
        int cls = in ? 1 : 0;
        if (in && ec.model != 1 && wz > ec.filter_res) cls = 2;
        ep.W[i] = make_float4(wx, wy, wz, __int_as_float(cls));
        ep.parent[i] = i;  // (cell i, if it comes to exist: there are at most as many cells as points)
        ep.cmin[i] = INT_MAX;
        ep.ccount[i] = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) { ep.bbox[6 * i + k] = INT_MAX; ep.bbox[6 * i + 3 + k] = INT_MIN; }
        ep.csize[i] = 0;
        ep.flag_g[i] = cls == 1;
        if (cls == 2) {
            const u64 key = est_pack(est_cell(wx, ec.inv_cell), est_cell(wy, ec.inv_cell), est_cell(wz, ec.inv_cell));
            unsigned h = est_hash(key) & ec.hash_mask;
            for (;;) {
                const u64 old = atomicCAS(&ep.hkey[h], ~0ull, key);
                if (old == ~0ull) {  // this thread created the cell
                    const int c = atomicAdd(&ep.cnt[EC_NCELLS], 1);
                    ep.cells[c] = h;
                    ep.cellid[h] = c;
                }
                if (old == ~0ull || old == key) break;
                h = (h + 1) & ec.hash_mask;
            }
            ep.cellof[i] = (int)h;
            ep.pos[i] = atomicAdd(&ep.hcnt[h], 1) + 1;
        }
    }
    // the last block to finish lets every cell reserve its stretch of the cell-sorted point array (in no particular order)
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ep.cnt[EC_TICKET], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    volatile int *vc = ep.cnt;
    const int ncells = vc[EC_NCELLS];
    for (int c = threadIdx.x; c < ncells; c += blockDim.x) {
        const int m = ((volatile int *)ep.hcnt)[((volatile unsigned *)ep.cells)[c]] + 1;
        ep.ccells[c] = m;
        ep.cbase[c] = atomicAdd(&ep.cnt[EC_CURSOR], m);
    }
}

// Union-find over the occupied CELLS (a cell's points are mutually linked, see k_est_link).  A root is hooked under the root
// with the smaller pseudo-random priority: with raster-ordered clouds, linking by index builds chains as long as the component
// (measured: 0.6 s for a 10 k-point wall); random linking keeps the trees shallow.  Roots carry no meaning: a component is
// named by its smallest POINT index afterwards (k_est_clusters).
__device__ __forceinline__ unsigned est_prio(int c) { return (unsigned)c * 0x9E3779B1u; }  // a bijection: no ties
// root of c's tree, halving the path on the way (a non-root only ever receives another ancestor as its parent: concurrent
// halving stores and hooks commute, and the hook below only ever succeeds on a root)
__device__ __forceinline__ int est_find(int *parent, int i) {
    volatile int *vp = parent;
    for (;;) {
        const int p = vp[i];
        if (p == i) return i;
        const int g = vp[p];
        if (g != p) vp[i] = g;
        i = g;
    }
}
__device__ __forceinline__ void est_union(int *parent, int a, int b) {
    for (;;) {
        a = est_find(parent, a);
        b = est_find(parent, b);
        if (a == b) return;
        if (est_prio(a) < est_prio(b)) { const int t = a; a = b; b = t; }
        if (atomicCAS(&parent[a], a, b) == a) return;
    }
}
__device__ __forceinline__ int est_ordered(float f) {  // int order == float order
    const int b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7fffffff;
}
// the non-ground points, sorted by cell; bounding boxes of the cells
__global__ void k_est_scatter(EstConst ec, EstPtrs ep) {
    pdl_enter();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ec.n; i += gridDim.x * blockDim.x) {
        const float4 w = ep.W[i];
        if (__float_as_int(w.w) != 2) continue;
        const int c = ep.cellid[ep.cellof[i]];
        ep.SW[ep.cbase[c] + ep.pos[i]] = make_float4(w.x, w.y, w.z, __int_as_float(i));
        int *bb = ep.bbox + 6 * c;
        atomicMin(bb, est_ordered(w.x)); atomicMin(bb + 1, est_ordered(w.y)); atomicMin(bb + 2, est_ordered(w.z));
        atomicMax(bb + 3, est_ordered(w.x)); atomicMax(bb + 4, est_ordered(w.y)); atomicMax(bb + 5, est_ordered(w.z));
    }
}
__device__ __forceinline__ float est_unordered(int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff); }
// Links.  The cell edge is 0.57 x tolerance (below tolerance / sqrt(3), like the host implementation): the points of one cell
// are mutually linked, and two cells belong together as soon as ONE pair of their points is within the tolerance; cells up
// to two apart can hold such a pair.  A lane per (occupied cell, cell of the lexicographically positive half of its
// 5 x 5 x 5 neighbourhood) looks the neighbour up and drops the pair if it is missing, already in the same component, or
// if the two cells' bounding boxes are further apart than the tolerance; the warp then searches the remaining pairs of its
// 32 lanes one after the other, 32 points of the neighbour at a time against the points of the cell.
__global__ void k_est_link(EstConst ec, EstPtrs ep) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const long long total = 62ll * ep.cnt[EC_NCELLS];
    const long long wstride = (long long)gridDim.x * blockDim.x;
    for (long long base = blockIdx.x * (long long)blockDim.x + (threadIdx.x & ~31); base < total; base += wstride) {
        const long long t = base + lane;
        int c = -1, cb = -1;
        if (t < total) {
            c = (int)(t / 62);
            const int q = (int)(t - 62ll * c) + 63;  // index in the 5 x 5 x 5 cube, behind its centre (62)
            const u64 ka = ep.hkey[ep.cells[c]];
            const int x = (int)((ka >> 42) & 0x1FFFFF) + q % 5 - 2, y = (int)((ka >> 21) & 0x1FFFFF) + (q / 5) % 5 - 2,
                      z = (int)(ka & 0x1FFFFF) + q / 25 - 2;
            const u64 kb = ((u64)(x & 0x1FFFFF) << 42) | ((u64)(y & 0x1FFFFF) << 21) | (u64)(z & 0x1FFFFF);
            unsigned h = est_hash(kb) & ec.hash_mask;
            for (;;) {
                const u64 k = ep.hkey[h];
                if (k == kb) { cb = ep.cellid[h]; break; }
                if (k == ~0ull) break;
                h = (h + 1) & ec.hash_mask;
            }
            if (cb >= 0) {
                const int *A = ep.bbox + 6 * c, *B = ep.bbox + 6 * cb;
                float d2 = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float g = fmaxf(fmaxf(est_unordered(B[k]) - est_unordered(A[k + 3]), est_unordered(A[k]) - est_unordered(B[k + 3])), 0.f);
                    d2 += g * g;
                }
                if (d2 > ec.tol2 * 1.0001f) cb = -1;  // no pair of the two cells can be within the tolerance (margin for the different rounding)
            }
            if (cb >= 0) {  // most neighbouring cells are linked by their first few points: every lane tries up to 3 x 3 pairs itself
                const int na = min(ep.ccells[c], 3), nb = min(ep.ccells[cb], 3);
                const float4 *PA = ep.SW + ep.cbase[c], *PB = ep.SW + ep.cbase[cb];
                bool hit = false;
                for (int i = 0; i < na; ++i) {
                    const float4 w = PA[i];
                    for (int j = 0; j < nb; ++j) {
                        const float4 v = PB[j];
                        const float ddx = v.x - w.x, ddy = v.y - w.y, ddz = v.z - w.z;
                        hit = hit || ddx * ddx + ddy * ddy + ddz * ddz <= ec.tol2;
                    }
                }
                if (hit) {
                    est_union(ep.parent, c, cb);
                    cb = -1;
                } else if (ep.ccells[c] <= 3 && ep.ccells[cb] <= 3) {
                    cb = -1;  // every pair has been tried
                }
            }
        }
        unsigned todo = __ballot_sync(FULLMASK, cb >= 0);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int a = __shfl_sync(FULLMASK, c, src), b = __shfl_sync(FULLMASK, cb, src);
            int same = 0;
            if (lane == 0) same = est_find(ep.parent, a) == est_find(ep.parent, b);
            if (__shfl_sync(FULLMASK, same, 0)) continue;
            const int na = ep.ccells[a], nb = ep.ccells[b];
            const float4 *PA = ep.SW + ep.cbase[a], *PB = ep.SW + ep.cbase[b];
            bool linked = false;
            for (int j0 = 0; j0 < nb && !linked; j0 += 32) {
                const bool have = j0 + lane < nb;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (have) v = PB[j0 + lane];
                bool hit = false;
                for (int i = 0; i < na && !hit; ++i) {
                    const float4 w = PA[i];
                    const float ddx = v.x - w.x, ddy = v.y - w.y, ddz = v.z - w.z;
                    hit = __any_sync(FULLMASK, have && ddx * ddx + ddy * ddy + ddz * ddz <= ec.tol2);
                }
                linked = hit;
            }
            if (linked && lane == 0) est_union(ep.parent, a, b);
        }
    }
}

// component of every non-ground point: root cell, smallest point index and size of the component
__global__ void k_est_label(EstConst ec, EstPtrs ep) {
    pdl_enter();
    int in_view = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ec.n; i += gridDim.x * blockDim.x) {
        const int cls = __float_as_int(ep.W[i].w);
        in_view += cls != 0;
        int rc = -1;
        if (cls == 2) {
            rc = est_find(ep.parent, ep.cellid[ep.cellof[i]]);
            atomicMin(&ep.cmin[rc], i);
            atomicAdd(&ep.ccount[rc], 1);
        }
        ep.rootc[i] = rc;
        ep.kidx[i] = -1;
    }
    for (int d = 16; d > 0; d >>= 1) in_view += __shfl_xor_sync(FULLMASK, in_view, d);
    if ((threadIdx.x & 31) == 0 && in_view) atomicAdd(&ep.cnt[EC_NV], in_view);
}
// One block: a component is named by its smallest point index; rank of every ground point, list of the clusters of admissible
// size in ascending order of their first member.
__global__ void __launch_bounds__(1024) k_est_clusters(EstConst ec, EstPtrs ep) {
    pdl_enter();
    __shared__ int wsum[32];
    const int n = ec.n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int rc = ep.rootc[i];
        const int l = rc >= 0 ? ep.cmin[rc] : -1;
        const int sz = rc >= 0 ? ep.ccount[rc] : 0;
        ep.label[i] = l;
        if (l == i) ep.csize[i] = sz;
        ep.flag_r[i] = l == i && sz >= EST_MIN_CLUSTER && sz <= EST_MAX_CLUSTER;
    }
    __syncthreads();
    block_exclusive_scan(ep.flag_g, ep.grank, n, wsum);
    block_exclusive_scan(ep.flag_r, ep.mrank, n, wsum);  // (mrank is scratch here; the member ranks are written by the next kernel)
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (ep.flag_r[i]) {
            const int k = ep.mrank[i];
            ep.croot[k] = i;
            ep.csz[k] = ep.csize[i];
            ep.kidx[i] = k;
        }
    if (threadIdx.x == 0) {
        ep.cnt[EC_NGROUND] = ep.grank[n];
        ep.cnt[EC_NC] = ep.mrank[n];
    }
}

// One block per cluster (the kernel is launched even without clusters: its last block lays out the frame): its position in PCL's output order, the rank of every member among the members (ascending index),
// and for clusters small enough to be dynamic the centroid, summed in that order (:1424-1434).
__global__ void __launch_bounds__(256) k_est_features(EstConst ec, EstPtrs ep) {
    pdl_enter();
    __shared__ int s_cnt[8], s_bef[8];
    __shared__ int s_list[EST_DYN_MAX_POINTS];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nc = ep.cnt[EC_NC], n = ec.n;
    for (int k = blockIdx.x; k < nc; k += gridDim.x) {
        const int r = ep.croot[k], s = ep.csz[k];
        const bool small = s <= EST_DYN_MAX_POINTS;
        int before = 0;
        for (int j = tid; j < nc; j += 256) {
            const int sj = ep.csz[j];
            before += (sj > s) || (sj == s && j < k);
        }
        // thread t owns the indices [r + t * per, r + (t + 1) * per): the members are counted, then ranked
        const int per = (n - r + 255) >> 8;
        const int i0 = min(n, r + tid * per), i1 = min(n, i0 + per);
        int mine = 0;
        for (int i = i0; i < i1; ++i) mine += ep.label[i] == r;
        int incl = mine;
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(FULLMASK, incl, d);
            if (lane >= d) incl += t;
        }
        for (int d = 16; d > 0; d >>= 1) before += __shfl_xor_sync(FULLMASK, before, d);
        if (lane == 31) s_cnt[wid] = incl;
        if (lane == 0) s_bef[wid] = before;
        __syncthreads();
        int run = incl - mine;
        for (int w = 0; w < wid; ++w) run += s_cnt[w];
        for (int i = i0; i < i1; ++i)
            if (ep.label[i] == r) {
                ep.mrank[i] = run;
                if (small) s_list[run] = i;
                ++run;
            }
        __syncthreads();
        if (wid == 0) {
            float sx = 0.f, sy = 0.f, sz = 0.f;
            if (small)
                for (int m0 = 0; m0 < s; m0 += 32) {
                    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (m0 + lane < s) w = ep.W[s_list[m0 + lane]];
                    const int m1 = min(32, s - m0);
                    for (int l = 0; l < m1; ++l) {
                        sx += __shfl_sync(FULLMASK, w.x, l);
                        sy += __shfl_sync(FULLMASK, w.y, l);
                        sz += __shfl_sync(FULLMASK, w.z, l);
                    }
                }
            if (lane == 0) {
                int bef = 0;
                for (int w = 0; w < 8; ++w) bef += s_bef[w];
                const float cx = sx / (float)s, cy = sy / (float)s, cz = sz / (float)s;
                ep.cfeat[k] = make_float4(cx, cy, cz, 0.f);
                ep.cdyn[k] = !(s > EST_DYN_MAX_POINTS || cz > EST_DYN_MAX_HEIGHT);
                ep.spos[k] = bef;
            }
        }
        __syncthreads();
    }
    // The last block to finish walks the clusters in output order — offsets of their points inside the dynamic / static parts
    // of the tagged cloud, sequence numbers of the dynamic ones — and hands the dynamic clusters' features to the host.
    __shared__ int s_last;
    __shared__ int wsum[32];
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&ep.cnt[EC_TICKET2], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int k = threadIdx.x; k < nc; k += blockDim.x) {
        const int p = ((volatile int *)ep.spos)[k], d = ((volatile int *)ep.cdyn)[k];
        ep.order[p] = k;
        ep.a_dyn[p] = d;
        ep.a_dsz[p] = d ? ep.csz[k] : 0;
        ep.a_ssz[p] = d ? 0 : ep.csz[k];
    }
    __syncthreads();
    block_exclusive_scan(ep.a_dyn, ep.p_dyn, nc, wsum);
    block_exclusive_scan(ep.a_dsz, ep.p_dsz, nc, wsum);
    block_exclusive_scan(ep.a_ssz, ep.p_ssz, nc, wsum);
    for (int p = threadIdx.x; p < nc; p += blockDim.x) {
        const int k = ep.order[p];
        if (ep.a_dyn[p]) {
            const int q = ep.p_dyn[p];
            ep.coff[k] = ep.p_dsz[p];
            ep.dseq[k] = q;
            const float4 f = __ldcg(ep.cfeat + k);
            EstFeature o;
            o.cx = f.x; o.cy = f.y; o.cz = f.z;
            o.size = ep.csz[k];
            o.sorted_pos = p;
            ep.h_feat[q] = o;
        } else {
            ep.coff[k] = ep.p_ssz[p];
            ep.dseq[k] = -1;
        }
    }
    if (threadIdx.x == 0) {
        const int nv = ep.cnt[EC_NV];
        const int ndyn = ep.p_dyn[nc], ndp = ep.p_dsz[nc], nsp = ep.p_ssz[nc];
        ep.cnt[EC_NDYN] = ndyn;
        ep.cnt[EC_NDYNPTS] = ndp;
        ep.cnt[EC_NSTATICPTS] = nsp;
        if (nv > 0) ep.cnt[EC_NTAGGED] = ndp + ep.cnt[EC_NGROUND] + nsp;  // nothing in view: the previous cloud is kept (:1379)
        else if (ec.nt_override >= 0) ep.cnt[EC_NTAGGED] = ec.nt_override;
        for (int c = 0; c < EC_COUNT; ++c) ep.h_hdr[c] = ((volatile int *)ep.cnt)[c];
        __threadfence_system();
    }
}

// the tagged cloud: positions, zero velocities (the dynamic clusters' are filled in by k_est_apply), padding
__global__ void k_est_write(EstConst ec, EstPtrs ep) {
    pdl_enter();
    const int nv = ep.cnt[EC_NV];
    const int nt = ep.cnt[EC_NTAGGED];
    const int ndp = ep.cnt[EC_NDYNPTS], ng = ep.cnt[EC_NGROUND];
    const int pad_from = nv > 0 ? nt : ec.n_pad_prev;  // nothing in view: only a grown tail is initialised
    const int span = max(ec.n, ec.n_pad);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < span; i += gridDim.x * blockDim.x) {
        if (nv > 0 && i < ec.n) {
            const float4 w = ep.W[i];
            const int cls = __float_as_int(w.w);
            int pos = -1, cid = -1;
            if (cls == 1) {
                pos = ndp + ep.grank[i];
            } else if (cls == 2) {
                const int k = ep.kidx[ep.label[i]];
                if (k >= 0) {
                    cid = ep.dseq[k];
                    pos = (cid >= 0 ? 0 : ndp + ng) + ep.coff[k] + ep.mrank[i];
                }
            }
            if (pos >= 0) {
                float *o = ep.tagged + 7 * (size_t)pos;
                o[0] = w.x; o[1] = w.y; o[2] = w.z;
                o[3] = 0.f; o[4] = 0.f; o[5] = 0.f; o[6] = 0.f;
                ep.tcid[pos] = cid;
            }
        }
        const int j = pad_from + i;
        if (j < ec.n_pad) {
            float *o = ep.tagged + 7 * (size_t)j;
            o[0] = EST_FAR; o[1] = EST_FAR; o[2] = EST_FAR;
            o[3] = 0.f; o[4] = 0.f; o[5] = 0.f; o[6] = 0.f;
            ep.tcid[j] = -1;
        }
    }
}

// velocities and colours of the dynamic clusters (host: match_clusters) onto their points (:1503-1524)
__global__ void k_est_apply(EstPtrs ep) {
    pdl_enter();
    const int ndp = ep.cnt[EC_NDYNPTS];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ndp; i += gridDim.x * blockDim.x) {
        const float4 v = ep.cvel[ep.tcid[i]];
        float *o = ep.tagged + 7 * (size_t)i;
        o[3] = v.x; o[4] = v.y; o[5] = v.z; o[6] = v.w;
    }
}
