// prefilter.cu — the application's point-cloud preprocessing on the GPU (product code).
//
// Replaces, in front of DSPMap::update, what g-ch/DSP-map src/map_sim_example.cpp:305-336 does on the CPU for every depth
// frame: pcl::VoxelGrid down-sampling at `res` (ex:312-316), the camera -> map axis swap x = z, y = -x, z = -y (ex:320-322),
// the open-interval crop to the map extent (ex:325, inRange ex:190-197) and the MAX_POINT_NUM cut (ex:332-334).
//
// pcl::VoxelGrid is third-party code absent from the reference tree (PCL 1.8 / 1.10 as bundled with ROS Melodic / Noetic,
// readme.md:23-25; filters/include/pcl/filters/impl/voxel_grid.hpp).  Its published algorithm, restated here:
//   * points with a non-finite coordinate are skipped;
//   * min_p / max_p = coordinate-wise extremes of the rest; inv = 1 / leaf (fp32);
//   * min_b = (int)floor(min_p * inv), max_b likewise, div_b = max_b - min_b + 1;
//   * leaf index of a point = ijk . (1, div_b.x, div_b.x * div_b.y) with ijk = (int)(floor(p * inv) - (float)min_b);
//   * one output point per occupied leaf, in ascending leaf index: the centroid of the leaf's points.
// Leaf membership, counts and output order are integer work and reproduced exactly.  PCL sums a leaf's points in fp32 in
// the order an UNSTABLE std::sort leaves them, i.e. its last bits are implementation-defined; here each coordinate is
// accumulated exactly in 2^-24 m fixed point (64-bit integer atomics: order-free, so the result is deterministic and
// bit-identical to the CPU restatement the tests hold) and divided once in fp64.  The two agree to PCL's own rounding error.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/dspmap_b200.h"

namespace {

thread_local std::string g_perr;

#define PCK(x)                                                             \
    do {                                                                   \
        cudaError_t e_ = (x);                                              \
        if (e_ != cudaSuccess) {                                           \
            g_perr = std::string(#x) + ": " + cudaGetErrorString(e_);      \
            return DSPMAP_E_CUDA;                                          \
        }                                                                  \
    } while (0)

const int kSMs = 148;
const int PF_SCAN_BLOCK = 2048;  // leaves per scan block
const double PF_FIX = 16777216.0;  // 2^24 fixed-point steps per metre

struct PfHeader {
    unsigned mm[6];      // order-preserving encodings of min x,y,z / max x,y,z over the finite points
    int min_b[3], div_b[3];
    long long vol;       // leaves in the bounding box
    int n_finite;
    int n_kept;          // leaves that survive the crop (before the capacity cut)
    int status;          // 0 ok, 1 no finite point, 2 bounding box larger than the leaf capacity
    float inv_leaf;
};

struct PfParams {
    const float *pts;
    int n, stride;
    float leaf;
    float lo[3], hi[3];  // crop, map axes, open interval
    int cap_out;
    long long cap_leaves;
};

__device__ __forceinline__ unsigned pf_enc(float f) {  // monotone float -> unsigned
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float pf_dec(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
__device__ __forceinline__ bool pf_finite(float x, float y, float z) { return isfinite(x) && isfinite(y) && isfinite(z); }

__global__ void k_pf_reset(PfHeader *h) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        for (int k = 0; k < 3; ++k) { h->mm[k] = 0xffffffffu; h->mm[3 + k] = 0u; }
        h->n_finite = 0; h->n_kept = 0; h->status = 0; h->vol = 0;
    }
}

// getMinMax3D over the finite points: warp-reduced, then one atomic per warp and coordinate
__global__ void __launch_bounds__(256) k_pf_minmax(PfParams P, PfHeader *h) {
    unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
    int cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += gridDim.x * blockDim.x) {
        const float *p = P.pts + (size_t)i * P.stride;
        const float x = p[0], y = p[1], z = p[2];
        if (!pf_finite(x, y, z)) continue;
        const unsigned e[3] = {pf_enc(x), pf_enc(y), pf_enc(z)};
        for (int k = 0; k < 3; ++k) { lo[k] = min(lo[k], e[k]); hi[k] = max(hi[k], e[k]); }
        ++cnt;
    }
    for (int d = 16; d > 0; d >>= 1) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = min(lo[k], __shfl_down_sync(0xffffffffu, lo[k], d));
            hi[k] = max(hi[k], __shfl_down_sync(0xffffffffu, hi[k], d));
        }
        cnt += __shfl_down_sync(0xffffffffu, cnt, d);
    }
    if ((threadIdx.x & 31) == 0 && cnt) {
        for (int k = 0; k < 3; ++k) { atomicMin(&h->mm[k], lo[k]); atomicMax(&h->mm[3 + k], hi[k]); }
        atomicAdd(&h->n_finite, cnt);
    }
}

// voxel_grid.hpp: inverse leaf size, min_b / max_b / div_b, and the capacity check
__global__ void k_pf_header(PfParams P, PfHeader *h) {
    if (threadIdx.x || blockIdx.x) return;
    if (h->n_finite == 0) { h->status = 1; return; }
    const float inv = 1.0f / P.leaf;
    h->inv_leaf = inv;
    long long vol = 1;
    for (int k = 0; k < 3; ++k) {
        const float fl = floorf(pf_dec(h->mm[k]) * inv), fh = floorf(pf_dec(h->mm[3 + k]) * inv);
        if (!(fabsf(fl) < 1e9f && fabsf(fh) < 1e9f)) { h->status = 2; h->vol = -1; return; }
        const int mnb = (int)fl, mxb = (int)fh;
        h->min_b[k] = mnb;
        h->div_b[k] = mxb - mnb + 1;
        vol *= (long long)(mxb - mnb + 1);
        if (vol > P.cap_leaves) { h->status = 2; h->vol = vol; return; }
    }
    h->vol = vol;
}

// Per point: leaf index, then exact fixed-point accumulation.  Neighbouring depth pixels share leaves, so lanes with the
// same leaf are combined first (integer sums: exact in any order) and only their leader touches memory.
__global__ void __launch_bounds__(256) k_pf_accumulate(PfParams P, const PfHeader *h, int *cnt, unsigned long long *sum) {
    if (h->status) return;
    const float inv = h->inv_leaf;
    const int mb0 = h->min_b[0], mb1 = h->min_b[1], mb2 = h->min_b[2];
    const int d0 = h->div_b[0], d01 = h->div_b[0] * h->div_b[1];
    const int nround = (P.n + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
    for (int r = 0; r < nround; ++r) {
        const int i = (r * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
        int leaf = -1;
        long long fx = 0, fy = 0, fz = 0;
        if (i < P.n) {
            const float *p = P.pts + (size_t)i * P.stride;
            const float x = p[0], y = p[1], z = p[2];
            if (pf_finite(x, y, z)) {
                const int i0 = (int)(floorf(x * inv) - (float)mb0), i1 = (int)(floorf(y * inv) - (float)mb1),
                          i2 = (int)(floorf(z * inv) - (float)mb2);
                leaf = i0 + i1 * d0 + i2 * d01;
                fx = __double2ll_rn((double)x * PF_FIX);
                fy = __double2ll_rn((double)y * PF_FIX);
                fz = __double2ll_rn((double)z * PF_FIX);
            }
        }
        const unsigned peers = __match_any_sync(0xffffffffu, leaf);
        const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
        int c = leaf >= 0 ? 1 : 0;
        // tree reduction inside each peer group (an arbitrary lane subset) onto its lowest lane: at step d the member of
        // rank r (r a multiple of 2d) adds the member of rank r + d
        const int rank = __popc(peers & ((1u << lane) - 1u)), members = __popc(peers);
        const int widest = __reduce_max_sync(0xffffffffu, members);
        for (int d = 1; d < widest; d <<= 1) {
            const bool take = (rank & (2 * d - 1)) == 0 && rank + d < members;
            const int src = take ? (int)__fns(peers, 0, rank + d + 1) : lane;
            const long long ax = __shfl_sync(0xffffffffu, fx, src), ay = __shfl_sync(0xffffffffu, fy, src),
                            az = __shfl_sync(0xffffffffu, fz, src);
            const int ac = __shfl_sync(0xffffffffu, c, src);
            if (take) { fx += ax; fy += ay; fz += az; c += ac; }
        }
        if (leaf >= 0 && lane == leader) {
            atomicAdd(&cnt[leaf], c);
            atomicAdd(&sum[3 * (size_t)leaf], (unsigned long long)fx);
            atomicAdd(&sum[3 * (size_t)leaf + 1], (unsigned long long)fy);
            atomicAdd(&sum[3 * (size_t)leaf + 2], (unsigned long long)fz);
        }
    }
}

// centroid of one leaf, swapped to map axes (ex:320-322); returns whether it passes the open-interval crop (ex:325)
__device__ __forceinline__ bool pf_leaf_point(const PfParams &P, int n, const unsigned long long *s, float *o) {
    const double dn = PF_FIX * (double)n;
    const float cx = (float)((double)(long long)s[0] / dn), cy = (float)((double)(long long)s[1] / dn),
                cz = (float)((double)(long long)s[2] / dn);
    o[0] = cz; o[1] = -cx; o[2] = -cy;
    return o[0] > P.lo[0] && o[0] < P.hi[0] && o[1] > P.lo[1] && o[1] < P.hi[1] && o[2] > P.lo[2] && o[2] < P.hi[2];
}

// per scan block: how many of its leaves are occupied and survive the crop
__global__ void __launch_bounds__(256) k_pf_blockcount(PfParams P, const PfHeader *h, const int *cnt, const unsigned long long *sum, int *blocksum) {
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    const long long vol = h->status ? 0 : h->vol;
    const long long b = (long long)blockIdx.x * PF_SCAN_BLOCK;
    int c = 0;
    if (b < vol)
        for (int i = threadIdx.x; i < PF_SCAN_BLOCK; i += blockDim.x) {
            const long long leaf = b + i;
            if (leaf >= vol) break;
            const int n = cnt[leaf];
            float o[3];
            if (n > 0 && pf_leaf_point(P, n, sum + 3 * leaf, o)) ++c;
        }
    for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s, c);
    __syncthreads();
    if (threadIdx.x == 0) blocksum[blockIdx.x] = s;
}

// exclusive scan of the block counts (one block; the counts of blocks past the bounding box are 0)
__global__ void __launch_bounds__(1024) k_pf_scan(const int *in, int *out, int n, PfHeader *h) {
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int x = i < n ? in[i] : 0;
        int incl = x;
        for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int y0 = wsum[lane], y = y0;
            for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, y, d); if (lane >= d) y += t; }
            wsum[lane] = y - y0;
        }
        __syncthreads();
        const int c0 = carry;
        if (i < n) out[i] = c0 + wsum[wid] + incl - x;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c0 + wsum[31] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) h->n_kept = carry;
}

// Emits the surviving centroids in ascending leaf order (the first cap_out of them, ex:332-334), and returns every
// touched accumulator to zero so that the next frame starts from a clean grid without a memset of the whole capacity.
__global__ void __launch_bounds__(256) k_pf_emit(PfParams P, const PfHeader *h, int *cnt, unsigned long long *sum, const int *blockoff,
                                                 float *out, int *n_out) {
    __shared__ int wsum[8];
    const long long vol = h->status ? 0 : h->vol;
    const long long b = (long long)blockIdx.x * PF_SCAN_BLOCK;
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_out = h->status == 1 ? 0 : (h->status ? -1 : min(h->n_kept, P.cap_out));
    if (b >= vol) return;
    int run = blockoff[blockIdx.x];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i0 = 0; i0 < PF_SCAN_BLOCK; i0 += 256) {
        const long long leaf = b + i0 + threadIdx.x;
        float o[3];
        bool keep = false;
        if (leaf < vol) {
            const int n = cnt[leaf];
            if (n > 0) {
                keep = pf_leaf_point(P, n, sum + 3 * leaf, o);
                cnt[leaf] = 0;
                sum[3 * leaf] = 0; sum[3 * leaf + 1] = 0; sum[3 * leaf + 2] = 0;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wsum[w] = __popc(bal);
        __syncthreads();
        int before = 0, tot = 0;
        for (int k = 0; k < 8; ++k) { const int x = wsum[k]; if (k < w) before += x; tot += x; }
        if (keep) {
            const int pos = run + before + __popc(bal & ((1u << lane) - 1u));
            if (pos < P.cap_out) { out[3 * pos] = o[0]; out[3 * pos + 1] = o[1]; out[3 * pos + 2] = o[2]; }
        }
        run += tot;
        __syncthreads();
    }
}

}  // namespace

struct dspmap_prefilter {
    int device = 0;
    long long cap_leaves = 0;
    int cap_raw_floats = 0, cap_out = 0, nblocks = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    float *d_raw = nullptr, *d_out = nullptr, *h_out = nullptr;
    int *d_cnt = nullptr, *d_blocksum = nullptr, *d_blockoff = nullptr, *d_nout = nullptr, *h_nout = nullptr;
    unsigned long long *d_sum = nullptr;
    PfHeader *d_hdr = nullptr, *h_hdr = nullptr;
    long long launches = 0;
};

extern "C" {

const char *dspmap_prefilter_last_error(void) { return g_perr.c_str(); }

int dspmap_prefilter_create(int device, int max_raw_floats, int max_out_points, long long max_leaves, dspmap_prefilter **out) {
    if (!out || max_raw_floats < 3 || max_out_points < 1) { g_perr = "bad argument"; return DSPMAP_E_BAD_ARG; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        g_perr = "no usable CUDA device (this library has no CPU path)";
        return DSPMAP_E_NO_DEVICE;
    }
    PCK(cudaSetDevice(device));
    dspmap_prefilter *p = new dspmap_prefilter();
    p->device = device;
    p->cap_leaves = std::min<long long>(max_leaves > 0 ? max_leaves : (1ll << 22), 1ll << 30);  // leaf indices are int32
    p->cap_raw_floats = max_raw_floats;
    p->cap_out = max_out_points;
    p->nblocks = (int)((p->cap_leaves + PF_SCAN_BLOCK - 1) / PF_SCAN_BLOCK);
    auto fail = [&](cudaError_t e) {
        g_perr = std::string("allocation: ") + cudaGetErrorString(e);
        dspmap_prefilter_destroy(p);
        return DSPMAP_E_CUDA;
    };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e);
    p->stream = p->own_stream;
    if ((e = cudaMalloc(&p->d_raw, sizeof(float) * (size_t)max_raw_floats)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&p->d_out, sizeof(float) * 3 * (size_t)max_out_points)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&p->d_cnt, sizeof(int) * (size_t)p->cap_leaves)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&p->d_sum, sizeof(unsigned long long) * 3 * (size_t)p->cap_leaves)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&p->d_blocksum, sizeof(int) * (size_t)(p->nblocks + 1))) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&p->d_blockoff, sizeof(int) * (size_t)(p->nblocks + 1))) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&p->d_nout, sizeof(int))) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&p->d_hdr, sizeof(PfHeader))) != cudaSuccess) return fail(e);
    if ((e = cudaMallocHost(&p->h_out, sizeof(float) * 3 * (size_t)max_out_points)) != cudaSuccess) return fail(e);
    if ((e = cudaMallocHost(&p->h_nout, sizeof(int))) != cudaSuccess) return fail(e);
    if ((e = cudaMallocHost(&p->h_hdr, sizeof(PfHeader))) != cudaSuccess) return fail(e);
    if ((e = cudaMemset(p->d_cnt, 0, sizeof(int) * (size_t)p->cap_leaves)) != cudaSuccess) return fail(e);
    if ((e = cudaMemset(p->d_sum, 0, sizeof(unsigned long long) * 3 * (size_t)p->cap_leaves)) != cudaSuccess) return fail(e);
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return fail(e);
    *out = p;
    return DSPMAP_OK;
}

void dspmap_prefilter_destroy(dspmap_prefilter *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    cudaFree(p->d_raw); cudaFree(p->d_out); cudaFree(p->d_cnt); cudaFree(p->d_sum); cudaFree(p->d_blocksum);
    cudaFree(p->d_blockoff); cudaFree(p->d_nout); cudaFree(p->d_hdr);
    if (p->h_out) cudaFreeHost(p->h_out);
    if (p->h_nout) cudaFreeHost(p->h_nout);
    if (p->h_hdr) cudaFreeHost(p->h_hdr);
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    cudaGetLastError();
    delete p;
}

int dspmap_prefilter_set_stream(dspmap_prefilter *p, void *stream) {
    if (!p) return DSPMAP_E_BAD_ARG;
    p->stream = stream ? (cudaStream_t)stream : p->own_stream;
    return DSPMAP_OK;
}

long long dspmap_prefilter_launches(const dspmap_prefilter *p) { return p ? p->launches : 0; }

// Device-resident variant: only enqueues (7 kernels).  d_n_out receives the number of points written (<= cap), or -1 when
// the cloud's bounding box needs more leaves than the capacity given at create time.
int dspmap_prefilter_run_device(dspmap_prefilter *p, int n, int stride, const float *d_pts, float leaf, const float *range_min,
                                const float *range_max, float *d_out, int cap, int *d_n_out) {
    if (!p || n < 0 || stride < 3 || !(leaf > 0.f) || !range_min || !range_max || !d_out || !d_n_out || cap < 1 || (n > 0 && !d_pts)) {
        g_perr = "bad argument";
        return DSPMAP_E_BAD_ARG;
    }
    PCK(cudaSetDevice(p->device));
    PfParams P;
    P.pts = d_pts; P.n = n; P.stride = stride; P.leaf = leaf; P.cap_out = cap; P.cap_leaves = p->cap_leaves;
    for (int k = 0; k < 3; ++k) { P.lo[k] = range_min[k]; P.hi[k] = range_max[k]; }
    const int B = 256;
    const int gpts = (int)std::min<long long>(std::max<long long>(((long long)n + B - 1) / B, 1), kSMs * 8);
    k_pf_reset<<<1, 32, 0, p->stream>>>(p->d_hdr);
    k_pf_minmax<<<gpts, B, 0, p->stream>>>(P, p->d_hdr);
    k_pf_header<<<1, 32, 0, p->stream>>>(P, p->d_hdr);
    k_pf_accumulate<<<gpts, B, 0, p->stream>>>(P, p->d_hdr, p->d_cnt, p->d_sum);
    k_pf_blockcount<<<p->nblocks, B, 0, p->stream>>>(P, p->d_hdr, p->d_cnt, p->d_sum, p->d_blocksum);
    k_pf_scan<<<1, 1024, 0, p->stream>>>(p->d_blocksum, p->d_blockoff, p->nblocks, p->d_hdr);
    k_pf_emit<<<p->nblocks, B, 0, p->stream>>>(P, p->d_hdr, p->d_cnt, p->d_sum, p->d_blockoff, d_out, d_n_out);
    p->launches += 7;
    PCK(cudaGetLastError());
    return DSPMAP_OK;
}

// Host variant (what the application calls instead of ex:305-336): raw cloud in (n points, `stride` floats apart, camera
// frame), filtered + swapped + cropped cloud out (xyz triples, ready for DSPMap::update).  One H2D of the raw cloud, one
// D2H of at most cap points; the copy is a DMA when the caller's buffers are page-locked.
int dspmap_prefilter_run(dspmap_prefilter *p, int n, int stride, const float *pts, float leaf, const float *range_min,
                         const float *range_max, float *out, int cap, int *n_out) {
    if (!p || !n_out || n < 0 || stride < 3 || (n > 0 && !pts)) { g_perr = "bad argument"; return DSPMAP_E_BAD_ARG; }
    if ((long long)n * stride > p->cap_raw_floats) { g_perr = "raw cloud larger than max_raw_floats"; return DSPMAP_E_CAPACITY; }
    PCK(cudaSetDevice(p->device));
    const int c = std::min(cap, p->cap_out);
    if (n > 0) PCK(cudaMemcpyAsync(p->d_raw, pts, sizeof(float) * (size_t)n * stride, cudaMemcpyHostToDevice, p->stream));
    int rc = dspmap_prefilter_run_device(p, n, stride, p->d_raw, leaf, range_min, range_max, p->d_out, c, p->d_nout);
    if (rc != DSPMAP_OK) return rc;
    PCK(cudaMemcpyAsync(p->h_nout, p->d_nout, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    PCK(cudaMemcpyAsync(p->h_out, p->d_out, sizeof(float) * 3 * (size_t)c, cudaMemcpyDeviceToHost, p->stream));
    PCK(cudaStreamSynchronize(p->stream));
    if (*p->h_nout < 0) {
        g_perr = "the cloud's bounding box needs more leaves than max_leaves";
        *n_out = 0;
        return DSPMAP_E_CAPACITY;
    }
    *n_out = *p->h_nout;
    if (out) memcpy(out, p->h_out, sizeof(float) * 3 * (size_t)*n_out);  // out == NULL: the result stays in the page-locked buffer
    return DSPMAP_OK;
}

// ex:305-353 in one call: preprocessing on the GPU, then DSPMap::update on the filtered cloud.  The raw cloud crosses PCIe
// once; what comes back to the host is the filtered cloud (<= max_out_points points) the velocity estimation needs.
int dspmap_update_raw(dspmap *m, dspmap_prefilter *p, int n, int stride, const float *raw, float leaf, const float *range_min,
                      const float *range_max, float px, float py, float pz, double t, float qw, float qx, float qy, float qz,
                      int *n_filtered) {
    int nf = 0;
    int rc = dspmap_prefilter_run(p, n, stride, raw, leaf, range_min, range_max, nullptr, p ? p->cap_out : 0, &nf);
    if (rc != DSPMAP_OK) return rc;
    if (n_filtered) *n_filtered = nf;
    return dspmap_update(m, nf, 3, p->h_out, px, py, pz, t, qw, qx, qy, qz);
}

}  // extern "C"
