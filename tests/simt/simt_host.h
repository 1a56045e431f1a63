// simt_host.h — a minimal host-side SIMT shim (TEST INFRASTRUCTURE): runs ONE warp of a warp-per-item CUDA kernel on the
// CPU, one OS thread per lane, so that the kernel's own source (sliced out of dspmap_frame.cuh by extract.py) can be
// checked against another kernel on the same inputs without a GPU.  Supported: threadIdx / blockIdx / blockDim / gridDim,
// __shared__ (one CTA at a time), the warp collectives used by those kernels (__ballot_sync, __shfl*_sync, __syncwarp,
// __reduce_min_sync; full masks only, convergent call sites), integer / float atomics, bit intrinsics.
// Not supported (and not needed here): __syncthreads, asynchronous copies, mbarriers, divergent collectives.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <barrier>
#include <climits>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)

namespace simt {
inline thread_local uint3 t_idx, b_idx;
inline dim3 b_dim(32, 1, 1), g_dim(1, 1, 1);
inline std::barrier<> *warp_barrier = nullptr;
inline unsigned long long bus[32];
inline void sync() { warp_barrier->arrive_and_wait(); }
template <typename T>
inline T exchange(T v, int src) {  // every lane publishes v, then reads lane src's
    static_assert(sizeof(T) <= 8, "collective payload");
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    bus[t_idx.x & 31] = raw;
    sync();
    T out;
    memcpy(&out, &bus[src & 31], sizeof(T));
    sync();
    return out;
}
// Runs kernel(args...) with one block of 32 threads (one warp).
template <typename F>
inline void launch_one_warp(F &&body) {
    std::barrier<> bar(32);
    warp_barrier = &bar;
    std::vector<std::thread> th;
    for (int l = 0; l < 32; ++l)
        th.emplace_back([&, l] {
            t_idx = uint3{(unsigned)l, 0, 0};
            b_idx = uint3{0, 0, 0};
            body();
        });
    for (auto &t : th) t.join();
    warp_barrier = nullptr;
}
}  // namespace simt

#define threadIdx simt::t_idx
#define blockIdx simt::b_idx
#define blockDim simt::b_dim
#define gridDim simt::g_dim

inline void __syncwarp(unsigned = 0xffffffffu) { simt::sync(); }
inline unsigned __ballot_sync(unsigned, bool p) {
    unsigned r = 0;
    simt::bus[threadIdx.x & 31] = p ? 1ull : 0ull;
    simt::sync();
    for (int l = 0; l < 32; ++l) r |= (unsigned)(simt::bus[l] & 1ull) << l;
    simt::sync();
    return r;
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src) { return simt::exchange(v, src); }
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int m) { return simt::exchange(v, (int)(threadIdx.x & 31) ^ m); }
template <typename T>
inline T __shfl_up_sync(unsigned, T v, int d) { const int l = threadIdx.x & 31; return simt::exchange(v, l >= d ? l - d : l); }
template <typename T>
inline T __shfl_down_sync(unsigned, T v, int d) { const int l = threadIdx.x & 31; return simt::exchange(v, l + d < 32 ? l + d : l); }
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    simt::bus[threadIdx.x & 31] = v;
    simt::sync();
    unsigned r = 0xffffffffu;
    for (int l = 0; l < 32; ++l) r = (unsigned)simt::bus[l] < r ? (unsigned)simt::bus[l] : r;
    simt::sync();
    return r;
}
inline unsigned __activemask() { return 0xffffffffu; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
template <typename T>
inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline float atomicAdd(float *p, float v) {
    // one warp, lanes run concurrently: serialise float adds in lane order so that two runs see the same order
    static std::atomic_flag lock = ATOMIC_FLAG_INIT;
    while (lock.test_and_set(std::memory_order_acquire)) {}
    const float o = *p;
    *p = o + v;
    lock.clear(std::memory_order_release);
    return o;
}
inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAnd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }
inline int atomicOr(int *p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
using std::max;
using std::min;
