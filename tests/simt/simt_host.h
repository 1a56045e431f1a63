// simt_host.h — a minimal host-side SIMT shim (TEST INFRASTRUCTURE): runs ONE thread block of a CUDA kernel on the CPU, one
// OS thread per CUDA thread, so that the kernel's own source (sliced out of dspmap_frame.cuh by extract.py) can be checked
// against another kernel on the same inputs without a GPU.  Supported: threadIdx / blockIdx / blockDim / gridDim,
// __shared__ (static + one dynamic buffer; one block at a time), __syncthreads, one named barrier, the warp collectives
// the kernels use (__ballot_sync, __shfl*_sync, __syncwarp, __reduce_min_sync; full masks, convergent call sites), integer
// / float atomics (add, or, and, compare-and-swap, exchange), bit intrinsics, cp.async pipelines (copies complete at once), and the mbarrier / cp.async.bulk subset
// of cuda::ptx (copies complete at once and are checked for the 16-byte rules of the hardware).
// Not modelled: memory ordering weaker than sequential consistency, divergent collectives, several blocks at once.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <barrier>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __align__
#define __align__(n)

namespace simt {
inline thread_local uint3 t_idx, b_idx;
inline dim3 b_dim(32, 1, 1), g_dim(1, 1, 1);
inline std::vector<std::unique_ptr<std::barrier<>>> warp_barriers;
inline std::unique_ptr<std::barrier<>> cta_barrier, named_bar;
inline unsigned long long bus[32][32];  // [warp][lane]
alignas(128) inline unsigned char dyn_smem[232448];
inline int warp_id() { return (int)(t_idx.x >> 5); }
inline void sync() { warp_barriers[warp_id()]->arrive_and_wait(); }
inline void named_barrier(int) { named_bar->arrive_and_wait(); }
template <typename T>
inline T exchange(T v, int src) {  // every lane publishes v, then reads lane src's
    static_assert(sizeof(T) <= 8, "collective payload");
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    bus[warp_id()][t_idx.x & 31] = raw;
    sync();
    T out;
    memcpy(&out, &bus[warp_id()][src & 31], sizeof(T));
    sync();
    return out;
}
// Runs body() on every thread of ONE block of `nthreads` threads (a multiple of 32); named_count = participants of the
// block's named barrier (bar.sync 1, n), 0 if unused.
inline unsigned current_block = 0;
template <typename F>
inline void launch_block(int nthreads, F &&body, int named_count = 0, unsigned nblocks = 1) {
    b_dim = dim3((unsigned)nthreads, 1, 1);
    g_dim = dim3(nblocks, 1, 1);
    warp_barriers.clear();
    for (int w = 0; w < nthreads / 32; ++w) warp_barriers.emplace_back(new std::barrier<>(32));
    cta_barrier.reset(new std::barrier<>(nthreads));
    named_bar.reset(named_count ? new std::barrier<>(named_count) : nullptr);
    std::vector<std::thread> th;
    for (int l = 0; l < nthreads; ++l)
        th.emplace_back([&, l] {
            t_idx = uint3{(unsigned)l, 0, 0};
            b_idx = uint3{current_block, 0, 0};
            body();
            // a thread that returns early must not leave the others waiting at a block barrier
            cta_barrier->arrive_and_drop();
        });
    for (auto &t : th) t.join();
}
template <typename F>
inline void launch_one_warp(F &&body) { launch_block(32, body); }
// A grid whose blocks do not communicate: one block after the other.
template <typename F>
inline void launch_grid(unsigned nblocks, int nthreads, F &&body) {
    for (unsigned b = 0; b < nblocks; ++b) {
        current_block = b;
        launch_block(nthreads, body, 0, nblocks);
    }
    current_block = 0;
}
}  // namespace simt

#define threadIdx simt::t_idx
#define blockIdx simt::b_idx
#define blockDim simt::b_dim
#define gridDim simt::g_dim

inline void __syncthreads() { simt::cta_barrier->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { simt::sync(); }
inline unsigned __ballot_sync(unsigned, bool p) {
    unsigned r = 0;
    const int w = simt::warp_id();
    simt::bus[w][threadIdx.x & 31] = p ? 1ull : 0ull;
    simt::sync();
    for (int l = 0; l < 32; ++l) r |= (unsigned)(simt::bus[w][l] & 1ull) << l;
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
    const int w = simt::warp_id();
    simt::bus[w][threadIdx.x & 31] = v;
    simt::sync();
    unsigned r = 0xffffffffu;
    for (int l = 0; l < 32; ++l) r = (unsigned)simt::bus[w][l] < r ? (unsigned)simt::bus[w][l] : r;
    simt::sync();
    return r;
}
inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0u; }
inline unsigned __activemask() { return 0xffffffffu; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
template <typename T>
inline T __ldg(const T *p) { return *p; }
template <typename T>
inline T __ldcg(const T *p) { return *p; }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
template <typename T>
inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline float atomicAdd(float *p, float v) {
    static std::atomic_flag lock = ATOMIC_FLAG_INIT;
    while (lock.test_and_set(std::memory_order_acquire)) {}
    const float o = *p;
    *p = o + v;
    lock.clear(std::memory_order_release);
    return o;
}
inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAnd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
inline int atomicOr(int *p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int atomicCAS(int *p, int cmp, int v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
inline int atomicMin(int *p, int v) { int o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
inline int atomicMax(int *p, int v) { int o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
inline int atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
using std::max;
using std::min;

// packed fp32 arithmetic of sm_100 (crt/sm_100_rt.h): each half is the scalar IEEE operation
inline float2 __fadd2_rn(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
inline float2 __fmul2_rn(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }

// cp.async pipelines (<cuda_pipeline.h>): the copy lands at once
inline void __pipeline_memcpy_async(void *dst, const void *src, size_t n) { memcpy(dst, src, n); }
inline void __pipeline_commit() {}
inline void __pipeline_wait_prior(int) {}

// the subset of cuda::ptx the kernels use.  An mbarrier is its 64-bit word in "shared memory": phase parity, pending
// arrivals, arrival count and the signed transaction count, all updated under one lock.
namespace cuda { namespace ptx {
struct sem_release_t {}; struct scope_cta_t {}; struct scope_cluster_t {}; struct space_shared_t {}; struct space_cluster_t {}; struct space_global_t {};
inline constexpr sem_release_t sem_release{};
inline constexpr scope_cta_t scope_cta{};
inline constexpr scope_cluster_t scope_cluster{};
inline constexpr space_shared_t space_shared{};
inline constexpr space_cluster_t space_cluster{};
inline constexpr space_global_t space_global{};
struct MbarState { int phase, pending, count; long long tx; };
inline std::mutex &mbar_lock() { static std::mutex m; return m; }
inline MbarState &mbar_state(uint64_t *bar) {  // side table keyed by the barrier's address
    static std::vector<std::pair<uint64_t *, MbarState>> tab;
    for (auto &e : tab) if (e.first == bar) return e.second;
    tab.push_back({bar, MbarState{0, 0, 0, 0}});
    return tab.back().second;
}
inline void mbar_check(MbarState &s) { if (s.pending == 0 && s.tx == 0) { s.phase ^= 1; s.pending = s.count; } }
inline void mbarrier_init(uint64_t *bar, uint32_t count) { std::lock_guard<std::mutex> g(mbar_lock()); mbar_state(bar) = MbarState{0, (int)count, (int)count, 0}; }
inline void fence_mbarrier_init(sem_release_t, scope_cluster_t) {}
inline void fence_proxy_async(space_shared_t) {}
inline uint64_t mbarrier_arrive(uint64_t *bar) { std::lock_guard<std::mutex> g(mbar_lock()); auto &s = mbar_state(bar); --s.pending; mbar_check(s); return 0; }
inline uint64_t mbarrier_arrive_expect_tx(sem_release_t, scope_cta_t, space_shared_t, uint64_t *bar, uint32_t tx) {
    std::lock_guard<std::mutex> g(mbar_lock());
    auto &s = mbar_state(bar);
    s.tx += tx;
    --s.pending;
    mbar_check(s);
    return 0;
}
inline void cp_async_bulk(space_cluster_t, space_global_t, void *dst, const void *src, uint32_t size, uint64_t *bar) {
    if (((uintptr_t)dst & 15) || ((uintptr_t)src & 15) || (size & 15) || size == 0) {
        fprintf(stderr, "cp.async.bulk: dst %p src %p size %u violate the 16-byte rules\n", dst, src, size);
        abort();
    }
    memcpy(dst, src, size);
    std::lock_guard<std::mutex> g(mbar_lock());
    auto &s = mbar_state(bar);
    s.tx -= size;
    mbar_check(s);
}
inline bool mbarrier_try_wait_parity(uint64_t *bar, uint32_t parity) {
    std::lock_guard<std::mutex> g(mbar_lock());
    const bool done = mbar_state(bar).phase != (int)parity;
    if (!done) std::this_thread::yield();
    return done;
}
}}  // namespace cuda::ptx
