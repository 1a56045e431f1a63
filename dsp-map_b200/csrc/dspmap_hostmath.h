// dspmap_hostmath.h — the bit-exact scalar helpers of dspmap_kernels.cuh for plain C++ translation units.
#pragma once
#include <cstdint>
#include "dspmap_types.h"
#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
inline void dsp_rotate(const float *v, const float *q, const float *qi, float *o) {
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
inline uint32_t dsp_u31(u64 seed, u64 k) {
    u64 z = seed + (k + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 33);
}
inline float dsp_uniform(u64 seed, u64 k, float lo, float hi) {
    int r = (int)dsp_u31(seed, k);
    return lo + (float)r / ((float)(2147483647 / (hi - lo)));
}
#endif
