// check_resample.cpp — CPU check of the resampling kernel against the round-1 kernel that was verified on a B200 against
// the reference (TEST INFRASTRUCTURE; built and run by tests/test_simt_cpu.py).  Both kernels' source is sliced out of the
// .cuh files (extract.py -> resample.inc from dsp-map_b200/csrc/dspmap_frame.cuh, legacy_resample.inc from
// tests/simt/legacy_frame_r01.cuh) and compiled for the host against simt_host.h, one OS thread per lane.  Same random
// inputs to both; every output (particles, masks, occupancy, future grid, counters) must be bit-identical.
#include "simt_host.h"


#include "dspmap_kernels.cuh"

#include "resample.inc"
#include "legacy_resample.inc"

#include <cstdio>
#include <random>

namespace {
MapConst make_mc() {
    MapConst mc;
    memset(&mc, 0, sizeof(mc));
    mc.nx = mc.ny = mc.nz = 4;
    mc.V = 64;
    mc.S = 48;
    mc.T = 3;
    mc.max_ppv = 24;
    mc.res = 0.5f;
    mc.hx = mc.hy = mc.hz = 1.0f;
    mc.ft[0] = 0.1f; mc.ft[1] = 0.5f; mc.ft[2] = 1.0f;
    mc.vlo = (1ull << 48) - 1ull;
    mc.vhi = 0ull;
    mc.fast_res = 0;
    mc.res_r = 1.f / mc.res;
    mc.v_hi = mc.V;
    return mc;
}
struct Store {
    std::vector<float4> PA, PB, OCCV;
    std::vector<ulonglong2> M;
    std::vector<float> FUT;
    DevState st;
};
bool same(const void *a, const void *b, size_t n, const char *what) {
    if (memcmp(a, b, n) == 0) return true;
    printf("MISMATCH: %s\n", what);
    return false;
}

int check_resample(unsigned seed) {
    const MapConst mc = make_mc();
    std::mt19937 rng(seed);
    auto uni = [&](float lo, float hi) { return lo + (hi - lo) * (float)(rng() >> 8) / 16777216.f; };
    Store s0;
    s0.PA.assign((size_t)mc.V * mc.S, make_float4(0, 0, 0, 0));
    s0.PB = s0.PA;
    s0.OCCV.assign(mc.V, make_float4(-1, -1, -1, -1));
    s0.M.assign(mc.V, make_ulonglong2(0, 0));
    s0.FUT.assign((size_t)mc.V * mc.T, 0.f);
    memset(&s0.st, 0, sizeof(s0.st));
    std::vector<int> E;
    for (int v = 0; v < mc.V; ++v) {
        const int kind = rng() % 6;  // empty, sparse (< 5), medium, above MAX, full, heavy-tailed weights
        const int cnt = kind == 0 ? 0 : kind == 1 ? 1 + rng() % 4 : kind == 2 ? 5 + rng() % 15 : kind == 3 ? 25 + rng() % 15 : kind == 4 ? 48 : 6 + rng() % 30;
        std::vector<int> slots(mc.S);
        for (int i = 0; i < mc.S; ++i) slots[i] = i;
        std::shuffle(slots.begin(), slots.end(), rng);
        for (int k = 0; k < cnt; ++k) {
            const int sl = slots[k];
            s0.M[v].x |= 1ull << sl;
            const int ix = v % 4, iy = (v / 4) % 4, iz = v / 16;
            // weights are multiples of 2^-12, so the atomically accumulated future sums do not depend on the lanes' order
            int wq = 1 + (int)(rng() % 200);
            if (kind == 5 && k % 5 == 0) wq *= 20;
            if (rng() % 9 == 0) wq = 1 + (int)(rng() % 4);  // below 1e-3: dropped
            const float flags[4] = {1.f, 7.f, 15.f, 0.6f};
            s0.PA[(size_t)v * mc.S + sl] = make_float4(-1.f + 0.5f * ix + uni(0.01f, 0.49f), -1.f + 0.5f * iy + uni(0.01f, 0.49f), -1.f + 0.5f * iz + uni(0.01f, 0.49f),
                                                       (float)wq / 4096.f);
            const bool still = rng() % 3 == 0;
            s0.PB[(size_t)v * mc.S + sl] = make_float4(still ? 0.f : uni(-1.5f, 1.5f), still ? 0.f : uni(-1.5f, 1.5f), 0.f, flags[rng() % 4]);
        }
        if (cnt) E.push_back(v);
    }
    s0.st.n_occ_voxels = (int)E.size();
    Store s[2] = {s0, s0};
    FrameConst fc;
    memset(&fc, 0, sizeof(fc));
    for (int k = 0; k < 2; ++k) {
        DevPtrs dp;
        memset(&dp, 0, sizeof(dp));
        dp.PA = s[k].PA.data(); dp.PB = s[k].PB.data(); dp.M = s[k].M.data(); dp.OCCV = s[k].OCCV.data(); dp.FUT = s[k].FUT.data();
        dp.E = E.data(); dp.st = &s[k].st;
        if (k == 0) simt::launch_one_warp([&] { legacy::k_resample(mc, fc, dp); });
        else simt::launch_grid(GRID, 32 * RS_WARPS, [&] { k_resample(mc, fc, dp); });
    }
    bool ok = same(s[0].PA.data(), s[1].PA.data(), sizeof(float4) * s0.PA.size(), "resample: PA");
    ok &= same(s[0].PB.data(), s[1].PB.data(), sizeof(float4) * s0.PB.size(), "resample: PB");
    ok &= same(s[0].M.data(), s[1].M.data(), sizeof(ulonglong2) * s0.M.size(), "resample: masks");
    ok &= same(s[0].OCCV.data(), s[1].OCCV.data(), sizeof(float4) * s0.OCCV.size(), "resample: occupancy / mean velocity");
    ok &= same(s[0].FUT.data(), s[1].FUT.data(), sizeof(float) * s0.FUT.size(), "resample: future grid");
    ok &= same(&s[0].st, &s[1].st, sizeof(DevState), "resample: counters");
    if (s[0].st.n_pre == 0 || s[0].st.n_out == s[0].st.n_pre) { printf("resample: the inputs did not exercise resampling\n"); ok = false; }
    printf("resample seed %u: %d occupied voxels, %d kept -> %d after, %d dropped for low weight: %s\n", seed, (int)E.size(), s[0].st.n_pre, s[0].st.n_out,
           s[0].st.n_low_weight, ok ? "identical" : "DIFFERENT");
    return ok ? 0 : 1;
}
}  // namespace

int main() {
    int bad = 0;
    for (unsigned seed = 1; seed <= 6; ++seed) bad += check_resample(seed);
    return bad ? 1 : 0;
}
