// check_misc_kernels.cpp — CPU check of two more kernel variants under simt_host.h (TEST INFRASTRUCTURE):
//   k_norm_fast against k_norm (one fp32 chain over all 1/C_z: same bits), and
//   k_fut_count / k_fut_compact (the sparse copy-out of the future grid) against the rows a host loop finds.
#include "simt_host.h"

#include "dspmap_kernels.cuh"
#include "misc_kernels.inc"

#include <random>

int main() {
    int bad = 0;
    std::mt19937 rng(5);
    auto uni = [&](float lo, float hi) { return lo + (hi - lo) * (float)(rng() >> 8) / 16777216.f; };
    // ---- newborn normaliser
    for (int n : {0, 1, 15, 16, 17, 1023, 1024, 1025, 4096, 4097, 9973}) {
        MapConst mc;
        FrameConst fc;
        memset(&mc, 0, sizeof(mc));
        memset(&fc, 0, sizeof(fc));
        mc.P = 3;
        fc.nb_weight = 1e-4f;
        std::vector<int> capoff = {0, n / 3, n / 2, n};
        std::vector<float> INV(n + 8);
        for (int i = 0; i < n; ++i) INV[i] = 1.f / uni(0.02f, 40.f);
        DevState st[2];
        for (int k = 0; k < 2; ++k) {
            memset(&st[k], 0, sizeof(DevState));
            DevPtrs dp;
            memset(&dp, 0, sizeof(dp));
            dp.obs_capoff = capoff.data(); dp.INV = INV.data(); dp.st = &st[k];
            if (k == 0) simt::launch_block(128, [&] { k_norm(mc, fc, dp); });
            else simt::launch_block(256, [&] { k_norm_fast(mc, fc, dp); });
        }
        const bool ok = memcmp(&st[0].norm, &st[1].norm, 4) == 0 && memcmp(&st[0].w_new, &st[1].w_new, 4) == 0;
        printf("normaliser over %5d terms: %.9g %s\n", n, st[0].norm, ok ? "identical" : "DIFFERENT");
        bad += ok ? 0 : 1;
    }
    // ---- sparse future rows
    for (int V : {1, 511, 512, 513, 5000}) {
        MapConst mc;
        memset(&mc, 0, sizeof(mc));
        mc.V = V;
        mc.T = 6;
        std::vector<float> fut((size_t)V * mc.T, 0.f);
        std::vector<int> want;
        for (int v = 0; v < V; ++v)
            if (rng() % 7 == 0) {
                fut[(size_t)v * mc.T + rng() % mc.T] = uni(0.001f, 2.f);
                if (rng() % 2) fut[(size_t)v * mc.T + rng() % mc.T] = uni(0.001f, 2.f);
                want.push_back(v);
            }
        const int nblocks = (V + OCC_BLOCK - 1) / OCC_BLOCK;
        std::vector<int> cnt(nblocks + 1, -1), off(nblocks + 1, 0), fidx(V, -1);
        std::vector<float> fval((size_t)V * mc.T, -1.f);
        int nf = -1;
        simt::launch_grid(nblocks, 256, [&] { k_fut_count(mc, fut.data(), cnt.data()); });
        for (int b = 0; b < nblocks; ++b) off[b + 1] = off[b] + cnt[b];  // k_scan_small
        simt::launch_grid(nblocks, 256, [&] { k_fut_compact(mc, fut.data(), off.data(), fidx.data(), fval.data(), &nf, nblocks); });
        bool ok = nf == (int)want.size();
        for (int k = 0; ok && k < nf; ++k)
            ok = fidx[k] == want[k] && memcmp(&fval[(size_t)k * mc.T], &fut[(size_t)want[k] * mc.T], 4 * mc.T) == 0;
        printf("sparse future rows, %4d voxels: %d rows %s\n", V, nf, ok ? "identical" : "DIFFERENT");
        bad += ok ? 0 : 1;
    }
    return bad ? 1 : 0;
}
