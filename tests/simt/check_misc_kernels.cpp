// check_misc_kernels.cpp — CPU check of two more kernel variants under simt_host.h (TEST INFRASTRUCTURE):
//   k_norm_fast against k_norm (one fp32 chain over all 1/C_z: same bits), and
//   k_fut_count / k_fut_compact (the sparse copy-out of the future grid) against the rows a host loop finds,
//   k_pyr_sort_w (warp-local sorting stages) against k_pyr_sort and against std::sort.
#include "simt_host.h"

#include "dspmap_kernels.cuh"
#include "misc_kernels.inc"
#include "sparse_rows.h"

#include <algorithm>
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
    // ---- scans fused into their producers (last block done): reader counts, newborn draw counts
    for (int V : {300, 512, 5000}) {
        MapConst mc;
        memset(&mc, 0, sizeof(mc));
        mc.V = V;
        mc.T = 2;
        const int nblocks = (V + OCC_BLOCK - 1) / OCC_BLOCK;
        std::vector<float4> occv(V);
        for (auto &o : occv) o = make_float4(uni(0.f, 0.5f), 0, 0, 0);
        std::vector<float> fut0((size_t)V * mc.T);
        for (auto &f : fut0) f = rng() % 3 ? 0.f : uni(0.1f, 1.f);
        std::vector<int> cnt[2], off[2];
        std::vector<float> fut[2], dfut[2];
        DevState st;
        memset(&st, 0, sizeof(st));
        for (int k = 0; k < 2; ++k) {
            cnt[k].assign(nblocks + 1, -1); off[k].assign(nblocks + 1, -1); fut[k] = fut0; dfut[k].assign(fut0.size(), -1.f);
            DevPtrs dp;
            memset(&dp, 0, sizeof(dp));
            dp.OCCV = occv.data(); dp.FUT = fut[k].data(); dp.st = &st;
            if (k == 0) {
                simt::launch_grid(nblocks, 256, [&] { k_occ_count(mc, dp, 0.2f, cnt[0].data(), dfut[0].data()); });
                off[0][0] = 0;
                for (int b = 0; b < nblocks; ++b) off[0][b + 1] = off[0][b] + cnt[0][b];
            } else {
                simt::launch_grid(nblocks, 256, [&] { k_occ_count_fs(mc, dp, 0.2f, cnt[1].data(), off[1].data(), nblocks, dfut[1].data()); });
            }
        }
        cnt[0][nblocks] = cnt[1][nblocks] = 0;
        const bool ok = cnt[0] == cnt[1] && off[0] == off[1] && fut[0] == fut[1] && dfut[0] == dfut[1] && st.tickets[2] == 0;
        printf("reader count + fused scan, %4d voxels (%d blocks): %s\n", V, nblocks, ok ? "identical" : "DIFFERENT");
        bad += ok ? 0 : 1;
    }
    {
        MapConst mc;
        FrameConst fc;
        memset(&mc, 0, sizeof(mc));
        memset(&fc, 0, sizeof(fc));
        mc.V = 40; mc.S = 48; mc.model = 0; mc.v_hi = mc.V;
        fc.n_tagged = 700; fc.nb_num = 20; fc.nb_model_gen = 16; fc.nb_min_static = 3;
        const int n = fc.n_tagged;
        std::vector<float4> PA((size_t)mc.V * mc.S), PB(PA.size()), NPC(n);
        std::vector<ulonglong2> M(mc.V);
        std::vector<int> ninmap(n + 1);
        std::vector<u64> nimask(n);
        std::vector<float> tagged((size_t)n * 7);
        for (int v = 0; v < mc.V; ++v) {
            M[v] = make_ulonglong2((((u64)rng() << 32) | rng()) & ((1ull << 48) - 1ull), 0ull);
            if (rng() % 6 == 0) M[v].x = 0ull;
            for (int sl = 0; sl < mc.S; ++sl) {
                PA[(size_t)v * mc.S + sl] = make_float4(0, 0, 0, uni(0.001f, 0.05f));
                const float fl[4] = {1.f, 7.f, 15.f, 0.6f};
                const int kind = rng() % 3;
                PB[(size_t)v * mc.S + sl] = make_float4(kind == 0 ? 0.f : uni(-0.6f, 0.6f), kind == 2 ? uni(-0.6f, 0.6f) : 0.f, 0.f, fl[rng() % 4]);
            }
        }
        for (int m = 0; m < n; ++m) {
            ninmap[m] = rng() % 8 != 0;
            NPC[m] = make_float4(0, 0, 0, __int_as_float((int)(rng() % mc.V)));
            nimask[m] = (u64)(rng() & 0xfffff);
            float *pt = &tagged[(size_t)m * 7];
            pt[3] = rng() % 3 ? uni(-1.f, 1.f) : -10000.f;
            pt[6] = rng() % 2 ? 1.f : 0.f;
        }
        std::vector<int> nstatic[2], nvcnt[2], nrcnt[2], nvoff[2], nroff[2];
        DevState st;
        memset(&st, 0, sizeof(st));
        const int nblocks = (n * 32 + 255) / 256;
        for (int k = 0; k < 2; ++k) {
            nstatic[k].assign(n, -9); nvcnt[k].assign(n + 1, -9); nrcnt[k].assign(n + 1, -9); nvoff[k].assign(n + 1, -9); nroff[k].assign(n + 1, -9);
            DevPtrs dp;
            memset(&dp, 0, sizeof(dp));
            dp.PA = PA.data(); dp.PB = PB.data(); dp.M = M.data(); dp.NPC = NPC.data(); dp.ninmap = ninmap.data(); dp.nimask = nimask.data();
            dp.tagged = tagged.data(); dp.nstatic = nstatic[k].data(); dp.nvcnt = nvcnt[k].data(); dp.nrcnt = nrcnt[k].data();
            dp.nvoff = nvoff[k].data(); dp.nroff = nroff[k].data(); dp.st = &st;
            if (k == 0) {
                simt::launch_grid(nblocks, 256, [&] { k_nb_point1(mc, fc, dp, 0); });
                nvoff[0][0] = nroff[0][0] = 0;
                for (int m = 0; m < n; ++m) { nvoff[0][m + 1] = nvoff[0][m] + nvcnt[0][m]; nroff[0][m + 1] = nroff[0][m] + nrcnt[0][m]; }
            } else {
                simt::launch_grid(nblocks, 256, [&] { k_nb_point1_fs(mc, fc, dp, 0); });
            }
        }
        nvcnt[0][n] = nvcnt[1][n] = nrcnt[0][n] = nrcnt[1][n] = 0;
        // points outside the map leave nstatic untouched in both
        const bool ok = nstatic[0] == nstatic[1] && nvcnt[0] == nvcnt[1] && nrcnt[0] == nrcnt[1] && nvoff[0] == nvoff[1] && nroff[0] == nroff[1] &&
                        nvoff[0][n] > 0 && nroff[0][n] > 0 && st.tickets[1] == 0;
        printf("newborn draw counts + fused scans, %d points (%d blocks): %d table draws, %d uniform draws: %s\n", n, nblocks, nvoff[0][n], nroff[0][n],
               ok ? "identical" : "DIFFERENT");
        bad += ok ? 0 : 1;
    }
    // ---- host side of the sparse copy-out: a sequence of grids through SparseRows::apply equals the dense grids
    {
        const int V = 3000, T = 6;
        std::vector<float> bufA((size_t)V * T, 9.f), bufB((size_t)V * T, -3.f), dense((size_t)V * T);
        SparseRows sr;
        bool ok = true;
        for (int frame = 0; frame < 12 && ok; ++frame) {
            std::fill(dense.begin(), dense.end(), 0.f);
            std::vector<int> idx;
            std::vector<float> val;
            for (int v = 0; v < V; ++v)
                if (rng() % (frame % 3 == 0 ? 5 : 40) == 0) {
                    for (int t = 0; t < T; ++t) dense[(size_t)v * T + t] = rng() % 2 ? uni(0.01f, 1.f) : 0.f;
                    dense[(size_t)v * T + rng() % T] = uni(0.01f, 1.f);
                    idx.push_back(v);
                    val.insert(val.end(), dense.begin() + (size_t)v * T, dense.begin() + (size_t)(v + 1) * T);
                }
            float *target = frame == 6 ? bufB.data() : bufA.data();  // the application switches arrays once, and back
            if (frame == 9) { std::fill(bufA.begin(), bufA.end(), 5.f); sr.invalidate(); }  // a dense copy overwrote it
            sr.apply(target, V, T, idx.data(), val.data(), (int)idx.size());
            ok = memcmp(target, dense.data(), sizeof(float) * dense.size()) == 0;
        }
        printf("sparse rows applied on the host over 12 frames: %s\n", ok ? "identical" : "DIFFERENT");
        bad += ok ? 0 : 1;
    }
    // ---- pyramid-list sort: warp-local stages against the block-barrier network
    for (unsigned seed = 1; seed <= 2; ++seed) {
        std::mt19937 r2(seed);
        MapConst mc;
        memset(&mc, 0, sizeof(mc));
        mc.P = 14; mc.S = 48; mc.L = 700;
        const int sizes[14] = {0, 1, 2, 3, 31, 32, 33, 64, 100, 513, 700, 1024, 1500, 2049};
        std::vector<int> pcount(sizes, sizes + 14), poff(15, 0);
        for (int q = 0; q < 14; ++q) poff[q + 1] = poff[q] + pcount[q];
        const int n = poff[14];
        std::vector<int> key(n), addr(n);
        std::vector<int> perm(n);
        for (int i = 0; i < n; ++i) perm[i] = i;
        std::shuffle(perm.begin(), perm.end(), r2);
        for (int i = 0; i < n; ++i) { key[i] = perm[i] * 3 + 1; addr[i] = (int)(r2() % 4000) * 1 + 0; }
        std::sort(addr.begin(), addr.end());
        addr.erase(std::unique(addr.begin(), addr.end()), addr.end());
        while ((int)addr.size() < n) addr.push_back((int)addr.size() + 5000);  // distinct slot addresses
        std::shuffle(addr.begin(), addr.end(), r2);
        std::vector<float4> PA(20000);
        for (auto &x : PA) x = make_float4((float)(r2() % 1000), 1.f, 2.f, (float)(1 + r2() % 200) / 4096.f);
        struct Out { std::vector<int> LA, plen; std::vector<float4> LP; std::vector<float> PW; std::vector<ulonglong2> M; DevState st; } o[2];
        for (int k = 0; k < 2; ++k) {
            o[k].LA.assign(n, -1); o[k].plen.assign(14, -1); o[k].LP.assign(n, make_float4(-1, -1, -1, -1)); o[k].PW.assign(n, -1.f);
            o[k].M.assign(500, make_ulonglong2(~0ull, ~0ull));
            memset(&o[k].st, 0, sizeof(DevState));
            DevPtrs dp;
            memset(&dp, 0, sizeof(dp));
            dp.pcount = pcount.data(); dp.poff = poff.data(); dp.plen = o[k].plen.data(); dp.PSkey = key.data(); dp.PSaddr = addr.data();
            dp.PA = PA.data(); dp.LA = o[k].LA.data(); dp.LP = o[k].LP.data(); dp.PW = o[k].PW.data(); dp.M = o[k].M.data(); dp.st = &o[k].st;
            if (k == 0) simt::launch_block(512, [&] { k_pyr_sort(mc, dp, 0.95f); });
            else simt::launch_block(512, [&] { k_pyr_sort_w(mc, dp, 0.95f); });
        }
        bool ok = o[0].LA == o[1].LA && o[0].plen == o[1].plen && memcmp(o[0].LP.data(), o[1].LP.data(), 16 * (size_t)n) == 0 &&
                  memcmp(o[0].PW.data(), o[1].PW.data(), 4 * (size_t)n) == 0 && memcmp(o[0].M.data(), o[1].M.data(), 16 * 500) == 0 &&
                  o[0].st.n_pyramid_full == o[1].st.n_pyramid_full && o[0].st.n_pyramid_full > 0;
        // and the reference result is really sorted by key
        for (int q = 0; ok && q < 14; ++q) {
            std::vector<std::pair<int, int>> kv;
            for (int i = poff[q]; i < poff[q + 1]; ++i) kv.push_back({key[i], addr[i]});
            std::sort(kv.begin(), kv.end());
            for (int i = 0; ok && i < std::min((int)kv.size(), mc.L); ++i) ok = o[1].LA[poff[q] + i] == kv[i].second;
        }
        printf("pyramid-list sort seed %u: %d keys in 14 lists, %d dropped beyond L: %s\n", seed, n, o[0].st.n_pyramid_full, ok ? "identical" : "DIFFERENT");
        bad += ok ? 0 : 1;
    }
    return bad ? 1 : 0;
}
