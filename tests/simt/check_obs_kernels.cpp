// check_obs_kernels.cpp — CPU check of the observation-pass kernel variants (TEST INFRASTRUCTURE; built and run by
// tests/test_simt_cpu.py).  The kernels' own source (extract.py -> obs_kernels.inc) runs under simt_host.h, one OS thread
// per CUDA thread, one block per launch.  Reference pipeline = the GPU-verified row-major kernels (k_pair_eval ->
// k_cz_chain -> k_weight2); every variant must reproduce its C_z, 1/C_z and new particle weights bit for bit:
//   k_cz_chain<.., STG>, k_cz_chain_tma            on the row-major buffer
//   k_weight2_t<true>  (dsp_quot fast path)        on the row-major buffer
//   k_pair_eval_col -> k_cz_chain_col -> k_weight_col<false / true>   on the column-major buffer
#include "simt_host.h"

#include "dspmap_kernels.cuh"
#include "obs_kernels.inc"

#include <cmath>
#include <random>

namespace {
int Nh = 4, Nv = 3, P = Nh * Nv, NBN = 1, NBW = 10;  // geometry of the current scene (3 x 3 or 5 x 5 neighbourhoods)
const int OBS = 100;

struct Scene {
    MapConst mc;
    FrameConst fc;
    std::vector<int> nbr, obs_cnt, obs_maxbits, obs_capoff, plen, poff, LA;
    std::vector<float4> OBSP, LP;
    std::vector<float> PW, lut, w0;
    int n_fov = 0, n_pts = 0;
};

Scene make_scene(unsigned seed, int nh, int nv, int nbn) {
    Nh = nh; Nv = nv; P = nh * nv; NBN = nbn; NBW = (2 * nbn + 1) * (2 * nbn + 1) + 1;
    Scene s;
    std::mt19937 rng(seed);
    auto uni = [&](float lo, float hi) { return lo + (hi - lo) * (float)(rng() >> 8) / 16777216.f; };
    memset(&s.mc, 0, sizeof(s.mc));
    memset(&s.fc, 0, sizeof(s.fc));
    s.mc.P = P; s.mc.NB = NBW - 1; s.mc.NBW = NBW; s.mc.OBS = OBS; s.mc.Nh = Nh; s.mc.Nv = Nv; s.mc.occl = 0.3f;
    s.mc.cap_pairs = 1ll << 40;
    s.fc.sigma = 0.1f; s.fc.sigma_r = 1.f / 0.1f; s.fc.fast_sigma = 0; s.fc.Pd = 0.95f; s.fc.one_minus_Pd = 1 - 0.95f;
    s.fc.kappa = 0.01f; s.fc.nb_weight = 1e-4f; s.fc.nb_num = 20;
    // PDF half table exactly as dspmap_create builds it (dsp_dynamic.h:1282-1292)
    s.lut.resize(DSP_LUT_HALF);
    for (int h = 0; h < 10000; ++h) {
        const float x = (float)h * 0.001f;
        s.lut[h] = (1.f / (sqrtf(2.f * 1.57079632679489661923))) * expf(-powf(x, 2) / (2));
    }
    s.lut[10000] = s.lut[9999];
    s.nbr.assign(P * NBW, 0);
    for (int p = 0; p < P; ++p) {
        const int h = p / Nv, v = p % Nv;
        int n = 0;
        for (int i = -NBN; i <= NBN; ++i)
            for (int j = -NBN; j <= NBN; ++j)
                if (h + i >= 0 && h + i < Nh && v + j >= 0 && v + j < Nv) s.nbr[p * NBW + 1 + n++] = (h + i) * Nv + v + j;
        s.nbr[p * NBW] = n;
    }
    // a surface patch per pyramid; points on it, particles scattered around it (some far: subnormal / zero pair terms)
    s.obs_cnt.assign(P, 0); s.obs_maxbits.assign(P, __float_as_int(-1.f)); s.obs_capoff.assign(P + 1, 0);
    s.plen.assign(P, 0); s.poff.assign(P + 1, 0);
    s.OBSP.assign((size_t)P * OBS, make_float4(0, 0, 0, 0));
    std::vector<float4> centre(P);
    for (int p = 0; p < P; ++p) {
        centre[p] = make_float4(2.f + 0.3f * (p / Nv), -0.4f + 0.35f * (p % Nv) , 0.2f * (p % 2), 0.f);
        const int kind = rng() % 6;
        const int np = kind == 0 ? 0 : kind == 1 ? 1 : kind == 2 ? 99 + (int)(rng() % 30) : 3 + (int)(rng() % 45);  // counts above 99 are clamped by the kernels
        s.obs_cnt[p] = np;
        float maxlen = -1.f;
        for (int z = 0; z < std::min(np, OBS - 1); ++z) {
            const float4 o = make_float4(centre[p].x + uni(-0.15f, 0.15f), centre[p].y + uni(-0.15f, 0.15f), centre[p].z + uni(-0.15f, 0.15f), 0.f);
            const float len = sqrtf(o.x * o.x + o.y * o.y + o.z * o.z);
            s.OBSP[(size_t)p * OBS + z] = make_float4(o.x, o.y, o.z, len);
            maxlen = std::max(maxlen, len);
        }
        if (np > 0) s.obs_maxbits[p] = __float_as_int(maxlen);
        const int pk = rng() % 6;
        s.plen[p] = pk == 0 ? 0 : pk == 1 ? 1 + (int)(rng() % 31) : pk == 2 ? 129 + (int)(rng() % 60) : 20 + (int)(rng() % 100);
    }
    for (int p = 0; p < P; ++p) {
        s.poff[p + 1] = s.poff[p] + s.plen[p];
        s.obs_capoff[p + 1] = s.obs_capoff[p] + std::min(s.obs_cnt[p], OBS - 1);
    }
    s.n_fov = s.poff[P];
    s.n_pts = s.obs_capoff[P];
    s.LP.resize(s.n_fov); s.PW.resize(s.n_fov + 8); s.LA.resize(s.n_fov); s.w0.resize(s.n_fov);
    for (int p = 0; p < P; ++p)
        for (int k = 0; k < s.plen[p]; ++k) {
            const int i = s.poff[p] + k;
            const int far = rng() % 10;  // 0: behind the surface by > 1 m in one axis, 1: in two axes, 2: beyond the occlusion margin
            float dx = uni(-0.25f, 0.25f), dy = uni(-0.25f, 0.25f), dz = uni(-0.25f, 0.25f);
            if (far == 0) dx += 1.2f;
            if (far == 1) { dx -= 1.1f; dy += 1.3f; }
            if (far == 2) dx += 0.8f;
            const float w = uni(0.002f, 0.08f);
            s.LP[i] = make_float4(centre[p].x + dx, centre[p].y + dy, centre[p].z + dz, w);
            s.PW[i] = s.fc.Pd * w;
            s.LA[i] = i;
            s.w0[i] = w;
        }
    return s;
}

struct Result { std::vector<float> CZ, INV, W; };
bool same(const std::vector<float> &a, const std::vector<float> &b, const char *what) {
    if (a.size() == b.size() && memcmp(a.data(), b.data(), 4 * a.size()) == 0) return true;
    size_t k = 0;
    while (k < a.size() && k < b.size() && memcmp(&a[k], &b[k], 4) == 0) ++k;
    printf("MISMATCH: %s (first at %zu: %.9g vs %.9g)\n", what, k, k < a.size() ? a[k] : 0.f, k < b.size() ? b[k] : 0.f);
    return false;
}

enum CzKernel { CZ_DEFAULT, CZ_STAGED, CZ_TMA, CZ_COL };
// One pass of the three observation kernels over the scene; col selects the column-major family, packed its evaluation
// kernel with packed fp32 arithmetic (taken only with the verified fast division, so fast_sigma is switched on for it: on
// the host the scalar path divides and the packed path runs the multiply / fused-correction sequence — the table index
// they produce is the same for every reachable argument, which is what k_verify_div establishes on the device).
Result run(const Scene &s, bool col, CzKernel czk, bool quot_fast, bool packed = false, bool fused_prep = false) {
    MapConst mc = s.mc;
    FrameConst fc = s.fc;
    if (packed) fc.fast_sigma = 1;
    DevState st;
    memset(&st, 0, sizeof(st));
    st.n_valid = s.n_pts;
    std::vector<int> cum(P * NBW), totlen(P), pairs(P + 1), rowbase(P + 1), chunks(P + 1), chunk_off(P + 1), cz_order(P);
    std::vector<float4> PA(s.n_fov);
    for (int i = 0; i < s.n_fov; ++i) PA[i] = make_float4(0, 0, 0, s.w0[i]);
    std::vector<float> CZ((size_t)P * OBS, -7.f), INV(s.n_pts + 8, -7.f);
    DevPtrs dp;
    memset(&dp, 0, sizeof(dp));
    dp.st = &st; dp.nbr = s.nbr.data(); dp.obs_cnt = const_cast<int *>(s.obs_cnt.data()); dp.obs_maxbits = const_cast<int *>(s.obs_maxbits.data());
    dp.obs_capoff = const_cast<int *>(s.obs_capoff.data()); dp.plen = const_cast<int *>(s.plen.data()); dp.poff = const_cast<int *>(s.poff.data());
    dp.OBSP = const_cast<float4 *>(s.OBSP.data()); dp.LP = const_cast<float4 *>(s.LP.data()); dp.PW = const_cast<float *>(s.PW.data());
    dp.LA = const_cast<int *>(s.LA.data()); dp.lut = s.lut.data(); dp.PA = PA.data(); dp.CZ = CZ.data(); dp.INV = INV.data();
    dp.cum = cum.data(); dp.totlen = totlen.data(); dp.pairs = pairs.data(); dp.rowbase = rowbase.data(); dp.chunks = chunks.data();
    dp.chunk_off = chunk_off.data();
    dp.cz_order = (col || czk == CZ_TMA) ? cz_order.data() : nullptr;
    if (fused_prep) {  // DSPMAP_FUSE_SCAN: preparation and both scans in one block
        simt::launch_block(1024, [&] { k_pair_prep_scan(mc, dp, col ? 1 : 0); });
    } else {
        simt::launch_block(32, [&] { k_pair_prep(mc, dp, col ? 1 : 0); });
        rowbase[0] = chunk_off[0] = 0;  // the two exclusive scans k_scan_small performs
        for (int p = 0; p < P; ++p) { rowbase[p + 1] = rowbase[p] + pairs[p]; chunk_off[p + 1] = chunk_off[p] + chunks[p]; }
    }
    float *G = static_cast<float *>(aligned_alloc(256, sizeof(float) * ((size_t)rowbase[P] + 64 + 64)));
    for (size_t i = 0; i < (size_t)rowbase[P] + 128; ++i) G[i] = NAN;  // every element a consumer reads must have been produced
    dp.G = G;
    if (col && packed) simt::launch_block(EVAL_THREADS, [&] { k_pair_eval_col<true>(mc, fc, dp, 0); });
    else if (col) simt::launch_block(EVAL_THREADS, [&] { k_pair_eval_col<false>(mc, fc, dp, 0); });
    else simt::launch_block(EVAL_THREADS, [&] { k_pair_eval(mc, fc, dp, 0); });
    switch (czk) {
    case CZ_DEFAULT: simt::launch_block(256, [&] { k_cz_chain<256, 8192, 128>(mc, fc, dp); }); break;
    case CZ_STAGED: simt::launch_block(256, [&] { k_cz_chain<256, 8192, 128, true>(mc, fc, dp); }); break;
    case CZ_TMA: simt::launch_block(CZT_THREADS, [&] { k_cz_chain_tma(mc, fc, dp); }); break;
    case CZ_COL: simt::launch_block(CZC_THREADS, [&] { k_cz_chain_col(mc, fc, dp); }); break;
    }
    if (col) {
        if (quot_fast) simt::launch_block(32, [&] { k_weight_col<true>(mc, fc, dp); });
        else simt::launch_block(32, [&] { k_weight_col<false>(mc, fc, dp); });
    } else {
#ifdef CHECK_W2W  // built with -DW2_SWITCH=0: the warp-per-chunk kernel takes every frame
        if (quot_fast) simt::launch_block(W2W_THREADS, [&] { k_weight2w_t<true>(mc, fc, dp); });
        else simt::launch_block(W2W_THREADS, [&] { k_weight2w_t<false>(mc, fc, dp); });
#else
        if (quot_fast) simt::launch_block(W2_THREADS, [&] { k_weight2_t<true>(mc, fc, dp); }, 96);
        else simt::launch_block(W2_THREADS, [&] { k_weight2_t<false>(mc, fc, dp); }, 96);
#endif
    }
    free(G);
    Result r;
    r.CZ = CZ;
    r.INV.assign(INV.begin(), INV.begin() + s.n_pts);
    r.W.resize(s.n_fov);
    for (int i = 0; i < s.n_fov; ++i) r.W[i] = PA[i].w;
    return r;
}
// The same pass sharded over `nranks` ranks the way dspmap_shard_phase runs it: rank r evaluates and chains the point
// pyramids i % nranks == r (mode 1) into a zero-initialised C_z buffer, the buffers are summed (all-reduce: one writer per
// element); then rank r evaluates and weighs the particle chunks c % nranks == r (mode 2) into a zero-initialised weight
// buffer, which is summed and applied through the list's slot addresses.
Result run_sharded(const Scene &s, bool col, int nranks) {
    const FrameConst fc = s.fc;
    std::vector<float> CZ((size_t)P * OBS, 0.f), INV(s.n_pts + 8, 0.f), NW(s.n_fov + 8, 0.f);
    std::vector<float4> PA(s.n_fov);
    for (int i = 0; i < s.n_fov; ++i) PA[i] = make_float4(0, 0, 0, s.w0[i]);
    for (int phase = 0; phase < 2; ++phase)
        for (int rank = 0; rank < nranks; ++rank) {
            MapConst mc = s.mc;
            mc.sharded = 1; mc.nranks = nranks; mc.rank = rank;
            DevState st;
            memset(&st, 0, sizeof(st));
            st.n_valid = s.n_pts;
            std::vector<int> cum(P * NBW), totlen(P), pairs(P + 1), rowbase(P + 1), chunks(P + 1), chunk_off(P + 1), cz_order(P);
            std::vector<float> cz_r((size_t)P * OBS, 0.f), inv_r(s.n_pts + 8, 0.f), nw_r(s.n_fov + 8, 0.f);
            DevPtrs dp;
            memset(&dp, 0, sizeof(dp));
            dp.st = &st; dp.nbr = s.nbr.data(); dp.obs_cnt = const_cast<int *>(s.obs_cnt.data()); dp.obs_maxbits = const_cast<int *>(s.obs_maxbits.data());
            dp.obs_capoff = const_cast<int *>(s.obs_capoff.data()); dp.plen = const_cast<int *>(s.plen.data()); dp.poff = const_cast<int *>(s.poff.data());
            dp.OBSP = const_cast<float4 *>(s.OBSP.data()); dp.LP = const_cast<float4 *>(s.LP.data()); dp.PW = const_cast<float *>(s.PW.data());
            dp.LA = const_cast<int *>(s.LA.data()); dp.lut = s.lut.data(); dp.PA = PA.data();
            dp.CZ = phase == 0 ? cz_r.data() : CZ.data();    // phase 1 reads the merged C_z
            dp.INV = phase == 0 ? inv_r.data() : INV.data();
            dp.NW = nw_r.data();
            dp.cum = cum.data(); dp.totlen = totlen.data(); dp.pairs = pairs.data(); dp.rowbase = rowbase.data(); dp.chunks = chunks.data();
            dp.chunk_off = chunk_off.data();
            dp.cz_order = col ? cz_order.data() : nullptr;
            simt::launch_block(32, [&] { k_pair_prep(mc, dp, col ? 1 : 0); });
            rowbase[0] = chunk_off[0] = 0;
            for (int p = 0; p < P; ++p) { rowbase[p + 1] = rowbase[p] + pairs[p]; chunk_off[p + 1] = chunk_off[p] + chunks[p]; }
            float *G = static_cast<float *>(aligned_alloc(256, sizeof(float) * ((size_t)rowbase[P] + 128)));
            for (size_t i = 0; i < (size_t)rowbase[P] + 128; ++i) G[i] = NAN;
            dp.G = G;
            const int mode = phase + 1;
            if (col) simt::launch_block(EVAL_THREADS, [&] { k_pair_eval_col<false>(mc, fc, dp, mode); });
            else simt::launch_block(EVAL_THREADS, [&] { k_pair_eval(mc, fc, dp, mode); });
            if (phase == 0) {
                if (col) simt::launch_block(CZC_THREADS, [&] { k_cz_chain_col(mc, fc, dp); });
                else simt::launch_block(256, [&] { k_cz_chain<256, 8192, 128>(mc, fc, dp); });
                for (size_t i = 0; i < CZ.size(); ++i) CZ[i] += cz_r[i];
                for (int i = 0; i < s.n_pts; ++i) INV[i] += inv_r[i];
            } else {
                if (col) simt::launch_block(32, [&] { k_weight_col<false>(mc, fc, dp); });
                else simt::launch_block(W2_THREADS, [&] { k_weight2_t<false>(mc, fc, dp); }, 96);
                for (int i = 0; i < s.n_fov; ++i) NW[i] += nw_r[i];
            }
            free(G);
        }
    Result r;
    r.CZ = CZ;
    r.INV.assign(INV.begin(), INV.begin() + s.n_pts);
    r.W.resize(s.n_fov);
    for (int i = 0; i < s.n_fov; ++i) r.W[i] = NW[s.LA[i]];  // k_shard_apply_weights: PA[LA[i]].w = NW[i], LA is the identity here
    return r;
}
// C_z entries that exist (pyramids with points): the sharded buffers are zero where the single-rank pass leaves them untouched
std::vector<float> valid_cz(const Scene &s, const std::vector<float> &cz) {
    std::vector<float> out;
    for (int p = 0; p < P; ++p)
        for (int z = 0; z < std::min(s.obs_cnt[p], OBS - 1); ++z) out.push_back(cz[(size_t)p * OBS + z]);
    return out;
}
}  // namespace

int main() {
    int bad = 0;
    for (unsigned seed = 1; seed <= 3; ++seed) {
        const Scene s = seed < 3 ? make_scene(seed, 4, 3, 1) : make_scene(seed, 5, 5, 2);  // the last one: multiple-neighbours variant
#ifdef CHECK_W2W
        const Result ref = run(s, true, CZ_COL, false);  // the column-major family (equal to k_weight2 by the other build of this check)
#else
        const Result ref = run(s, false, CZ_DEFAULT, false);
#endif
        int changed = 0, tiny = 0;
        for (int i = 0; i < s.n_fov; ++i) changed += ref.W[i] != s.w0[i];
        printf("scene %u: %d pyramids, %d registered particles, %d binned points; reference pass changed %d weights\n", seed, P, s.n_fov, s.n_pts, changed);
        if (changed < s.n_fov / 2) { printf("the scene does not exercise the weight pass\n"); ++bad; }
        (void)tiny;
        struct Case { const char *name; bool col; CzKernel cz; bool qf; bool packed = false; bool fused_prep = false; } cases[] = {
#ifdef CHECK_W2W
            {"k_weight2w (warp per chunk)", false, CZ_DEFAULT, false},
            {"k_weight2w<QF>", false, CZ_DEFAULT, true},
#else
            {"k_cz_chain<STG>", false, CZ_STAGED, false},
            {"k_cz_chain_tma", false, CZ_TMA, false},
            {"k_weight2<QF>", false, CZ_DEFAULT, true},
            {"column-major family", true, CZ_COL, false},
            {"column-major family + dsp_quot fast path", true, CZ_COL, true},
            {"column-major family, packed evaluation", true, CZ_COL, false, true},
            {"k_pair_prep_scan (row-major)", false, CZ_DEFAULT, false, false, true},
            {"k_pair_prep_scan (column-major)", true, CZ_COL, false, false, true},
#endif
        };
        for (const Case &c : cases) {
            const Result r = run(s, c.col, c.cz, c.qf, c.packed, c.fused_prep);
            bool ok = same(ref.CZ, r.CZ, "C_z");
            ok &= same(ref.INV, r.INV, "1 / C_z");
            ok &= same(ref.W, r.W, "particle weights");
            printf("  %-42s %s\n", c.name, ok ? "identical" : "DIFFERENT");
            bad += ok ? 0 : 1;
        }
#ifndef CHECK_W2W
        for (int col = 0; col < 2; ++col) {
            const Result r = run_sharded(s, col != 0, seed == 2 ? 3 : 2);
            bool ok = same(valid_cz(s, ref.CZ), valid_cz(s, r.CZ), "C_z");
            ok &= same(ref.INV, r.INV, "1 / C_z");
            ok &= same(ref.W, r.W, "particle weights");
            printf("  %-42s %s\n", col ? "column-major family, sharded" : "row-major family, sharded", ok ? "identical" : "DIFFERENT");
            bad += ok ? 0 : 1;
        }
#endif
    }
    return bad ? 1 : 0;
}
