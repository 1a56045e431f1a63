// check_nb_place.cpp — CPU check of the newborn placement kernel (one thread per candidate, rank by counting) against the
// round-1 kernel that was verified on a B200 against the reference (a warp per destination voxel, minimum extraction): the
// same candidates must land in the same slots.  TEST INFRASTRUCTURE; built and run by tests/test_simt_cpu.py.
#include "simt_host.h"


#include "dspmap_kernels.cuh"

#include "nb_place.inc"
namespace legacy {  // the members of the round-1 DevPtrs that its placement kernel touches
struct DevPtrs {
    float4 *PA, *PB, *CA, *CB;
    ulonglong2 *M;
    int *cowner, *cbase, *ccnt, *cseg, *csegi;
    DevState *st;
};
}  // namespace legacy
#include "legacy_nb_place.inc"

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

int check_nb_place(unsigned seed) {
    const MapConst mc = make_mc();
    std::mt19937 rng(seed);
    std::vector<ulonglong2> M0(mc.V, make_ulonglong2(0, 0));
    std::vector<int> cowner, cbase(mc.V, 0), ccnt(mc.V, 0), cseg, csegi;
    std::vector<float4> CA, CB;
    int next_key = 0;
    for (int v = 0; v < mc.V; ++v) {
        const int live = rng() % 5 == 0 ? 48 : (int)(rng() % 49);  // some voxels are full
        std::vector<int> slots(mc.S);
        for (int i = 0; i < mc.S; ++i) slots[i] = i;
        std::shuffle(slots.begin(), slots.end(), rng);
        for (int k = 0; k < live; ++k) M0[v].x |= 1ull << slots[k];
        if (rng() % 4 == 0) continue;  // no candidates for this voxel
        const int kind = rng() % 5;
        const int c = kind == 0 ? 1 + rng() % 3 : kind == 1 ? 20 + rng() % 40 : kind == 2 ? 100 + rng() % 100 : kind == 3 ? 257 + rng() % 80 : 30 + rng() % 230;
        cowner.push_back(v);
        cbase[v] = (int)cseg.size();
        ccnt[v] = c;
        std::vector<int> keys(c);
        for (int j = 0; j < c; ++j) keys[j] = next_key + j * 3 + (int)(rng() % 3);  // unique, not sorted after the shuffle
        next_key += 3 * c + 7;
        std::shuffle(keys.begin(), keys.end(), rng);
        for (int j = 0; j < c; ++j) {
            cseg.push_back(keys[j]);
            csegi.push_back((int)CA.size());
            CA.push_back(make_float4((float)keys[j], (float)v, 1.f, 0.25f));
            CB.push_back(make_float4((float)j, 2.f, 0.f, 15.f));
        }
    }
    std::shuffle(cowner.begin(), cowner.end(), rng);
    struct Out { std::vector<float4> PA, PB; std::vector<ulonglong2> M; DevState st; } o[2];
    FrameConst fc;
    memset(&fc, 0, sizeof(fc));
    for (int k = 0; k < 2; ++k) {
        o[k].PA.assign((size_t)mc.V * mc.S, make_float4(0, 0, 0, 0));
        o[k].PB = o[k].PA;
        o[k].M = M0;
        memset(&o[k].st, 0, sizeof(DevState));
        o[k].st.n_cand_owner = (int)cowner.size();
        DevPtrs dp;
        memset(&dp, 0, sizeof(dp));
        dp.PA = o[k].PA.data(); dp.PB = o[k].PB.data(); dp.M = o[k].M.data(); dp.st = &o[k].st;
        dp.cowner = cowner.data(); dp.cbase = cbase.data(); dp.ccnt = ccnt.data(); dp.cseg = cseg.data(); dp.csegi = csegi.data();
        dp.CA = CA.data();
        std::vector<int> Cdst(CA.size()), Caddr(CA.size(), -1);
        for (size_t i = 0; i < CA.size(); ++i) Cdst[i] = (int)CA[i].y;   // (the scene stores the voxel in CA.y)
        std::vector<ulonglong2> MS = M0;                                  // the snapshot the grouping pass takes
        if (k == 0) {
            legacy::DevPtrs lp;
            memset(&lp, 0, sizeof(lp));
            lp.PA = dp.PA; lp.PB = dp.PB; lp.M = dp.M; lp.st = dp.st; lp.cowner = dp.cowner; lp.cbase = dp.cbase; lp.ccnt = dp.ccnt;
            lp.cseg = dp.cseg; lp.csegi = dp.csegi; lp.CA = dp.CA; lp.CB = CB.data();
            simt::launch_one_warp([&] { legacy::k_nb_place(mc, fc, lp); });
        } else {
            dp.Cdst = Cdst.data(); dp.Caddr = Caddr.data(); dp.MS = MS.data(); dp.cap_cand = (int)CA.size();
            o[k].st.cand_top = (int)cseg.size();
            simt::launch_grid(3, 256, [&] { k_nb_place(mc, fc, dp); });
            // the new kernel leaves the velocity to k_nb_fill: identify who landed where through Caddr
            for (size_t i = 0; i < CA.size(); ++i)
                if (Caddr[i] >= 0) o[k].PB[Caddr[i]] = CB[i];
            o[k].st.cand_top = 0;
        }
    }
    bool ok = same(o[0].PA.data(), o[1].PA.data(), sizeof(float4) * o[0].PA.size(), "nb_place: PA");
    ok &= same(o[0].PB.data(), o[1].PB.data(), sizeof(float4) * o[0].PB.size(), "nb_place: which candidate took which slot");
    ok &= same(o[0].M.data(), o[1].M.data(), sizeof(ulonglong2) * o[0].M.size(), "nb_place: masks");
    ok &= o[0].st.n_born == o[1].st.n_born && o[0].st.n_born > 0;
    printf("nb_place seed %u: %d destination voxels, %d candidates, %d born: %s\n", seed, (int)cowner.size(), (int)cseg.size(), o[0].st.n_born,
           ok ? "identical" : "DIFFERENT");
    return ok ? 0 : 1;
}
}  // namespace

int main() {
    int bad = 0;
    for (unsigned seed = 1; seed <= 4; ++seed) bad += check_nb_place(seed);
    return bad ? 1 : 0;
}
