// check_estimator.cpp — CPU check of the device-side velocity-estimation front end (dspmap_estimator.cuh: hash-grid
// union-find clustering, cluster order, centroids, layout of the tagged cloud) against the host implementation
// (velocity_estimator.cpp, which tests/test_host.py pins against the reference's own side thread): the same clouds through
// both, frame after frame (the cluster matching carries state), must give the same tagged cloud bit for bit.
// TEST INFRASTRUCTURE; built and run by tests/test_simt_cpu.py.
#include "simt_host.h"

#include "dspmap_kernels.cuh"

#include "est_scan.inc"

#include "dspmap_estimator.cuh"
#include "velocity_estimator.h"

#include <cstdio>
#include <random>

namespace {
struct Device {  // the buffers dspmap_create allocates, as plain host memory
    std::vector<float4> W, cfeat, cvel, SW;
    std::vector<int> parent, csize, label, pos, cbase, ccells, bbox, grank, mrank, kidx, flag_g, flag_r, hhead, tcid, cellid, cellof, cmin, ccount, rootc;
    std::vector<unsigned> cells;
    std::vector<u64> hkey;
    std::vector<int> croot, csz, spos, cdyn, coff, dseq, order, a_dyn, a_dsz, a_ssz, p_dyn, p_dsz, p_ssz;
    std::vector<int> cnt, h_hdr;
    std::vector<EstFeature> h_feat;
    std::vector<float> tagged;
    EstPtrs ep;
    unsigned hash_mask;
    explicit Device(int MP) {
        const size_t NP = (size_t)MP + 1, NCL = (size_t)MP / EST_MIN_CLUSTER + 2;
        size_t H = 1024;
        while (H < 4 * (size_t)MP) H <<= 1;
        hash_mask = (unsigned)(H - 1);
        W.resize(NP); cfeat.resize(NCL); cvel.resize(NCL); SW.resize(NP); bbox.assign(6 * NP, 0);
        for (auto *v : {&parent, &csize, &label, &pos, &cbase, &ccells, &grank, &mrank, &kidx, &flag_g, &flag_r, &tcid, &cellof, &cmin, &ccount, &rootc}) v->assign(NP, 0);
        cellid.assign(H, 0);
        for (auto *v : {&croot, &csz, &spos, &cdyn, &coff, &dseq, &order, &a_dyn, &a_dsz, &a_ssz, &p_dyn, &p_dsz, &p_ssz}) v->assign(NCL, 0);
        hkey.assign(H, ~0ull); hhead.assign(H, -1); cells.assign(NP, 0u);
        cnt.assign(EC_COUNT, 0); h_hdr.assign(EC_COUNT, 0);
        h_feat.resize(NCL);
        tagged.assign(7 * NP, -7.f);
        memset(&ep, 0, sizeof(ep));
        ep.W = W.data(); ep.parent = parent.data(); ep.csize = csize.data(); ep.label = label.data(); ep.pos = pos.data(); ep.cbase = cbase.data(); ep.ccells = ccells.data(); ep.bbox = bbox.data(); ep.SW = SW.data();
        ep.grank = grank.data(); ep.mrank = mrank.data(); ep.kidx = kidx.data(); ep.flag_g = flag_g.data(); ep.flag_r = flag_r.data();
        ep.hkey = hkey.data(); ep.hcnt = hhead.data(); ep.cells = cells.data(); ep.cellid = cellid.data(); ep.cellof = cellof.data(); ep.cmin = cmin.data(); ep.ccount = ccount.data(); ep.rootc = rootc.data();
        ep.croot = croot.data(); ep.csz = csz.data(); ep.spos = spos.data(); ep.cdyn = cdyn.data(); ep.coff = coff.data();
        ep.dseq = dseq.data(); ep.order = order.data(); ep.a_dyn = a_dyn.data(); ep.a_dsz = a_dsz.data(); ep.a_ssz = a_ssz.data();
        ep.p_dyn = p_dyn.data(); ep.p_dsz = p_dsz.data(); ep.p_ssz = p_ssz.data(); ep.cfeat = cfeat.data();
        ep.cnt = cnt.data(); ep.h_hdr = h_hdr.data(); ep.h_feat = h_feat.data();
        ep.tagged = tagged.data(); ep.tcid = tcid.data(); ep.cvel = cvel.data();
    }
};

// a camera looking along +x with a 90 x 60 degree field of view: the four outer planes in the order the estimator indexes them
// with Nh = Nv = 1 (h first, h last, v first, v last)
const float kPlanes0[12] = {0.70710678f, 0.70710678f, 0.f, -0.70710678f, 0.70710678f, 0.f, -0.5f, 0.f, 0.8660254f, 0.5f, 0.f, 0.8660254f};

struct Scene {
    std::mt19937 rng;
    struct Blob { float c[3], v[3], r; int n; };
    std::vector<Blob> blobs;
    explicit Scene(unsigned seed) : rng(seed) {
        std::uniform_real_distribution<float> U(0.f, 1.f);
        const int nb = 12 + (int)(U(rng) * 10);
        for (int b = 0; b < nb; ++b) {
            Blob B;
            B.c[0] = 2.f + 6.f * U(rng); B.c[1] = -3.f + 6.f * U(rng); B.c[2] = 0.2f + 2.2f * U(rng);  // some centroids above 1.5 m
            B.v[0] = U(rng) - 0.5f; B.v[1] = U(rng) - 0.5f; B.v[2] = 0.f;
            B.r = 0.05f + 0.3f * U(rng);
            const float u = U(rng);
            B.n = u < 0.2f ? 1 + (int)(U(rng) * 4) : (u < 0.8f ? 5 + (int)(U(rng) * 150) : 201 + (int)(U(rng) * 400));  // dropped / dynamic / static by size
            blobs.push_back(B);
        }
    }
    // sensor-frame cloud of frame f (sensor at cur, attitude q)
    void cloud(int f, const float *cur, const float *q, int n_ground, int n_wall, int n_behind, std::vector<float> &pts) {
        std::uniform_real_distribution<float> U(0.f, 1.f);
        std::normal_distribution<float> N(0.f, 1.f);
        std::vector<float> w;
        for (const auto &B : blobs)
            for (int k = 0; k < B.n; ++k)
                for (int a = 0; a < 3; ++a) w.push_back(B.c[a] + 0.1f * f * B.v[a] + B.r * 0.5f * N(rng));
        for (int k = 0; k < n_ground; ++k) { w.push_back(1.f + 8.f * U(rng)); w.push_back(-4.f + 8.f * U(rng)); w.push_back(-0.05f + 0.2f * U(rng)); }  // around the ground threshold
        for (int k = 0; k < n_wall; ++k) { w.push_back(9.f + 0.05f * N(rng)); w.push_back(-5.f + 10.f * U(rng)); w.push_back(0.2f + 3.f * U(rng)); }
        for (int k = 0; k < n_behind; ++k) { w.push_back(-1.f - 5.f * U(rng)); w.push_back(-4.f + 8.f * U(rng)); w.push_back(1.f); }
        const int n = (int)w.size() / 3;
        std::vector<int> perm(n);
        for (int i = 0; i < n; ++i) perm[i] = i;
        std::shuffle(perm.begin(), perm.end(), rng);
        float qi[4];
        dsp_quat_inverse(q, qi);
        pts.resize(3 * (size_t)n);
        for (int i = 0; i < n; ++i) {  // world -> sensor: rotate (w - cur) by the inverse attitude
            const float d[3] = {w[3 * perm[i]] - cur[0], w[3 * perm[i] + 1] - cur[1], w[3 * perm[i] + 2] - cur[2]};
            dsp_rotate(d, qi, q, &pts[3 * (size_t)i]);
        }
    }
};

template <typename K>
void run(K kernel, unsigned blocks, int threads, const EstConst &ec, const EstPtrs &ep) {
    simt::launch_grid(blocks, threads, [&] { kernel(ec, ep); });
}

int check(unsigned seed, int model, int frames) {
    Scene scene(seed);
    const int MP = 6000;
    Device dev(MP);
    VelocityEstimator host, devhost;  // the full host implementation; the host half of the device path
    host.reset(seed);
    devhost.reset(seed);
    host.filter_res = devhost.filter_res = 0.1f;
    MapConst mc;
    memset(&mc, 0, sizeof(mc));
    mc.Nh = mc.Nv = 1;
    mc.model = model;
    std::vector<float> ref_tagged, pts;
    int n_pad = 0, bad = 0;
    for (int f = 0; f < frames; ++f) {
        FrameConst fc;
        memset(&fc, 0, sizeof(fc));
        const float yaw = 0.05f * f;
        fc.q[0] = cosf(yaw / 2); fc.q[3] = sinf(yaw / 2);
        dsp_quat_inverse(fc.q, fc.qi);
        fc.cur[0] = 0.2f * f; fc.cur[1] = 0.05f * f; fc.cur[2] = 1.0f;
        fc.dt = f == 0 ? 0.f : 0.1f;
        const bool blind = f == 3;  // a frame with nothing in view: everything is kept (:1379)
        if (blind) {
            pts.assign(3 * 50, 0.f);
            for (int i = 0; i < 50; ++i) pts[3 * i] = -1.f - i;
        } else {
            scene.cloud(f, fc.cur, fc.q, 400 + 100 * f, f == 2 ? 0 : 900, 60, pts);
        }
        const int n = (int)pts.size() / 3;
        if (n > MP) { printf("scene too large\n"); return 1; }
        host.estimate(mc, fc, kPlanes0, pts.data(), n, model, ref_tagged);

        EstConst ec;
        memset(&ec, 0, sizeof(ec));
        dsp_rotate(kPlanes0, fc.q, fc.qi, ec.nrm);
        dsp_rotate(kPlanes0 + 3, fc.q, fc.qi, ec.nrm + 3);
        dsp_rotate(kPlanes0 + 6, fc.q, fc.qi, ec.nrm + 6);
        dsp_rotate(kPlanes0 + 9, fc.q, fc.qi, ec.nrm + 9);
        for (int k = 0; k < 4; ++k) { ec.q[k] = fc.q[k]; ec.qi[k] = fc.qi[k]; }
        for (int k = 0; k < 3; ++k) ec.cur[k] = fc.cur[k];
        ec.filter_res = 0.1f;
        const float tol = 2 * ec.filter_res;
        ec.tol2 = tol * tol;
        ec.inv_cell = 1.f / (tol * 0.57f);
        ec.n = n; ec.model = model; ec.hash_mask = dev.hash_mask;
        ec.n_pad_prev = n_pad; ec.nt_override = -1;
        ec.n_pad = n_pad = std::max(n_pad, n);
        EstPtrs ep = dev.ep;
        ep.pts = pts.data();
        std::fill(dev.hkey.begin(), dev.hkey.end(), ~0ull);
        std::fill(dev.hhead.begin(), dev.hhead.end(), -1);
        for (int c = 0; c < EC_PER_FRAME; ++c) dev.cnt[c] = 0;
        run(k_est_classify, 5, 256, ec, ep);
        if (model != 1) {
            run(k_est_scatter, 5, 256, ec, ep);
            run(k_est_link, 12, 256, ec, ep);
        }
        run(k_est_label, 4, 256, ec, ep);
        run(k_est_clusters, 1, 1024, ec, ep);
        run(k_est_features, model != 1 ? 5 : 1, 256, ec, ep);
        run(k_est_write, 7, 256, ec, ep);
        const int *hdr = dev.h_hdr.data();
        if (hdr[EC_NV] > 0) {
            devhost.finish_device(dev.h_feat.data(), hdr[EC_NDYN], hdr[EC_NC], fc.dt, (float *)dev.cvel.data());
            if (hdr[EC_NDYN] > 0) simt::launch_grid(4, 256, [&] { k_est_apply(ep); });
        }
        const int nt = hdr[EC_NTAGGED];
        bool ok = (size_t)nt * 7 == ref_tagged.size() && memcmp(dev.tagged.data(), ref_tagged.data(), sizeof(float) * ref_tagged.size()) == 0;
        for (int j = nt; j < n_pad && ok; ++j) ok = dev.tagged[7 * (size_t)j] >= 1e29f;
        ok = ok && host.draws == devhost.draws && host.last.size() == devhost.last.size();
        int moving = 0;
        for (int i = 0; i < nt; ++i) moving += dev.tagged[7 * (size_t)i + 3] > -100.f && dev.tagged[7 * (size_t)i + 6] > 0.01f;
        printf("seed %u model %d frame %d: %d points, %d in view, %d clusters (%d dynamic), tagged %d (%d with a velocity): %s\n", seed, model, f, n,
               hdr[EC_NV], hdr[EC_NC], hdr[EC_NDYN], nt, moving, ok ? "identical" : "DIFFERENT");
        bad += !ok;
    }
    return bad;
}
}  // namespace

int main() {
    int bad = 0;
    bad += check(1, 0, 5);
    bad += check(2, 0, 5);
    bad += check(3, 1, 4);
    return bad ? 1 : 0;
}
