// oracle/dsp_oracle.cpp — CPU restatement of the DSP map's per-frame particle loop.
//
// TEST INFRASTRUCTURE ONLY. Nothing in the product path (dsp-map_b200/, include/) includes, links or calls
// this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg load it (through
// oracle/oracle.py).  It is a sequential, runtime-configurable restatement of g-ch/DSP-map
// include/dsp_dynamic.h (and of the two variants dsp_dynamic_multiple_neighbors.h / dsp_static.h, which
// differ only in the places marked "variant:" below); every function cites the reference lines it
// follows.  Compile with -O2 -ffp-contract=off and WITHOUT -ffast-math so each fp32 operation rounds
// once, in the written order.
//
// PARITY PIN: the reference ships no tests or golden vectors, so this restatement is
// pinned against the reference itself: tests/test_oracle_vs_reference.py runs the unmodified reference
// header (oracle/_ref/libdspref_*.so, built by oracle/build_ref.py) and this file on identical
// streams and requires bit-identical particle stores, pyramid lists, observation bins, occupancy
// and future grids after every frame; tests/golden/ holds vectors produced by the reference
// for machines where /root/reference is absent.
//
// The side thread (clustering + Hungarian matching, dsp_dynamic.h:1377-1544) is NOT restated here: its
// output input_cloud_with_velocity is an explicit input of oracle_update() (world frame, 7 floats per
// point: x y z vx vy vz intensity), exactly the array getKMClusterResult() returns (dsp_dynamic.h:441).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

namespace {

struct Config {
    int32_t nx, ny, nz;
    float resolution;
    int32_t angle_resolution;
    int32_t half_fov_h, half_fov_v;
    int32_t max_ppv;
    int32_t safe_ppv;      // S (0 = derive)
    int32_t safe_pyramid;  // L (0 = derive)
    int32_t neighbor_n;    // PYRAMID_NEIGHBOR_N
    int32_t model;         // 0 dynamic, 1 static
    int32_t prediction_times;
    float future_time[8];
    float occlusion_margin;
    int32_t init_particle_num;
    float init_weight;
    uint64_t table_seed;
    uint64_t uniform_seed;
    int32_t gaussian_table_size;
    int32_t obs_max_per_pyramid;  // 100
    int32_t pi_is_double;         // mn:78 / st:74 redefine M_PIf32 as a double literal; dsp_dynamic.h keeps glibc's float one
};

enum { F_FLAG = 0, F_VX, F_VY, F_VZ, F_PX, F_PY, F_PZ, F_W, F_N };

inline uint32_t u31(uint64_t seed, uint64_t k) {  // the counter-based uniform stream (see ref_driver.cpp)
    uint64_t z = seed + (k + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 33);
}

struct Map {
    Config c;
    int V, S, P, L, T, Nh, Nv, NBW, OBS;
    float hx, hy, hz, res;
    std::vector<float> part;  // [V][S][8]   dsp_dynamic.h:116 (update-time field [8] is write-only, dropped)
    std::vector<float> vox;   // [V][4+T]    dsp_dynamic.h:120
    std::vector<int> pyr;     // [P][L][3]   dsp_dynamic.h:124
    std::vector<int> nbr;     // [P][NBW]    dsp_dynamic.h:127
    std::vector<float> obs;   // [P][OBS][5] dsp_dynamic.h:498
    std::vector<int> obs_n;   // dsp_dynamic.h:501
    std::vector<float> obs_maxlen;  // dsp_dynamic.h:515
    std::vector<float> plane_h0, plane_v0, plane_h, plane_v;  // [(Nh+1)][3], [(Nv+1)][3]
    std::vector<float> p_rand, v_rand;
    float pdf[20000];
    int64_t p_cur = 0, v_cur = 0;
    uint64_t u_cur = 0;
    float p_std = 0.2f, v_std = 0.1f, sigma_ob = 0.2f, kappa = 0.01f, P_d = 0.95f;
    float nb_weight = 0.04f;
    int nb_num = 20;
    float expected_new_born = 0.f;
    float cur_pos[3] = {0, 0, 0};
    bool have_last = false;
    float last_p[3];
    double last_t;
    bool nb_latched = false;
    int nb_min_static = 0, nb_model_generated = 0;
    std::vector<float> tagged;  // last newborn input (kept when a frame provides none, dsp_dynamic.h:1379)
    // per-frame counters
    int64_t ctr[16];
    int stage_limit = 4;  // debugging aid: 1 predict, 2 +observe, 3 +newborn, 4 all

    float *slot(int v, int s) { return &part[((size_t)v * S + s) * F_N]; }

    // dsp_dynamic.h:1551-1553
    float uniform(float lo, float hi) {
        int r = (int)u31(c.uniform_seed, u_cur++);
        return lo + static_cast<float>(r) / (static_cast<float>(RAND_MAX / (hi - lo)));
    }
    // dsp_dynamic.h:1162-1178
    float pnoise() {
        float d = p_rand[p_cur];
        if (++p_cur >= c.gaussian_table_size) p_cur = 0;
        return d;
    }
    float vnoise() {
        float d = v_rand[v_cur];
        if (++v_cur >= c.gaussian_table_size) v_cur = 0;
        return d;
    }
    // dsp_dynamic.h:1150-1160 (seeded from time(NULL) there; from table_seed here)
    void gen_tables() {
        std::default_random_engine random(c.table_seed);
        std::normal_distribution<double> n1(0, p_std);
        std::normal_distribution<double> n2(0, v_std);
        p_rand.resize(c.gaussian_table_size);
        v_rand.resize(c.gaussian_table_size);
        for (int i = 0; i < c.gaussian_table_size; i++) {
            p_rand[i] = n1(random);
            v_rand[i] = n2(random);
        }
    }
    // dsp_dynamic.h:1282-1292: note the normaliser uses pi/2 (M_PI_2f32), i.e. 1/sqrt(pi)
    void gen_pdf() {
        for (int i = 0; i < 20000; ++i) {
            float x = (float)(i - 10000) * 0.001f;
            pdf[i] = (1.f / (sqrtf(2.f * 1.57079632679489661923))) * expf(-powf(x, 2) / (2));
        }
    }
    // dsp_dynamic.h:1294-1301
    float query_pdf(float x, float mu, float sigma) const {
        float cx = (x - mu) / sigma;
        if (cx > 9.9f) cx = 9.9f;
        else if (cx < -9.9f) cx = -9.9f;
        return pdf[(int)(cx * 1000 + 10000)];
    }
    // dsp_dynamic.h:1118-1125
    bool is_out(float x, float y, float z) const {
        return x >= hx || x <= -hx || y >= hy || y <= -hy || z >= hz || z <= -hz;
    }
    // dsp_dynamic.h:1076-1088
    bool voxel_index(float x, float y, float z, int &idx) const {
        if (is_out(x, y, z)) return false;
        int ix = (int)((x + hx) / res), iy = (int)((y + hy) / res), iz = (int)((z + hz) / res);
        idx = iz * c.ny * c.nx + iy * c.nx + ix;
        return !(idx < 0 || idx >= V);
    }
    // dsp_dynamic.h:1090-1107
    void voxel_center(int idx, float *o) const {
        int zs = c.ny * c.nx, iz = idx / zs, rem = idx - iz * zs, iy = rem / c.nx, ix = rem - iy * c.nx;
        float cx = -hx + res * 0.5f, cy = -hy + res * 0.5f, cz = -hz + res * 0.5f;
        o[0] = (float)ix * res + cx;
        o[1] = (float)iy * res + cy;
        o[2] = (float)iz * res + cz;
    }
    // dsp_dynamic.h:1303-1322 with the scalar quaternion arithmetic of oracle/shim/Eigen/Eigen
    static void rotate(const float *v, const float *q, float *o) {
        float aw = q[0], ax = q[1], ay = q[2], az = q[3];
        float bw = 0.f, bx = v[0], by = v[1], bz = v[2];
        float tw = aw * bw - ax * bx - ay * by - az * bz;
        float tx = aw * bx + ax * bw + ay * bz - az * by;
        float ty = aw * by + ay * bw + az * bx - ax * bz;
        float tz = aw * bz + az * bw + ax * by - ay * bx;
        float n2 = ax * ax + ay * ay + az * az + aw * aw;
        float iw, ix, iy, iz;
        if (n2 > 0.f) { iw = aw / n2; ix = -ax / n2; iy = -ay / n2; iz = -az / n2; }
        else { iw = ix = iy = iz = 0.f; }
        o[0] = tw * ix + tx * iw + ty * iz - tz * iy;
        o[1] = tw * iy + ty * iw + tz * ix - tx * iz;
        o[2] = tw * iz + tz * iw + tx * iy - ty * ix;
    }
    static float dot(const float *n, float x, float y, float z) { return x * n[0] + y * n[1] + z * n[2]; }  // :1324
    // dsp_dynamic.h:1329-1339
    bool in_fov(float x, float y, float z) const {
        return dot(&plane_h[0], x, y, z) >= 0.f && dot(&plane_h[3 * Nh], x, y, z) <= 0.f &&
               dot(&plane_v[0], x, y, z) <= 0.f && dot(&plane_v[3 * Nv], x, y, z) >= 0.f;
    }
    // dsp_dynamic.h:1341-1353 / 1355-1367
    int pyr_h(float x, float y, float z) const {
        float last = 1.f;
        for (int i = 0; i < Nh; i++) {
            float d = dot(&plane_h[3 * (i + 1)], x, y, z);
            if (last * d <= 0.f) return i;
            last = d;
        }
        return -1;
    }
    int pyr_v(float x, float y, float z) const {
        float last = -1.f;
        for (int j = 0; j < Nv; j++) {
            float d = dot(&plane_v[3 * (j + 1)], x, y, z);
            if (last * d <= 0.f) return j;
            last = d;
        }
        return -1;
    }

    // dsp_dynamic.h:145-175, 525-591 (+ dsp_static.h:63 for S)
    void init(const Config &cfg) {
        c = cfg;
        V = c.nx * c.ny * c.nz;
        res = c.resolution;
        hx = (res * (float)c.nx) * 0.5f;
        hy = (res * (float)c.ny) * 0.5f;
        hz = (res * (float)c.nz) * 0.5f;
        Nh = c.half_fov_h * 2 / c.angle_resolution;
        Nv = c.half_fov_v * 2 / c.angle_resolution;
        P = Nh * Nv;
        int pyramid_num = 360 * 180 / c.angle_resolution / c.angle_resolution;
        int safe_particle_num = V * c.max_ppv + 1e5;
        S = c.safe_ppv > 0 ? c.safe_ppv : c.max_ppv * (c.model == 1 ? 5 : 2);
        L = c.safe_pyramid > 0 ? c.safe_pyramid : safe_particle_num / pyramid_num * 2;
        T = c.prediction_times;
        OBS = c.obs_max_per_pyramid;
        NBW = (2 * c.neighbor_n + 1) * (2 * c.neighbor_n + 1) + 1;
        part.assign((size_t)V * S * F_N, 0.f);
        vox.assign((size_t)V * (4 + T), 0.f);
        pyr.assign((size_t)P * L * 3, 0);
        nbr.assign((size_t)P * NBW, 0);
        obs.assign((size_t)P * OBS * 5, 0.f);
        obs_n.assign(P, 0);
        obs_maxlen.assign(P, 0.f);
        // :543 (dsp_dynamic.h: M_PIf32 is glibc's float literal, fp32 product; mn / static: their own double literal)
        float ang = c.pi_is_double ? (float)((float)c.angle_resolution / 180.f * 3.14159265358979323846) : (float)c.angle_resolution / 180.f * 3.14159265358979323846f;
        plane_h0.resize(3 * (Nh + 1));
        plane_v0.resize(3 * (Nv + 1));
        plane_h = plane_h0;
        plane_v = plane_v0;
        int h0 = -c.half_fov_h / c.angle_resolution, h1 = -h0;  // :564-570
        for (int i = h0; i <= h1; i++) {
            plane_h0[3 * (i + h1) + 0] = -std::sin((float)i * ang);
            plane_h0[3 * (i + h1) + 1] = std::cos((float)i * ang);
            plane_h0[3 * (i + h1) + 2] = 0.f;
        }
        int v0 = -c.half_fov_v / c.angle_resolution, v1 = -v0;  // :572-578
        for (int i = v0; i <= v1; i++) {
            plane_v0[3 * (i + v1) + 0] = std::sin((float)i * ang);
            plane_v0[3 * (i + v1) + 1] = 0.f;
            plane_v0[3 * (i + v1) + 2] = std::cos((float)i * ang);
        }
        for (int p = 0; p < P; p++) {  // :1128-1147 (mn:1135-1136 widens the block)
            int h = p / Nv, v = p % Nv, n = 0;
            for (int i = -c.neighbor_n; i <= c.neighbor_n; ++i)
                for (int j = -c.neighbor_n; j <= c.neighbor_n; ++j) {
                    int hh = h + i, vv = v + j;
                    if (hh >= 0 && hh < Nh && vv >= 0 && vv < Nv) nbr[(size_t)p * NBW + 1 + n++] = hh * Nv + vv;
                }
            nbr[(size_t)p * NBW] = n;
        }
        gen_tables();
        gen_pdf();
        add_random_particles(c.init_particle_num, c.init_weight);
        std::memset(ctr, 0, sizeof(ctr));
    }
    // dsp_dynamic.h:1183-1201
    int add_particle(int v, float px, float py, float pz, float vx, float vy, float vz, float w) {
        for (int i = 0; i < S; i++) {
            float *q = slot(v, i);
            if (q[F_FLAG] < 0.1f) {
                q[F_FLAG] = 15.f;
                q[F_VX] = vx; q[F_VY] = vy; q[F_VZ] = vz;
                q[F_PX] = px; q[F_PY] = py; q[F_PZ] = pz;
                q[F_W] = w;
                return 1;
            }
        }
        return 0;
    }
    // dsp_dynamic.h:594-624
    void add_random_particles(int n, float w) {
        for (int i = 0; i < n; i++) {
            float px = uniform(-hx, hx), py = uniform(-hy, hy), pz = uniform(-hz, hz);
            float vx = uniform(-1.f, 1.f), vy = uniform(-1.f, 1.f), vz = uniform(-1.f, 1.f);
            int idx;
            if (voxel_index(px, py, pz, idx)) add_particle(idx, px, py, pz, vx, vy, vz, w);
        }
    }

    // dsp_dynamic.h:1206-1274
    int move_particle(int nv, int cv, int cs) {
        int ns = cs;
        float *src = slot(cv, cs);
        if (nv != cv) {
            src[F_FLAG] = 0.f;
            bool ok = false;
            for (int i = 0; i < S; ++i) {
                float *d = slot(nv, i);
                if (d[F_FLAG] < 0.1f) {
                    ns = i;
                    ok = true;
                    d[F_FLAG] = 7.f;
                    for (int k = 1; k < F_N; ++k) d[k] = src[k];
                    break;
                }
            }
            if (!ok) return -1;
        }
        float *q = slot(nv, ns);
        if (in_fov(q[F_PX], q[F_PY], q[F_PZ])) {
            int h = pyr_h(q[F_PX], q[F_PY], q[F_PZ]), v = pyr_v(q[F_PX], q[F_PY], q[F_PZ]);
            int pid = h * Nv + v;
            bool ok = false;
            for (int j = 0; j < L; j++) {
                int *e = &pyr[((size_t)pid * L + j) * 3];
                if (e[0] == 0) {
                    e[0] |= 1;
                    e[1] = nv;
                    e[2] = ns;
                    ok = true;
                    break;
                }
            }
            if (!ok) {
                q[F_FLAG] = 0.f;
                return -2;
            }
            if (!(std::fabs(q[F_VX] * q[F_VY] * q[F_VZ]) < 1e-6)) {  // :1262-1269
                q[F_VX] += vnoise();
                q[F_VY] += vnoise();
                q[F_VZ] = 0.f;
            }
        }
        return 1;
    }

    // dsp_dynamic.h:627-701; variant: dsp_static.h:640-646 zeroes v and applies only the ego-motion shift
    void predict(float ox, float oy, float oz, float dt) {
        for (size_t i = 0; i < (size_t)P * L; ++i) pyr[i * 3] &= 0;  // :637-642
        for (int v = 0; v < V; ++v)
            for (int s = 0; s < S; s++) {
                float *q = slot(v, s);
                if (!(q[F_FLAG] > 0.1f && q[F_FLAG] < 6.f)) continue;
                q[F_FLAG] = 1.f;
                ++ctr[0];  // N_in
                if (c.model == 1) {
                    q[F_VX] = 0.f; q[F_VY] = 0.f; q[F_VZ] = 0.f;
                    q[F_PX] += ox; q[F_PY] += oy; q[F_PZ] += oz;
                } else {
                    if (!(std::fabs(q[F_VX] * q[F_VY] * q[F_VZ]) < 1e-6)) {  // :653-659
                        q[F_VX] += vnoise();
                        q[F_VY] += vnoise();
                        q[F_VZ] += vnoise();
                    }
                    q[F_VZ] = 0.f;  // LIMIT_MOVEMENT_IN_XY_PLANE 1 (:44, :661-663)
                    q[F_PX] += dt * q[F_VX] + ox;
                    q[F_PY] += dt * q[F_VY] + oy;
                    q[F_PZ] += dt * q[F_VZ] + oz;
                }
                int nv;
                if (voxel_index(q[F_PX], q[F_PY], q[F_PZ], nv)) {
                    int r = move_particle(nv, v, s);
                    if (r == -2) ++ctr[3];       // pyramid full
                    else if (r == -1) ++ctr[2];  // voxel full
                    else if (nv != v) ++ctr[4];  // moved to another voxel
                } else {
                    q[F_FLAG] = 0.f;
                    ++ctr[1];  // left the map
                }
            }
    }

    // dsp_dynamic.h:704-793
    void observe_update() {
        for (int i = 0; i < P; ++i)
            for (int j = 0; j < obs_n[i]; ++j) {
                float *z = &obs[((size_t)i * OBS + j) * 5];
                for (int n = 0; n < nbr[(size_t)i * NBW]; ++n) {
                    int pc = nbr[(size_t)i * NBW + 1 + n];
                    for (int k = 0; k < L; ++k) {
                        const int *e = &pyr[((size_t)pc * L + k) * 3];
                        if (!(e[0] & 1)) continue;
                        const float *q = slot(e[1], e[2]);
                        float gk = query_pdf(q[F_PX], z[0], sigma_ob) * query_pdf(q[F_PY], z[1], sigma_ob) *
                                   query_pdf(q[F_PZ], z[2], sigma_ob);
                        z[3] += P_d * q[F_W] * gk;
                    }
                }
                z[3] += (expected_new_born + kappa);
            }
        for (int i = 0; i < P; i++)
            for (int k = 0; k < L; k++) {
                const int *e = &pyr[((size_t)i * L + k) * 3];
                if (!(e[0] & 1)) continue;
                ++ctr[5];  // N_fov
                float *q = slot(e[1], e[2]);
                float px = q[F_PX], py = q[F_PY], pz = q[F_PZ];
                float dist = sqrtf(px * px + py * py + pz * pz);
                if (obs_maxlen[i] > 0.f && dist > obs_maxlen[i] + c.occlusion_margin) continue;  // :761 / mn:761
                float sum = 0.f;
                for (int n = 0; n < nbr[(size_t)i * NBW]; ++n) {
                    int ni = nbr[(size_t)i * NBW + 1 + n];
                    for (int zs = 0; zs < obs_n[ni]; ++zs) {
                        float *z = &obs[((size_t)ni * OBS + zs) * 5];
                        float gk = query_pdf(px, z[0], sigma_ob) * query_pdf(py, z[1], sigma_ob) *
                                   query_pdf(pz, z[2], sigma_ob);
                        sum += P_d * gk / z[3];
                    }
                }
                q[F_W] *= ((1 - P_d) + sum);
            }
    }

    // dsp_dynamic.h:796-921; variant: dsp_static.h:779-829 (no in-map test of the point, no split, v = 0)
    void newborn() {
        float norm = 0.f;
        for (int i = 0; i < P; i++)
            for (int j = 0; j < obs_n[i]; j++) norm += 1.f / obs[((size_t)i * OBS + j) * 5 + 3];
        float w_new = nb_weight * norm;
        if (!nb_latched) {  // function-local statics, frozen at the first call (:808-811; static: 0.2f)
            nb_min_static = (int)((float)nb_num * (c.model == 1 ? 0.2f : 0.15f));
            nb_model_generated = (int)((float)nb_num * 0.8f);
            nb_latched = true;
        }
        size_t n_pts = tagged.size() / 7;
        for (size_t m = 0; m < n_pts; ++m) {
            const float *pt = &tagged[7 * m];
            float cx = pt[0] - cur_pos[0], cy = pt[1] - cur_pos[1], cz = pt[2] - cur_pos[2];
            int n_static = 0;
            if (c.model == 0) {
                int pv;
                float ws = 0.f, wd = 0.f, wsd = 0.f;
                if (!voxel_index(cx, cy, cz, pv)) continue;
                for (int k = 0; k < S; ++k) {
                    const float *q = slot(pv, k);
                    if (q[F_FLAG] > 0.9f && q[F_FLAG] < 14.f) {
                        float va = std::fabs(q[F_VX]) + std::fabs(q[F_VY]) + std::fabs(q[F_VZ]);
                        if (va < 0.1f) ws += q[F_W];
                        else if (va < 0.5f) wsd += q[F_W];
                        else wd += q[F_W];
                    }
                }
                float tot = ws + wd + wsd;  // :851-866
                float m_s = ws / tot, m_d = wd / tot, m_sd = wsd / tot;
                float p_s = (m_s + m_s + m_sd) * 0.5f, p_d = (m_d + m_d + m_sd) * 0.5f;
                float np_ = p_s + p_d;
                float ps_n = p_s / np_;
                float prod = (float)nb_model_generated * ps_n;
                n_static = (prod != prod) ? INT_MIN : (int)prod;  // (int)NaN is INT_MIN on x86 (cvttss2si)
                n_static = std::max(nb_min_static, n_static);
            }
            for (int p = 0; p < nb_num; p++) {
                float px = cx + pnoise(), py = cy + pnoise(), pz = cz + pnoise();
                int idx;
                if (!voxel_index(px, py, pz, idx)) continue;
                float vx = 0.f, vy = 0.f, vz = 0.f;
                if (c.model == 0) {
                    if (p < n_static) {
                    } else if (pt[3] > -100.f && p < nb_model_generated) {
                        if (pt[6] > 0.01f) {
                            vx = pt[3] + 4 * vnoise();
                            vy = pt[4] + 4 * vnoise();
                            vz = pt[5] + 4 * vnoise();
                        }
                    } else if (pt[6] > 0.01f) {
                        vx = uniform(-1.5f, 1.5f);
                        vy = uniform(-1.5f, 1.5f);
                        vz = uniform(-0.5f, 0.5f);
                    }
                    vz = 0.f;
                }
                ++ctr[6];  // candidates inside the map
                if (add_particle(idx, px, py, pz, vx, vy, vz, w_new)) ++ctr[7];  // N_born
            }
        }
    }

    // dsp_dynamic.h:924-1057
    void resample() {
        for (int v = 0; v < V; ++v) {
            float wsum = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
            int n = 0, n_old = 0;
            for (int s = 0; s < S; s++) {
                float *q = slot(v, s);
                if (!(q[F_FLAG] > 0.1f)) continue;
                if (q[F_W] < 1e-3) {
                    q[F_FLAG] = 0.f;
                    ++ctr[8];  // dropped for low weight
                    continue;
                }
                if (q[F_FLAG] < 10.f) {
                    ++n_old;
                    sx += q[F_VX]; sy += q[F_VY]; sz += q[F_VZ];
                    for (int t = 0; t < T; ++t) {
                        float ft = c.future_time[t];
                        float fx = q[F_PX] + q[F_VX] * ft, fy = q[F_PY] + q[F_VY] * ft, fz = q[F_PZ] + q[F_VZ] * ft;
                        int fi;
                        if (voxel_index(fx, fy, fz, fi)) vox[(size_t)fi * (4 + T) + 4 + t] += q[F_W];
                    }
                }
                q[F_FLAG] = 1.f;
                ++n;
                wsum += q[F_W];
            }
            ctr[9] += n;       // N_pre
            ctr[10] += n_old;  // N_old
            float *o = &vox[(size_t)v * (4 + T)];
            o[0] = wsum;
            if (n_old > 0) { o[1] = sx / (float)n_old; o[2] = sy / (float)n_old; o[3] = sz / (float)n_old; }
            else { o[1] = o[2] = o[3] = 0.f; }
            if (n < 5) { ctr[11] += n; continue; }
            int n_after = n > c.max_ppv ? c.max_ppv : n;
            float w_after = wsum / (float)n_after;
            float acc_ori = 0.f, acc_new = w_after * 0.5f;
            for (int s = 0; s < S; ++s) {
                float *q = slot(v, s);
                if (!(q[F_FLAG] > 0.7f)) continue;
                acc_ori += q[F_W];
                if (acc_ori > acc_new) {
                    q[F_W] = w_after;
                    acc_new += w_after;
                    bool full = false;
                    int pi = 0;
                    while (acc_ori > acc_new) {
                        bool found = false;
                        if (!full)
                            for (; pi < S; ++pi) {
                                float *d = slot(v, pi);
                                if (d[F_FLAG] < 0.1f) {
                                    d[F_FLAG] = 0.6f;
                                    for (int k = 1; k < F_N; k++) d[k] = q[k];
                                    found = true;
                                    break;
                                }
                            }
                        if (!found) {
                            q[F_W] += w_after;
                            full = true;
                        }
                        acc_new += w_after;
                    }
                } else {
                    q[F_FLAG] = 0.f;
                }
            }
            for (int s = 0; s < S; ++s)
                if (slot(v, s)[F_FLAG] > 0.1f) ++ctr[11];  // N_out
        }
    }

    // dsp_dynamic.h:181-353 (without the side thread and the CSV dump)
    int update(int n, int stride, const float *pts, float px, float py, float pz, double t, const float *q,
               const float *tag, int n_tag) {
        if (!have_last) {  // function-local statics initialised from the first call's arguments (:187-190)
            last_p[0] = px; last_p[1] = py; last_p[2] = pz;
            last_t = t;
            have_last = true;
        }
        if (std::fabs(q[0]) > 1.001f || std::fabs(q[1]) > 1.001f || std::fabs(q[2]) > 1.001f || std::fabs(q[3]) > 1.001f)
            return 0;
        float ox = px - last_p[0], oy = py - last_p[1], oz = pz - last_p[2];
        float dt = (float)(t - last_t);
        if (std::fabs(ox) > 10.f || std::fabs(oy) > 10.f || std::fabs(oz) > 10.f || dt < 0.f || dt > 10.f) return 0;
        std::memset(ctr, 0, sizeof(ctr));
        cur_pos[0] = last_p[0] = px;
        cur_pos[1] = last_p[1] = py;
        cur_pos[2] = last_p[2] = pz;
        last_t = t;
        for (int i = 0; i < Nh + 1; i++) rotate(&plane_h0[3 * i], q, &plane_h[3 * i]);  // :226-232
        for (int j = 0; j < Nv + 1; j++) rotate(&plane_v0[3 * j], q, &plane_v[3 * j]);
        for (int i = 0; i < P; i++) { obs_n[i] = 0; obs_maxlen[i] = -1.f; }  // :235-238
        int valid = 0;
        for (int k = 0; k < n; ++k) {  // :244-290
            float r[3];
            rotate(pts + (size_t)k * stride, q, r);
            if (!in_fov(r[0], r[1], r[2])) continue;
            int pid = pyr_h(r[0], r[1], r[2]) * Nv + pyr_v(r[0], r[1], r[2]);
            int j = obs_n[pid];
            float len = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
            float *z = &obs[((size_t)pid * OBS + j) * 5];
            z[0] = r[0]; z[1] = r[1]; z[2] = r[2]; z[3] = 0.f; z[4] = len;
            if (obs_maxlen[pid] < len) obs_maxlen[pid] = len;
            obs_n[pid] += 1;
            if (obs_n[pid] >= OBS) obs_n[pid] = OBS - 1;  // overflow: the last slot is overwritten (:281-284)
            ++valid;
        }
        ctr[12] = valid;
        expected_new_born = nb_weight * (float)valid * (float)nb_num;  // :292
        if (tag && n_tag >= 0) tagged.assign(tag, tag + (size_t)7 * n_tag);  // null = keep the previous cloud (:1379)
        predict(-ox, -oy, -oz, dt);                 // :300
        if (stage_limit >= 2) observe_update();     // :304
        if (stage_limit >= 3) newborn();            // :315
        if (stage_limit >= 4) resample();           // :322
        return 1;
    }

    // dsp_dynamic.h:385-426
    int occupancy(float thr, float *xyz, int cap, float *future) {
        int n = 0;
        for (int i = 0; i < V; i++) {
            float *o = &vox[(size_t)i * (4 + T)];
            if (o[0] > thr) {
                if (xyz && n < cap) voxel_center(i, xyz + 3 * n);
                ++n;
            }
            for (int t = 0; t < T; ++t) {
                if (future) future[(size_t)i * T + t] = o[4 + t];
                o[4 + t] = 0.f;
            }
        }
        return n;
    }
};

}  // namespace

extern "C" {
void *oracle_create(const Config *c) {
    Map *m = new Map();
    m->init(*c);
    return m;
}
void oracle_destroy(void *h) { delete (Map *)h; }
void oracle_dims(void *h, int *d) {
    Map *m = (Map *)h;
    int v[] = {m->V, m->S, m->P, m->L, m->T, m->Nh, m->Nv, m->NBW, m->c.max_ppv, m->c.nx, m->c.ny, m->c.nz, m->OBS, m->c.model};
    std::memcpy(d, v, sizeof(v));
}
void oracle_set_prediction_variance(void *h, float p, float v) {  // dsp_dynamic.h:355-360
    Map *m = (Map *)h;
    m->p_std = p;
    m->v_std = v;
    m->gen_tables();
}
void oracle_set_observation_stddev(void *h, float s) { ((Map *)h)->sigma_ob = s; }
void oracle_set_newborn_weight(void *h, float w) { ((Map *)h)->nb_weight = w; }
void oracle_set_newborn_number(void *h, int n) { ((Map *)h)->nb_num = n; }
int oracle_update(void *h, int n, int stride, const float *pts, float px, float py, float pz, double t, float qw,
                  float qx, float qy, float qz, const float *tagged, int n_tagged) {
    float q[4] = {qw, qx, qy, qz};
    return ((Map *)h)->update(n, stride, pts, px, py, pz, t, q, tagged, n_tagged);
}
int oracle_get_occupancy(void *h, float thr, float *xyz, int cap, float *future) {
    return ((Map *)h)->occupancy(thr, xyz, cap, future);
}
void oracle_clear_prediction(void *h) {  // dsp_dynamic.h:431-438
    Map *m = (Map *)h;
    for (int i = 0; i < m->V; i++)
        for (int t = 0; t < m->T; ++t) m->vox[(size_t)i * (4 + m->T) + 4 + t] = 0.f;
}
int oracle_dump_particles(void *h, int *ids, float *vals, int cap) {
    Map *m = (Map *)h;
    int n = 0;
    for (int v = 0; v < m->V; ++v)
        for (int s = 0; s < m->S; ++s) {
            const float *q = m->slot(v, s);
            if (q[F_FLAG] > 0.1f) {
                if (ids && n < cap) {
                    ids[2 * n] = v;
                    ids[2 * n + 1] = s;
                    std::memcpy(vals + 8 * n, q, 8 * sizeof(float));
                }
                ++n;
            }
        }
    return n;
}
// state injection: replaces the particle store (ids[n][2], vals[n][8] as dumped) — tests only
void oracle_load_particles(void *h, const int *ids, const float *vals, int n) {
    Map *m = (Map *)h;
    std::fill(m->part.begin(), m->part.end(), 0.f);
    for (int i = 0; i < n; ++i) std::memcpy(m->slot(ids[2 * i], ids[2 * i + 1]), vals + 8 * i, 8 * sizeof(float));
}
void oracle_dump_voxel_objects(void *h, float *out) {
    Map *m = (Map *)h;
    std::memcpy(out, m->vox.data(), m->vox.size() * sizeof(float));
}
void oracle_dump_observations(void *h, int *counts, float *maxlen, float *pts) {
    Map *m = (Map *)h;
    std::memcpy(counts, m->obs_n.data(), sizeof(int) * m->P);
    std::memcpy(maxlen, m->obs_maxlen.data(), sizeof(float) * m->P);
    if (pts) std::memcpy(pts, m->obs.data(), m->obs.size() * sizeof(float));
}
int oracle_dump_pyramid_lists(void *h, int *offsets, int *entries, int cap) {
    Map *m = (Map *)h;
    int n = 0;
    for (int p = 0; p < m->P; ++p) {
        offsets[p] = n;
        for (int j = 0; j < m->L; ++j) {
            const int *e = &m->pyr[((size_t)p * m->L + j) * 3];
            if (e[0] & 1) {
                if (entries && n < cap) { entries[2 * n] = e[1]; entries[2 * n + 1] = e[2]; }
                ++n;
            }
        }
    }
    offsets[m->P] = n;
    return n;
}
void oracle_dump_neighbors(void *h, int *out) {
    Map *m = (Map *)h;
    std::memcpy(out, m->nbr.data(), m->nbr.size() * sizeof(int));
}
void oracle_cursors(void *h, int64_t *c) {
    Map *m = (Map *)h;
    c[0] = m->p_cur;
    c[1] = m->v_cur;
    c[2] = (int64_t)m->u_cur;
    c[3] = 0;
}
void oracle_set_cursors(void *h, int64_t p, int64_t v, int64_t u) {
    Map *m = (Map *)h;
    m->p_cur = p;
    m->v_cur = v;
    m->u_cur = (uint64_t)u;
}
void oracle_set_last_pose(void *h, float px, float py, float pz, double t) {
    Map *m = (Map *)h;
    m->last_p[0] = px; m->last_p[1] = py; m->last_p[2] = pz;
    m->last_t = t;
    m->have_last = true;
}
void oracle_set_stage_limit(void *h, int k) { ((Map *)h)->stage_limit = k; }
void oracle_counters(void *h, int64_t *out) { std::memcpy(out, ((Map *)h)->ctr, sizeof(int64_t) * 16); }
void oracle_gaussian_tables(void *h, float *p, float *v, int n) {
    Map *m = (Map *)h;
    std::memcpy(p, m->p_rand.data(), sizeof(float) * n);
    std::memcpy(v, m->v_rand.data(), sizeof(float) * n);
}
void oracle_pdf_table(void *h, float *out) { std::memcpy(out, ((Map *)h)->pdf, sizeof(float) * 20000); }
void oracle_plane_normals(void *h, float *ph, float *pv) {
    Map *m = (Map *)h;
    std::memcpy(ph, m->plane_h.data(), m->plane_h.size() * sizeof(float));
    std::memcpy(pv, m->plane_v.data(), m->plane_v.size() * sizeof(float));
}
int oracle_voxel_index(void *h, float x, float y, float z) {
    int idx = -1;
    return ((Map *)h)->voxel_index(x, y, z, idx) ? idx : -1;
}
void oracle_voxel_center(void *h, int idx, float *xyz) { ((Map *)h)->voxel_center(idx, xyz); }
}
