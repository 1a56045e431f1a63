// velocity_estimator.cpp — see velocity_estimator.h.  Compiled with -ffp-contract=off: the fp32 expressions below are
// evaluated in the order the reference writes them.
#include "velocity_estimator.h"

#include <algorithm>
#include <cmath>
#include <limits>

#include "dspmap_hostmath.h"

float VelocityEstimator::uniform(float lo, float hi) { return dsp_uniform(seed, draws++, lo, hi); }

namespace {
struct Grid {  // open-addressing hash from integer cell to the head of a linked list of points
    std::vector<uint64_t> keys;
    std::vector<int> head, next;
    uint64_t mask;
    static uint64_t pack(int64_t a, int64_t b, int64_t c) {
        return ((uint64_t)((a + (1 << 20)) & 0x1FFFFF) << 42) | ((uint64_t)((b + (1 << 20)) & 0x1FFFFF) << 21) |
               (uint64_t)((c + (1 << 20)) & 0x1FFFFF);
    }
    static uint64_t mix(uint64_t k) {
        k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33;
        return k;
    }
    void build(const std::vector<int64_t> &cells, int n) {
        size_t cap = 16;
        while (cap < (size_t)n * 2) cap <<= 1;
        mask = cap - 1;
        keys.assign(cap, ~0ull);
        head.assign(cap, -1);
        next.assign(n, -1);
        for (int i = n - 1; i >= 0; --i) {  // reverse, so each list ascends by index
            uint64_t k = pack(cells[3 * i], cells[3 * i + 1], cells[3 * i + 2]);
            size_t h = mix(k) & mask;
            while (keys[h] != ~0ull && keys[h] != k) h = (h + 1) & mask;
            keys[h] = k;
            next[i] = head[h];
            head[h] = i;
        }
    }
    int find(int64_t a, int64_t b, int64_t c) const {
        uint64_t k = pack(a, b, c);
        size_t h = mix(k) & mask;
        while (keys[h] != ~0ull) {
            if (keys[h] == k) return head[h];
            h = (h + 1) & mask;
        }
        return -1;
    }
};
}  // namespace

void euclidean_clusters(const float *xyz, int n, float tol, int min_size, int max_size, std::vector<std::vector<int>> &out) {
    // Connected components of the graph "d2 <= tol^2" through a union-find over grid cells.  The cell edge is below
    // tol / sqrt(3), so all points of one cell are mutually linked; links between cells up to two cells apart need one
    // witness pair.  Cost is linear in the points for the dense clouds a nearby surface produces (a per-point
    // neighbourhood search is quadratic there).  Components do not depend on the traversal, so the result is identical.
    out.clear();
    if (n == 0 || !(tol > 0.f)) return;
    const float tol2 = tol * tol;
    const float edge = tol * 0.57f;
    std::vector<int64_t> cells(3 * (size_t)n);
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) cells[3 * i + k] = (int64_t)std::floor(xyz[3 * i + k] / edge);
    Grid g;
    g.build(cells, n);
    // one representative point per occupied cell (the head of its list), cells numbered in order of first appearance
    std::vector<int> cell_of(n, -1), rep;
    for (int i = 0; i < n; ++i) {
        const int h = g.find(cells[3 * i], cells[3 * i + 1], cells[3 * i + 2]);  // smallest index in the cell
        if (h == i) { cell_of[i] = (int)rep.size(); rep.push_back(i); }
    }
    for (int i = 0; i < n; ++i)
        if (cell_of[i] < 0) cell_of[i] = cell_of[g.find(cells[3 * i], cells[3 * i + 1], cells[3 * i + 2])];
    const int nc = (int)rep.size();
    std::vector<int> parent(nc);
    for (int c = 0; c < nc; ++c) parent[c] = c;
    auto root = [&](int c) {
        while (parent[c] != c) { parent[c] = parent[parent[c]]; c = parent[c]; }
        return c;
    };
    auto linked = [&](int ha, int hb) {  // is there a pair (i in cell a, j in cell b) within tol?
        for (int i = ha; i >= 0; i = g.next[i]) {
            const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
            for (int j = hb; j >= 0; j = g.next[j]) {
                const float dx = xyz[3 * j] - px, dy = xyz[3 * j + 1] - py, dz = xyz[3 * j + 2] - pz;
                if (dx * dx + dy * dy + dz * dz <= tol2) return true;
            }
        }
        return false;
    };
    for (int pass = 1; pass <= 2; ++pass)  // adjacent cells first: most far links are then already implied
        for (int c = 0; c < nc; ++c) {
            const int i = rep[c];
            const int64_t cx = cells[3 * i], cy = cells[3 * i + 1], cz = cells[3 * i + 2];
            for (int64_t a = -2; a <= 2; ++a)
                for (int64_t b = -2; b <= 2; ++b)
                    for (int64_t d = -2; d <= 2; ++d) {
                        const int64_t m = std::max(std::max(a < 0 ? -a : a, b < 0 ? -b : b), d < 0 ? -d : d);
                        if (m != pass) continue;
                        const int hb = g.find(cx + a, cy + b, cz + d);
                        if (hb < 0) continue;
                        const int cb = cell_of[hb];
                        if (cb < c) continue;  // each unordered pair once
                        int ra = root(c), rb = root(cb);
                        if (ra == rb) continue;
                        if (linked(i, hb)) parent[std::max(ra, rb)] = std::min(ra, rb);
                    }
        }
    // gather components; a component is discovered at its smallest point index, like a seeded flood fill would
    std::vector<int> comp_of_root(nc, -1);
    std::vector<std::vector<int>> comps;
    for (int i = 0; i < n; ++i) {
        const int r = root(cell_of[i]);
        if (comp_of_root[r] < 0) { comp_of_root[r] = (int)comps.size(); comps.emplace_back(); }
        comps[comp_of_root[r]].push_back(i);  // ascending indices by construction
    }
    for (auto &c : comps)
        if ((int)c.size() >= min_size && (int)c.size() <= max_size) out.push_back(std::move(c));
    std::stable_sort(out.begin(), out.end(), [](const std::vector<int> &a, const std::vector<int> &b) { return a.size() > b.size(); });
}

void hungarian(const std::vector<float> &cost, int R, int C, std::vector<int> &assign) {
    assign.assign(R, -1);
    const int n = std::max(R, C);
    if (n == 0) return;
    double mx = 0;
    for (int i = 0; i < R * C; ++i) mx = std::max(mx, (double)cost[i]);
    const int W = n + 1;
    std::vector<double> a((size_t)W * W, mx);
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) a[(size_t)(r + 1) * W + (c + 1)] = (double)cost[(size_t)r * C + c];
    const double INF = std::numeric_limits<double>::infinity();
    std::vector<double> u(W, 0), v(W, 0), minv(W);
    std::vector<int> p(W, 0), way(W, 0);
    std::vector<char> used(W);
    for (int i = 1; i <= n; ++i) {
        p[0] = i;
        int j0 = 0;
        std::fill(minv.begin(), minv.end(), INF);
        std::fill(used.begin(), used.end(), 0);
        do {
            used[j0] = 1;
            int i0 = p[j0], j1 = 0;
            double delta = INF;
            for (int j = 1; j <= n; ++j)
                if (!used[j]) {
                    double cur = a[(size_t)i0 * W + j] - u[i0] - v[j];
                    if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
                    if (minv[j] < delta) { delta = minv[j]; j1 = j; }
                }
            for (int j = 0; j <= n; ++j)
                if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
                else minv[j] -= delta;
            j0 = j1;
        } while (p[j0] != 0);
        do { int j1 = way[j0]; p[j0] = p[j1]; j0 = j1; } while (j0);
    }
    for (int j = 1; j <= n; ++j)
        if (p[j] >= 1 && p[j] <= R && j <= C) assign[p[j] - 1] = j - 1;
}

void VelocityEstimator::estimate(const MapConst &mc, const FrameConst &fc, const float *planes0, const float *pts, int n,
                                 int model, std::vector<float> &out) {
    // rotated boundary planes and the in-view rotated cloud, same arithmetic as the device (dsp_dynamic.h:226-257)
    const int np = mc.Nh + mc.Nv + 2;
    std::vector<float> planes(3 * (size_t)np);
    for (int i = 0; i < np; ++i) dsp_rotate(planes0 + 3 * i, fc.q, fc.qi, &planes[3 * i]);
    const float *ph = planes.data(), *pv = ph + 3 * (mc.Nh + 1);
    auto dot = [](const float *nn, float x, float y, float z) { return x * nn[0] + y * nn[1] + z * nn[2]; };
    rotated.clear();
    for (int i = 0; i < n; ++i) {
        float r[3];
        dsp_rotate(pts + 3 * (size_t)i, fc.q, fc.qi, r);
        if (dot(ph, r[0], r[1], r[2]) >= 0.f && dot(ph + 3 * mc.Nh, r[0], r[1], r[2]) <= 0.f && dot(pv, r[0], r[1], r[2]) <= 0.f &&
            dot(pv + 3 * mc.Nv, r[0], r[1], r[2]) >= 0.f)
            rotated.insert(rotated.end(), r, r + 3);
    }
    const int nv = (int)rotated.size() / 3;
    if (nv == 0) return;  // :1379 — the previous cloud is kept
    out.clear();
    auto push = [&](float x, float y, float z, float vx, float vy, float vz, float inten) {
        const float rec[7] = {x, y, z, vx, vy, vz, inten};
        out.insert(out.end(), rec, rec + 7);
    };
    if (model == 1) {  // dsp_static.h:1285-1309
        for (int i = 0; i < nv; ++i) push(rotated[3 * i] + fc.cur[0], rotated[3 * i + 1] + fc.cur[1], rotated[3 * i + 2] + fc.cur[2], 0.f, 0.f, 0.f, 0.f);
        return;
    }
    std::vector<float> statics, nonground;  // xyz triples, world frame (:1387-1398)
    for (int i = 0; i < nv; ++i) {
        float x = rotated[3 * i] + fc.cur[0], y = rotated[3 * i + 1] + fc.cur[1], z = rotated[3 * i + 2] + fc.cur[2];
        std::vector<float> &dst = (z > filter_res) ? nonground : statics;
        dst.push_back(x); dst.push_back(y); dst.push_back(z);
    }
    std::vector<ClusterFeature> cur;
    if (!nonground.empty()) {
        std::vector<std::vector<int>> clusters;
        euclidean_clusters(nonground.data(), (int)nonground.size() / 3, 2 * filter_res, 5, 10000, clusters);  // :1410-1417
        std::vector<char> dynamic_flag;
        for (const auto &cl : clusters) {  // :1419-1447
            ClusterFeature f;
            f.intensity = uniform(0.1f, 1.f);
            for (int idx : cl) {
                f.cx += nonground[3 * idx]; f.cy += nonground[3 * idx + 1]; f.cz += nonground[3 * idx + 2];
                ++f.point_num;
            }
            f.cx /= (float)f.point_num; f.cy /= (float)f.point_num; f.cz /= (float)f.point_num;
            if (cl.size() > 200 || f.cz > 1.5) {  // DYNAMIC_CLUSTER_MAX_POINT_NUM / _MAX_CENTER_HEIGHT (:52-53)
                for (int idx : cl) { statics.push_back(nonground[3 * idx]); statics.push_back(nonground[3 * idx + 1]); statics.push_back(nonground[3 * idx + 2]); }
                dynamic_flag.push_back(0);
            } else {
                cur.push_back(f);
                dynamic_flag.push_back(1);
            }
        }
        const float distance_gate = 1.5f, maximum_velocity = 5.f;  // :1449-1451
        const int point_num_gate = 100;
        if (!last.empty() && !cur.empty() && fc.dt > 0.00001 && fc.dt < 10.0) {  // :1454-1455
            const int R = (int)cur.size(), C = (int)last.size();
            std::vector<float> cost((size_t)R * C), gate((size_t)R * C);
            for (int r = 0; r < R; ++r)
                for (int c = 0; c < C; ++c) {
                    float dx = cur[r].cx - last[c].cx, dy = cur[r].cy - last[c].cy, dz = cur[r].cz - last[c].cz;
                    float d = sqrtf(dx * dx + dy * dy + dz * dz);
                    if (std::abs(cur[r].point_num - last[c].point_num) > point_num_gate || d >= distance_gate) {
                        gate[(size_t)r * C + c] = 0.f;
                        cost[(size_t)r * C + c] = distance_gate * 5000.f;
                    } else {
                        gate[(size_t)r * C + c] = 1.f;
                        cost[(size_t)r * C + c] = d / distance_gate * 1000.f;
                    }
                }
            std::vector<int> assign;
            hungarian(cost, R, C, assign);
            for (int r = 0; r < R; ++r) {  // :1477-1499
                int c = assign[r];
                if (c < 0 || !(gate[(size_t)r * C + c] > 0.01f)) continue;
                ClusterFeature &f = cur[r];
                f.vx = (f.cx - last[c].cx) / fc.dt;
                f.vy = (f.cy - last[c].cy) / fc.dt;
                f.vz = (f.cz - last[c].cz) / fc.dt;
                f.v = sqrtf(f.vx * f.vx + f.vy * f.vy + f.vz * f.vz);
                f.intensity = last[c].intensity;
                if (f.v > maximum_velocity) { f.v = 0.f; f.vx = f.vy = f.vz = 0.f; }
            }
        }
        int seq = 0, dseq = 0;  // :1503-1524
        for (const auto &cl : clusters) {
            if (dynamic_flag[seq]) {
                const ClusterFeature &f = cur[dseq];
                for (int idx : cl) push(nonground[3 * idx], nonground[3 * idx + 1], nonground[3 * idx + 2], f.vx, f.vy, f.vz, f.intensity);
                ++dseq;
            }
            ++seq;
        }
    }
    for (size_t i = 0; i < statics.size() / 3; ++i) push(statics[3 * i], statics[3 * i + 1], statics[3 * i + 2], 0.f, 0.f, 0.f, 0.f);  // :1529-1540
    last = cur;  // :1542 (only reached when non-ground points exist in the reference too? no: assigned unconditionally)
}
