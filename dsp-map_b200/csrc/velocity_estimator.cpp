// velocity_estimator.cpp — see velocity_estimator.h.  Compiled with -ffp-contract=off: the fp32 expressions below are
// evaluated in the order the reference writes them.
#include "velocity_estimator.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

#include "dspmap_hostmath.h"

#ifdef EST_TIMING  // build.py -DEST_TIMING: where a frame's estimation time goes, printed at exit (diagnosis)
#include <chrono>
#include <cstdio>
#include <cstdlib>
namespace {
double est_us[10];
long est_frames;
void est_report() {
    fprintf(stderr, "estimator: front sweep %.1f  clustering %.1f  features + matching %.1f  tagged cloud %.1f  total %.1f us per frame (%ld frames)\n",
            est_us[0] / est_frames, est_us[1] / est_frames, est_us[2] / est_frames, est_us[3] / est_frames, est_us[4] / est_frames, est_frames);
    fprintf(stderr, "clustering: cells %.1f  links (2 passes) %.1f  components %.1f us per frame\n", est_us[5] / est_frames, est_us[6] / est_frames, est_us[7] / est_frames);
}
std::chrono::steady_clock::time_point est_last;
void est_mark(int k) {  // k < 0: start; else: charge the time since the last mark to slot k
    const auto now = std::chrono::steady_clock::now();
    if (k >= 0) est_us[k] += std::chrono::duration<double, std::micro>(now - est_last).count();
    est_last = now;
}
struct EstTimer {
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    int k;
    explicit EstTimer(int k_) : k(k_) {}
    ~EstTimer() { est_us[k] += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t).count(); }
};
}  // namespace
#define EST_TIME(k) EstTimer est_timer_##k(k)
#define EST_MARK(k) est_mark(k)
#else
#define EST_TIME(k)
#define EST_MARK(k)
#endif

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

namespace {
// Hash-grid path for clouds whose bounding box is too large for the dense grid (same components).
void hashed_clusters(const float *xyz, int n, float tol, int min_size, int max_size, std::vector<std::vector<int>> &out) {
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
}
}  // namespace

namespace {
// Scratch of the dense-grid clustering path, reused from frame to frame (no allocation in steady state).
struct ClusterScratch {
    std::vector<int> grid;  // padded dense cell grid -> compact cell number; all -1 between calls
    std::vector<float> box;
    std::vector<unsigned char> bits, rowocc;
    std::vector<int> cell_of, next, rep, tail, cell_lin, parent, croot, comp_of_root, comp_size;
};
const long long kDenseCells = 1ll << 23;  // 32 MB of int; larger bounding boxes take the hash-grid path

// Connected components on a dense cell grid over the cloud's bounding box.  Returns false when the box is too large.
bool dense_clusters(const float *xyz, int n, float tol, int min_size, int max_size, std::vector<std::vector<int>> &out,
                    const float *bbox = nullptr) {
    static thread_local ClusterScratch S;
    const float tol2 = tol * tol;
    const float inv_edge = 1.f / (tol * 0.57f);  // cells only need edge < tol / sqrt(3) up to rounding (1.3 % slack)
    auto cell = [inv_edge](float v) {
        const float f = v * inv_edge;
        int c = (int)f;
        return c - (int)(f < (float)c);  // floor
    };
    // bounding box first (floor is monotone, so the extreme cells are the cells of the extreme coordinates)
    float fl[3] = {xyz[0], xyz[1], xyz[2]}, fh[3] = {xyz[0], xyz[1], xyz[2]};
    if (bbox) {  // the caller already swept the points (VelocityEstimator::estimate)
        for (int k = 0; k < 3; ++k) { fl[k] = bbox[k]; fh[k] = bbox[3 + k]; }
    } else {
        for (int i = 1; i < n; ++i)
            for (int k = 0; k < 3; ++k) {
                const float v = xyz[3 * i + k];
                fl[k] = v < fl[k] ? v : fl[k];
                fh[k] = v > fh[k] ? v : fh[k];
            }
    }
    int lo[3];
    long long dim[3];
    for (int k = 0; k < 3; ++k) {
        if (!(fl[k] * inv_edge > -1e9f && fh[k] * inv_edge < 1e9f)) return false;  // also rejects NaN extremes
        lo[k] = cell(fl[k]);
        dim[k] = (long long)cell(fh[k]) - lo[k] + 5;  // 2 cells of padding on each side
    }
    const long long dx = dim[0], dy = dim[1], dz = dim[2];
    if (dx * dy > kDenseCells || dx * dy * dz > kDenseCells) return false;
    const size_t vol = (size_t)(dx * dy * dz);
    if (S.grid.size() < vol) S.grid.resize(vol, -1);
    if (S.bits.size() < vol / 8 + 16) S.bits.resize(vol / 8 + 16, 0);
    int *grid = S.grid.data();
    unsigned char *bits = S.bits.data();  // one occupancy bit per cell: the neighbour scan below stays in L1
    S.cell_of.resize(n); S.next.resize(n);
    S.rep.clear(); S.tail.clear(); S.cell_lin.clear(); S.box.clear();
    EST_MARK(-1);
    // cells numbered in order of first appearance; each cell's point list ascends by index, its head is rep[c]
    const int ox = 2 - lo[0], oy = 2 - lo[1], oz = 2 - lo[2];
    for (int i = 0; i < n; ++i) {
        const int cx = cell(xyz[3 * i]) + ox, cy = cell(xyz[3 * i + 1]) + oy, cz = cell(xyz[3 * i + 2]) + oz;
        if ((unsigned)cx >= (unsigned)dx || (unsigned)cy >= (unsigned)dy || (unsigned)cz >= (unsigned)dz) {  // a NaN coordinate
            for (size_t c = 0; c < S.cell_lin.size(); ++c) { grid[S.cell_lin[c]] = -1; bits[S.cell_lin[c] >> 3] = 0; }
            return false;
        }
        const int l = (int)(((long long)cz * dy + cy) * dx + cx);
        int c = grid[l];
        const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        if (c < 0) {
            c = (int)S.rep.size();
            grid[l] = c;
            bits[l >> 3] |= (unsigned char)(1u << (l & 7));
            S.rep.push_back(i); S.tail.push_back(i); S.cell_lin.push_back(l);
            const float b6[6] = {x, y, z, x, y, z};
            S.box.insert(S.box.end(), b6, b6 + 6);
        } else {
            S.next[S.tail[c]] = i;
            S.tail[c] = i;
            float *bb = &S.box[6 * (size_t)c];  // the cell's points' bounding box: a cheap lower bound for linked()
            bb[0] = std::min(bb[0], x); bb[1] = std::min(bb[1], y); bb[2] = std::min(bb[2], z);
            bb[3] = std::max(bb[3], x); bb[4] = std::max(bb[4], y); bb[5] = std::max(bb[5], z);
        }
        S.next[i] = -1;
        S.cell_of[i] = c;
    }
    EST_MARK(5);
    const int nc = (int)S.rep.size();
    const float *box = S.box.data();
    auto apart = [&](int ca, int cb) {  // no pair of the two cells can be within tol
        const float *A = box + 6 * (size_t)ca, *B = box + 6 * (size_t)cb;
        float d2 = 0.f;
        for (int k = 0; k < 3; ++k) {
            const float g = std::max(std::max(B[k] - A[k + 3], A[k] - B[k + 3]), 0.f);
            d2 += g * g;
        }
        return d2 > tol2 * 1.0001f;  // margin: the bound is evaluated in a different rounding than the pair test
    };
    S.parent.resize(nc);
    int *parent = S.parent.data();
    for (int c = 0; c < nc; ++c) parent[c] = c;
    auto root = [&](int c) {
        while (parent[c] != c) { parent[c] = parent[parent[c]]; c = parent[c]; }
        return c;
    };
    const int *next = S.next.data();
    auto linked = [&](int ha, int hb) {  // is there a pair (i in cell a, j in cell b) within tol?
        for (int i = ha; i >= 0; i = next[i]) {
            const float px = xyz[3 * i], py = xyz[3 * i + 1], pz = xyz[3 * i + 2];
            for (int j = hb; j >= 0; j = next[j]) {
                const float ddx = xyz[3 * j] - px, ddy = xyz[3 * j + 1] - py, ddz = xyz[3 * j + 2] - pz;
                if (ddx * ddx + ddy * ddy + ddz * ddz <= tol2) return true;
            }
        }
        return false;
    };
    // Each unordered pair of cells once: the lexicographically positive half of the 5 x 5 x 5 neighbourhood, scanned as
    // rows of five x-adjacent cells (one cache line; an all-empty row is skipped with one test).  Adjacent cells (pass 0)
    // before cells two apart (pass 1): most of the far links are then already implied.
    struct Row { int off; int amask[2]; };  // bit (a + 2) of amask[pass] set: offset a of this row belongs to the pass
    Row rows[13];
    int nrows = 0;
    for (int d = 0; d <= 2; ++d)
        for (int b = -2; b <= 2; ++b) {
            if (d == 0 && b < 0) continue;
            int m[2] = {0, 0};
            for (int a = -2; a <= 2; ++a) {
                if (d == 0 && b == 0 && a <= 0) continue;
                const int ch = std::max(std::max(a < 0 ? -a : a, b < 0 ? -b : b), d);
                m[ch - 1] |= 1 << (a + 2);
            }
            if (m[0] | m[1]) rows[nrows++] = Row{(int)((d * dy + b) * dx), {m[0], m[1]}};
        }
    // the occupancy of every row is read from the bitmap once (pass 0) and kept for pass 1: 13 bytes per cell
    S.rowocc.resize((size_t)nc * 13);
    unsigned char *rowocc = S.rowocc.data();
    for (int pass = 0; pass < 2; ++pass)
        for (int c = 0; c < nc; ++c) {
            const int base = S.cell_lin[c];
            const int hc = S.rep[c];
            for (int k = 0; k < nrows; ++k) {
                const int pos = base + rows[k].off - 2;  // >= 0: the grid is padded by two cells
                unsigned row5;
                if (pass == 0) {
                    uint64_t w;
                    memcpy(&w, bits + (pos >> 3), 8);
                    row5 = (unsigned)(w >> (pos & 7)) & 31u;
                    rowocc[(size_t)c * 13 + k] = (unsigned char)row5;
                } else {
                    row5 = rowocc[(size_t)c * 13 + k];
                }
                unsigned occ = row5 & (unsigned)rows[k].amask[pass];
                while (occ) {
                    const int a = __builtin_ctz(occ);
                    occ &= occ - 1;
                    const int cb = grid[pos + a];
                    const int ra = root(c), rb = root(cb);
                    if (ra == rb) continue;
                    if (!apart(c, cb) && linked(hc, S.rep[cb])) parent[std::max(ra, rb)] = std::min(ra, rb);
                }
            }
        }
    EST_MARK(6);
    for (int c = 0; c < nc; ++c) { grid[S.cell_lin[c]] = -1; bits[S.cell_lin[c] >> 3] = 0; }  // leave the grid clean for the next call
    // gather components; a component is discovered at its smallest point index, like a seeded flood fill would
    S.croot.resize(nc);
    for (int c = 0; c < nc; ++c) S.croot[c] = root(c);
    S.comp_of_root.assign(nc, -1);
    S.comp_size.clear();
    for (int i = 0; i < n; ++i) {
        const int r = S.croot[S.cell_of[i]];
        if (S.comp_of_root[r] < 0) { S.comp_of_root[r] = (int)S.comp_size.size(); S.comp_size.push_back(0); }
        ++S.comp_size[S.comp_of_root[r]];
    }
    const int ncomp = (int)S.comp_size.size();
    std::vector<int> slot(ncomp, -1);
    for (int k = 0; k < ncomp; ++k)
        if (S.comp_size[k] >= min_size && S.comp_size[k] <= max_size) {
            slot[k] = (int)out.size();
            out.emplace_back();
            out.back().reserve(S.comp_size[k]);
        }
    for (int i = 0; i < n; ++i) {
        const int s = slot[S.comp_of_root[S.croot[S.cell_of[i]]]];
        if (s >= 0) out[s].push_back(i);  // ascending indices by construction
    }
    EST_MARK(7);
    return true;
}
}  // namespace

void euclidean_clusters(const float *xyz, int n, float tol, int min_size, int max_size, std::vector<std::vector<int>> &out) {
    // Connected components of the graph "d2 <= tol^2" through a union-find over grid cells.  The cell edge is below
    // tol / sqrt(3), so all points of one cell are mutually linked; links between cells up to two cells apart need one
    // witness pair.  Cost is linear in the points for the dense clouds a nearby surface produces (a per-point
    // neighbourhood search is quadratic there).  Components do not depend on the traversal, so the result is identical.
    euclidean_clusters_path(xyz, n, tol, min_size, max_size, 0, out);
}

namespace {
bool clusters_impl(const float *xyz, int n, float tol, int min_size, int max_size, int path, std::vector<std::vector<int>> &out, const float *bbox);
}
bool euclidean_clusters_path(const float *xyz, int n, float tol, int min_size, int max_size, int path, std::vector<std::vector<int>> &out) {
    return clusters_impl(xyz, n, tol, min_size, max_size, path, out, nullptr);
}
namespace {
bool clusters_impl(const float *xyz, int n, float tol, int min_size, int max_size, int path, std::vector<std::vector<int>> &out, const float *bbox) {
    out.clear();
    if (n == 0 || !(tol > 0.f)) return true;
    bool done = false;
    if (path != 1) {
        done = dense_clusters(xyz, n, tol, min_size, max_size, out, bbox);
        if (!done) out.clear();
        if (!done && path == 2) return false;
    }
    if (!done) hashed_clusters(xyz, n, tol, min_size, max_size, out);
    std::stable_sort(out.begin(), out.end(), [](const std::vector<int> &a, const std::vector<int> &b) { return a.size() > b.size(); });
    return true;
}
}  // namespace

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

namespace {
typedef float vf8 __attribute__((vector_size(32)));
typedef int vi8 __attribute__((vector_size(32)));

// Rotates n sensor-frame points into the world-aligned frame and keeps those inside the four outer FOV planes
// (dsp_dynamic.h:226-257), eight points per step.  Every lane evaluates dsp_rotate's expressions in dsp_rotate's order
// (no contraction: the file is compiled with -ffp-contract=off and neither target has FMA enabled), so the kept points are
// bit-identical to the scalar helper's.  nrm = the four rotated plane normals (h first, h last, v first, v last).
__attribute__((target_clones("avx2", "default")))
int rotate_in_view(const float *pts, int n, const float *q, const float *qi, const float *nrm, float *kept) {
    const float aw = q[0], ax = q[1], ay = q[2], az = q[3];
    const float iw = qi[0], ix = qi[1], iy = qi[2], iz = qi[3];
    int nk = 0;
    for (int i0 = 0; i0 < n; i0 += 8) {
        alignas(32) float sx[8], sy[8], sz[8];
        const int m = std::min(8, n - i0);
        for (int l = 0; l < m; ++l) {
            sx[l] = pts[3 * (size_t)(i0 + l)];
            sy[l] = pts[3 * (size_t)(i0 + l) + 1];
            sz[l] = pts[3 * (size_t)(i0 + l) + 2];
        }
        for (int l = m; l < 8; ++l) sx[l] = sy[l] = sz[l] = 0.f;
        const vf8 bx = *(const vf8 *)sx, by = *(const vf8 *)sy, bz = *(const vf8 *)sz;
        const vf8 bw = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // scalar component of the pure quaternion (0, v)
        const vf8 tw = aw * bw - ax * bx - ay * by - az * bz;
        const vf8 tx = aw * bx + ax * bw + ay * bz - az * by;
        const vf8 ty = aw * by + ay * bw + az * bx - ax * bz;
        const vf8 tz = aw * bz + az * bw + ax * by - ay * bx;
        const vf8 rx = tw * ix + tx * iw + ty * iz - tz * iy;
        const vf8 ry = tw * iy + ty * iw + tz * ix - tx * iz;
        const vf8 rz = tw * iz + tz * iw + tx * iy - ty * ix;
        const vf8 d0 = rx * nrm[0] + ry * nrm[1] + rz * nrm[2];
        const vf8 d1 = rx * nrm[3] + ry * nrm[4] + rz * nrm[5];
        const vf8 d2 = rx * nrm[6] + ry * nrm[7] + rz * nrm[8];
        const vf8 d3 = rx * nrm[9] + ry * nrm[10] + rz * nrm[11];
        const vi8 in = (d0 >= 0.f) & (d1 <= 0.f) & (d2 <= 0.f) & (d3 >= 0.f);
        for (int l = 0; l < m; ++l)
            if (in[l]) {
                kept[3 * (size_t)nk] = rx[l];
                kept[3 * (size_t)nk + 1] = ry[l];
                kept[3 * (size_t)nk + 2] = rz[l];
                ++nk;
            }
    }
    return nk;
}
}  // namespace

namespace {
// The dynamic model's front end in ONE sweep over the cloud: rotation + FOV test as in rotate_in_view, the shift into the
// world frame, the ground / non-ground split (dsp_dynamic.h:1387-1398) and the non-ground bounding box the clustering grid
// needs.  Both destinations are written and only one cursor advances (the side a point falls on is data-dependent).  Every
// lane evaluates the scalar expressions in the scalar order, so the outputs are bit-identical to the three separate passes.
// Returns the number of points in view; counts[0] = floats written to sp, counts[1] = floats written to gp.
__attribute__((target_clones("avx2", "default")))
int rotate_split(const float *pts, int n, const float *q, const float *qi, const float *nrm, const float *cur, float filter_res,
                 float *sp, float *gp, size_t *counts, float *bbox) {
    const float aw = q[0], ax = q[1], ay = q[2], az = q[3];
    const float iw = qi[0], ix = qi[1], iy = qi[2], iz = qi[3];
    const float cx = cur[0], cy = cur[1], cz = cur[2];
    float *sp0 = sp, *gp0 = gp;
    const vf8 big = {3.4e38f, 3.4e38f, 3.4e38f, 3.4e38f, 3.4e38f, 3.4e38f, 3.4e38f, 3.4e38f};
    vf8 lo0 = big, lo1 = big, lo2 = big, hi0 = -big, hi1 = -big, hi2 = -big;  // per-lane running box of the non-ground points
    int nk = 0;
    for (int i0 = 0; i0 < n; i0 += 8) {
        alignas(32) float sx[8], sy[8], sz[8];
        const int m = std::min(8, n - i0);
        for (int l = 0; l < m; ++l) {
            sx[l] = pts[3 * (size_t)(i0 + l)];
            sy[l] = pts[3 * (size_t)(i0 + l) + 1];
            sz[l] = pts[3 * (size_t)(i0 + l) + 2];
        }
        for (int l = m; l < 8; ++l) sx[l] = sy[l] = sz[l] = 0.f;
        const vf8 bx = *(const vf8 *)sx, by = *(const vf8 *)sy, bz = *(const vf8 *)sz;
        const vf8 bw = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const vf8 tw = aw * bw - ax * bx - ay * by - az * bz;
        const vf8 tx = aw * bx + ax * bw + ay * bz - az * by;
        const vf8 ty = aw * by + ay * bw + az * bx - ax * bz;
        const vf8 tz = aw * bz + az * bw + ax * by - ay * bx;
        const vf8 rx = tw * ix + tx * iw + ty * iz - tz * iy;
        const vf8 ry = tw * iy + ty * iw + tz * ix - tx * iz;
        const vf8 rz = tw * iz + tz * iw + tx * iy - ty * ix;
        const vf8 d0 = rx * nrm[0] + ry * nrm[1] + rz * nrm[2];
        const vf8 d1 = rx * nrm[3] + ry * nrm[4] + rz * nrm[5];
        const vf8 d2 = rx * nrm[6] + ry * nrm[7] + rz * nrm[8];
        const vf8 d3 = rx * nrm[9] + ry * nrm[10] + rz * nrm[11];
        const vi8 in = (d0 >= 0.f) & (d1 <= 0.f) & (d2 <= 0.f) & (d3 >= 0.f);
        const vf8 wx = rx + cx, wy = ry + cy, wz = rz + cz;
        const vi8 ngm = in & (wz > filter_res);  // in view and above the ground threshold (padding lanes are never in view)
        const vf8 lx = ngm ? wx : big, ly = ngm ? wy : big, lz = ngm ? wz : big;
        const vf8 hx = ngm ? wx : -big, hy = ngm ? wy : -big, hz = ngm ? wz : -big;
        lo0 = lx < lo0 ? lx : lo0; lo1 = ly < lo1 ? ly : lo1; lo2 = lz < lo2 ? lz : lo2;
        hi0 = hx > hi0 ? hx : hi0; hi1 = hy > hi1 ? hy : hi1; hi2 = hz > hi2 ? hz : hi2;
        for (int l = 0; l < m; ++l)
            if (in[l]) {
                const float x = wx[l], y = wy[l], z = wz[l];
                sp[0] = x; sp[1] = y; sp[2] = z;
                gp[0] = x; gp[1] = y; gp[2] = z;
                const int ng = z > filter_res;
                gp += 3 * ng;
                sp += 3 * (1 - ng);
                ++nk;
            }
    }
    counts[0] = (size_t)(sp - sp0);
    counts[1] = (size_t)(gp - gp0);
    const vf8 *lo[3] = {&lo0, &lo1, &lo2}, *hi[3] = {&hi0, &hi1, &hi2};
    for (int k = 0; k < 3; ++k) {
        float a = (*lo[k])[0], b = (*hi[k])[0];
        for (int l = 1; l < 8; ++l) { a = (*lo[k])[l] < a ? (*lo[k])[l] : a; b = (*hi[k])[l] > b ? (*hi[k])[l] : b; }
        bbox[k] = a;
        bbox[3 + k] = b;
    }
    return nk;
}
}  // namespace

// Hungarian matching of this frame's dynamic clusters to the previous frame's and the velocities that follow (:1449-1499)
void VelocityEstimator::match(std::vector<ClusterFeature> &cur, float dt) {
    const float distance_gate = 1.5f, maximum_velocity = 5.f;  // :1449-1451
    const int point_num_gate = 100;
    if (!last.empty() && !cur.empty() && dt > 0.00001 && dt < 10.0) {  // :1454-1455
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
            f.vx = (f.cx - last[c].cx) / dt;
            f.vy = (f.cy - last[c].cy) / dt;
            f.vz = (f.cz - last[c].cz) / dt;
            f.v = sqrtf(f.vx * f.vx + f.vy * f.vy + f.vz * f.vz);
            f.intensity = last[c].intensity;
            if (f.v > maximum_velocity) { f.v = 0.f; f.vx = f.vy = f.vz = 0.f; }
        }
    }
}

// Second half of a frame whose front end ran on the device (dspmap_estimator.cuh): feat = the dynamic clusters in output
// order, n_clusters = all clusters of admissible size (each drew a colour, :1422).  cvel: 4 floats per dynamic cluster.
void VelocityEstimator::finish_device(const EstFeature *feat, int n_dynamic, int n_clusters, float dt, float *cvel) {
    std::vector<ClusterFeature> cur((size_t)n_dynamic);
    for (int r = 0; r < n_dynamic; ++r) {
        ClusterFeature &f = cur[r];
        f.cx = feat[r].cx; f.cy = feat[r].cy; f.cz = feat[r].cz;
        f.point_num = feat[r].size;
        f.intensity = dsp_uniform(seed, draws + (u64)feat[r].sorted_pos, 0.1f, 1.f);
    }
    draws += (u64)n_clusters;
    if (n_dynamic > 0) match(cur, dt);
    for (int r = 0; r < n_dynamic; ++r) {
        cvel[4 * r] = cur[r].vx; cvel[4 * r + 1] = cur[r].vy; cvel[4 * r + 2] = cur[r].vz; cvel[4 * r + 3] = cur[r].intensity;
    }
    last = cur;
}

void VelocityEstimator::estimate(const MapConst &mc, const FrameConst &fc, const float *planes0, const float *pts, int n,
                                 int model, std::vector<float> &out) {
    // rotated outer boundary planes and the in-view rotated cloud, same arithmetic as the device (dsp_dynamic.h:226-257)
#ifdef EST_TIMING
    if (est_frames++ == 0) atexit(est_report);
#endif
    EST_TIME(4);
    float nrm[12];
    dsp_rotate(planes0, fc.q, fc.qi, nrm);
    dsp_rotate(planes0 + 3 * mc.Nh, fc.q, fc.qi, nrm + 3);
    dsp_rotate(planes0 + 3 * (mc.Nh + 1), fc.q, fc.qi, nrm + 6);
    dsp_rotate(planes0 + 3 * (mc.Nh + 1 + mc.Nv), fc.q, fc.qi, nrm + 9);
    int nv;
    size_t split_counts[2] = {0, 0};
    float ng_box[6];
    if (model == 1) {
        rotated.resize(3 * (size_t)n);
        nv = rotate_in_view(pts, n, fc.q, fc.qi, nrm, rotated.data());
    } else {  // rotation, FOV test, world shift, ground split and the clustering grid's bounding box in one sweep
        statics.resize(3 * (size_t)n + 3);
        nonground.resize(3 * (size_t)n + 3);
        EST_TIME(0);
        nv = rotate_split(pts, n, fc.q, fc.qi, nrm, fc.cur, filter_res, statics.data(), nonground.data(), split_counts, ng_box);
    }
    if (nv == 0) return;  // :1379 — the previous cloud is kept
    out.resize(7 * (size_t)nv);  // every in-view point is written at most once
    size_t no = 0;
    float *o = out.data();
    auto push = [&](float x, float y, float z, float vx, float vy, float vz, float inten) {
        o[no] = x; o[no + 1] = y; o[no + 2] = z; o[no + 3] = vx; o[no + 4] = vy; o[no + 5] = vz; o[no + 6] = inten;
        no += 7;
    };
    if (model == 1) {  // dsp_static.h:1285-1309
        for (int i = 0; i < nv; ++i) push(rotated[3 * i] + fc.cur[0], rotated[3 * i + 1] + fc.cur[1], rotated[3 * i + 2] + fc.cur[2], 0.f, 0.f, 0.f, 0.f);
        return;
    }
    // ground / non-ground split, xyz triples in the world frame (:1387-1398).  Both destinations are written and only
    // one cursor advances: the side a point falls on is data-dependent, a branch here mispredicts
    statics.resize(split_counts[0]);
    nonground.resize(split_counts[1]);
    statics.reserve(3 * (size_t)nv);  // clusters reclassified as static are appended below
    std::vector<ClusterFeature> cur;
    if (!nonground.empty()) {
        std::vector<std::vector<int>> clusters;
        {
            EST_TIME(1);
            clusters_impl(nonground.data(), (int)nonground.size() / 3, 2 * filter_res, 5, 10000, 0, clusters, ng_box);  // :1410-1417
        }
        EST_TIME(2);
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
        match(cur, fc.dt);
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
    out.resize(no);
    last = cur;  // :1542 (only reached when non-ground points exist in the reference too? no: assigned unconditionally)
}
