// Stand-in for saebyn/munkres-cpp "munkres.h" — TEST INFRASTRUCTURE ONLY.
// Used by the reference's side thread only (dsp_dynamic.h:1456-1481); NOT on the hot path.
// Contract kept: Matrix<T>(rows, cols) with operator()(r, c); Munkres<T>::solve(m) computes a
// minimum-cost assignment of the matrix padded to square with its maximum element and leaves
// 0 at assigned cells, -1 elsewhere.
#pragma once
#include <algorithm>
#include <limits>
#include <vector>
template <typename T>
class Matrix {
public:
    Matrix(size_t r, size_t c) : r_(r), c_(c), d_(r * c, T(0)) {}
    T &operator()(size_t r, size_t c) { return d_[r * c_ + c]; }
    const T &operator()(size_t r, size_t c) const { return d_[r * c_ + c]; }
    size_t rows() const { return r_; }
    size_t columns() const { return c_; }
private:
    size_t r_, c_;
    std::vector<T> d_;
};
template <typename T>
class Munkres {
public:
    void solve(Matrix<T> &m) {
        const int R = (int)m.rows(), C = (int)m.columns(), n = std::max(R, C);
        if (n == 0) return;
        double mx = 0;
        for (int r = 0; r < R; ++r)
            for (int c = 0; c < C; ++c) mx = std::max(mx, (double)m(r, c));
        std::vector<double> a((size_t)(n + 1) * (n + 1), mx);
        for (int r = 0; r < R; ++r)
            for (int c = 0; c < C; ++c) a[(size_t)(r + 1) * (n + 1) + (c + 1)] = (double)m(r, c);
        // shortest-augmenting-path Hungarian, 1-based potentials
        const double INF = std::numeric_limits<double>::infinity();
        std::vector<double> u(n + 1, 0), v(n + 1, 0), minv(n + 1);
        std::vector<int> p(n + 1, 0), way(n + 1, 0);
        std::vector<char> used(n + 1);
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
                        double cur = a[(size_t)i0 * (n + 1) + j] - u[i0] - v[j];
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
        for (int r = 0; r < R; ++r)
            for (int c = 0; c < C; ++c) m(r, c) = T(-1);
        for (int j = 1; j <= n; ++j)
            if (p[j] >= 1 && p[j] <= R && j <= C) m(p[j] - 1, j - 1) = T(0);
    }
};
