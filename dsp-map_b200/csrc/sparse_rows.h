// sparse_rows.h — host side of the sparse copy-out of the future-status grid (product code, DSPMAP_SPARSE_FUTURE).
// The caller's V x T array holds exactly what the previous call wrote; a call brings the non-zero voxel rows of the new
// grid.  Rows that were non-zero last time are cleared, the new rows are written, and the row list is remembered: the
// array then equals the dense grid without 4 * V * T bytes having crossed PCIe or been written by the host.
#pragma once
#include <cstring>
#include <vector>

struct SparseRows {
    float *owner = nullptr;  // the array that holds exactly the previous call's result (nullptr: unknown content)
    std::vector<int> prev;   // its non-zero rows

    void invalidate() {
        owner = nullptr;
        prev.clear();
    }
    // First half, independent of the new data: the rows that were non-zero last time are cleared.  The reader calls it while it
    // waits for the device (the patching is bound by cache misses on rows scattered over a few MB, ~5 ns each: 10 k rows cleared
    // and 10 k written took 0.1 ms behind the synchronisation; the clearing now hides in the wait and leaves the lines warm).
    void clear_previous(float *future, int V, int T) {
        const size_t row = sizeof(float) * (size_t)T;
        if (owner != future) {  // first use of this array (or something else wrote it in between): clear all of it once
            memset(future, 0, row * (size_t)V);
            owner = future;
        } else {
            for (int v : prev) memset(future + (size_t)v * T, 0, row);
        }
        prev.clear();
    }
    // future: V x T floats; idx[k] / val[k * T .. k * T + T): the nf non-zero rows of the new grid
    void apply(float *future, int V, int T, const int *idx, const float *val, int nf) {
        const size_t row = sizeof(float) * (size_t)T;
        clear_previous(future, V, T);  // (nothing left to do when the reader has called it already)
        for (int k = 0; k < nf; ++k) memcpy(future + (size_t)idx[k] * T, val + (size_t)k * T, row);
        prev.assign(idx, idx + nf);
    }
};
