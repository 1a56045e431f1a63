"""CPU restatement of the application's point-cloud preprocessing (TEST INFRASTRUCTURE — only tests/, smoke() and bench.py's
baseline legs may import this; the product is dsp-map_b200/csrc/prefilter.cu).

Follows g-ch/DSP-map src/map_sim_example.cpp:305-336: pcl::VoxelGrid at leaf size `res` (ex:312-316), axis swap x = z,
y = -x, z = -y (ex:320-322), open-interval crop (ex:325, inRange ex:190-197), cut at MAX_POINT_NUM (ex:332-334).

pcl::VoxelGrid is a third-party dependency absent from /root/reference (PCL 1.8 / 1.10, the versions ROS Melodic / Noetic
bundle, readme.md:23-25; unpinned).  Its published algorithm (filters/include/pcl/filters/impl/voxel_grid.hpp, applyFilter)
is restated in `leaf_indices`: skip non-finite points; min_b = floor(min_p * inv_leaf); leaf index
ijk . (1, div_b.x, div_b.x div_b.y) with ijk = (int)(floor(p * inv_leaf) - (float)min_b); one centroid per occupied leaf in
ascending leaf index.  **Parity unpinned** for this row: the reference tree has no golden vectors for it and PCL cannot be
built here.  Two centroid arithmetics are given:
  * `centroids_exact`  — exact sum in 2^-24 m fixed point, one fp64 division, rounded to fp32: what the CUDA path computes,
                         compared bit for bit;
  * `centroids_pcl_fp32` — PCL's arithmetic (fp32 running sum of the leaf's points, then a division by the fp32 count) in
                         input order, i.e. what PCL would give with a stable sort; PCL's actual order after its unstable
                         std::sort is implementation-defined, so this is compared with a tolerance only.
"""
import numpy as np

FIX = 16777216.0  # 2^24


def leaf_indices(pts, leaf):
    """pts (n, >=3) float32 -> (finite mask, leaf index per finite point (int64), min_b, div_b); None when nothing is finite."""
    p = np.asarray(pts, np.float32)[:, :3]
    fin = np.isfinite(p).all(axis=1)
    q = p[fin]
    if len(q) == 0:
        return fin, None, None, None
    inv = np.float32(1.0) / np.float32(leaf)
    min_b = np.floor(q.min(axis=0) * inv).astype(np.int32)
    max_b = np.floor(q.max(axis=0) * inv).astype(np.int32)
    div_b = (max_b - min_b + 1).astype(np.int64)
    ijk = (np.floor(q * inv) - min_b.astype(np.float32)).astype(np.int32).astype(np.int64)
    idx = ijk[:, 0] + ijk[:, 1] * div_b[0] + ijk[:, 2] * div_b[0] * div_b[1]
    return fin, idx, min_b, div_b


def centroids_exact(q, idx):
    """(ascending occupied leaf ids, counts, fp32 centroids) with the exact fixed-point sum."""
    leaves, inverse, counts = np.unique(idx, return_inverse=True, return_counts=True)
    fixed = np.rint(q.astype(np.float64) * FIX).astype(np.int64)
    sums = np.zeros((len(leaves), 3), np.int64)
    np.add.at(sums, inverse, fixed)
    cen = (sums.astype(np.float64) / (FIX * counts.astype(np.float64))[:, None]).astype(np.float32)
    return leaves, counts, cen


def centroids_pcl_fp32(q, idx):
    """PCL's fp32 running sum in input order (stable sort), centroid /= (float)n."""
    order = np.argsort(idx, kind="stable")
    leaves, start, counts = np.unique(idx[order], return_index=True, return_counts=True)
    cen = np.zeros((len(leaves), 3), np.float32)
    for k, (s, c) in enumerate(zip(start, counts)):
        acc = np.zeros(3, np.float32)
        for i in order[s:s + c]:
            acc = acc + q[i]
        cen[k] = acc / np.float32(c)
    return leaves, counts, cen


def preprocess(pts, leaf, range_min, range_max, cap, arithmetic="exact"):
    """The whole of ex:305-336: returns the (m, 3) float32 cloud handed to DSPMap::update (m <= cap)."""
    fin, idx, _, _ = leaf_indices(pts, leaf)
    if idx is None:
        return np.zeros((0, 3), np.float32)
    q = np.asarray(pts, np.float32)[:, :3][fin]
    _, _, cen = (centroids_exact if arithmetic == "exact" else centroids_pcl_fp32)(q, idx)
    out = np.stack([cen[:, 2], -cen[:, 0], -cen[:, 1]], axis=1).astype(np.float32)
    lo, hi = np.asarray(range_min, np.float32), np.asarray(range_max, np.float32)
    keep = ((out > lo) & (out < hi)).all(axis=1)
    return np.ascontiguousarray(out[keep][:cap])


def leaf_volume(pts, leaf):
    """Leaves in the bounding box (what the CUDA path checks against its capacity)."""
    _, idx, _, div_b = leaf_indices(pts, leaf)
    return 0 if idx is None else int(div_b[0] * div_b[1] * div_b[2])
