#!/usr/bin/env python3
"""Writes tests/golden/prefilter_96x72.npz: a small raw depth cloud (camera frame, NaN holes) and the cloud the
application's preprocessing (map_sim_example.cpp:305-336) hands to DSPMap::update for it, computed by the CPU restatement
oracle/prefilter_oracle.py (exact fixed-point centroids).  pcl::VoxelGrid itself cannot be built here (PCL is absent), so
this fixture pins the RESTATEMENT against regressions and pins the CUDA path to it bit for bit; it is not a PCL output.

    python tests/golden/make_prefilter_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "dsp-map_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from dspmap_b200.streams import make_depth_cloud  # noqa: E402
import prefilter_oracle as po  # noqa: E402

LEAF = 0.1                                              # ex:40
LO, HI = (-4.95, -4.95, -3.0), (4.95, 4.95, 3.0)        # ex:52-57 for the 66 x 66 x 40 map at 0.15 m
CAP = 5000                                              # ex:48


def main():
    raw = make_depth_cloud(96, 72, seed=7, stride=4)
    fin, idx, min_b, div_b = po.leaf_indices(raw, LEAF)
    leaves, counts, cen = po.centroids_exact(raw[:, :3][fin], idx)
    out = po.preprocess(raw, LEAF, LO, HI, CAP)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "prefilter_96x72.npz")
    np.savez_compressed(path, raw=raw, leaf=np.float32(LEAF), lo=np.float32(LO), hi=np.float32(HI), cap=np.int32(CAP),
                        min_b=min_b, div_b=div_b, leaves=leaves, counts=counts, centroids=cen, out=out)
    print(path, raw.shape, "->", out.shape, "leaves", len(leaves))


if __name__ == "__main__":
    main()
