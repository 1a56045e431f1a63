"""Application-side preprocessing (map_sim_example.cpp:305-336: pcl::VoxelGrid, axis swap, crop, cut) — SURVEY.md §8f row 3.
CPU tests pin the restatement (oracle/prefilter_oracle.py) to its committed fixture and to PCL's fp32 arithmetic within its
rounding error; GPU tests compare dspmap_prefilter_* with the restatement bit for bit through the C-ABI."""
import os

import numpy as np
import pytest

import dspmap_b200 as dm
import prefilter_oracle as po
from dspmap_b200.streams import make_depth_cloud
from parity import same

HERE = os.path.dirname(os.path.abspath(__file__))
LEAF = 0.1
LO, HI = (-4.95, -4.95, -3.0), (4.95, 4.95, 3.0)


def golden():
    return np.load(os.path.join(HERE, "golden", "prefilter_96x72.npz"))


def test_restatement_reproduces_its_fixture():
    G = golden()
    fin, idx, min_b, div_b = po.leaf_indices(G["raw"], float(G["leaf"]))
    leaves, counts, cen = po.centroids_exact(G["raw"][:, :3][fin], idx)
    assert np.array_equal(min_b, G["min_b"]) and np.array_equal(div_b, G["div_b"])
    assert np.array_equal(leaves, G["leaves"]) and np.array_equal(counts, G["counts"]) and same(cen, G["centroids"])
    assert same(po.preprocess(G["raw"], float(G["leaf"]), G["lo"], G["hi"], int(G["cap"])), G["out"])
    assert counts.sum() == fin.sum() and not fin.all()          # the fixture has NaN pixels and they are skipped


def test_exact_centroids_agree_with_pcl_fp32_arithmetic():
    raw = make_depth_cloud(160, 120, seed=3)
    a = po.preprocess(raw, LEAF, LO, HI, 5000, "exact")
    b = po.preprocess(raw, LEAF, LO, HI, 5000, "pcl")
    # same leaves survive (a centroid within 1e-6 of the crop bound could flip; none does on this cloud), values to fp32 rounding
    assert a.shape == b.shape and np.allclose(a, b, rtol=0, atol=2e-6)


def test_restatement_semantics():
    # leaf index formula of voxel_grid.hpp on a hand-checked cloud: leaf 0.5, two points per leaf, one NaN, one lone point
    raw = np.array([[0.1, 0.1, 1.0], [0.3, 0.2, 1.25], [np.nan, 0, 1], [-0.7, 0.1, 1.1], [-0.6, 0.4, 1.4], [2.2, -0.3, 3.3]], np.float32)
    fin, idx, min_b, div_b = po.leaf_indices(raw, 0.5)
    assert list(min_b) == [-2, -1, 2] and list(div_b) == [7, 2, 5]
    assert list(idx) == [2 + 7 + 0, 2 + 7 + 0, 0 + 7 + 0, 0 + 7 + 0, 6 + 0 + 4 * 14]
    out = po.preprocess(raw, 0.5, (-10, -10, -10), (10, 10, 10), 10)
    want = np.array([[1.25, 0.65, -0.25], [1.125, -0.2, -0.15], [3.3, -2.2, 0.3]], np.float32)   # ascending leaf: 7, 9, 62
    assert np.allclose(out, want, atol=1e-6)
    assert len(po.preprocess(raw, 0.5, (-10, -10, -10), (10, 10, 10), 2)) == 2                    # the cut keeps the first ones
    assert len(po.preprocess(raw, 0.5, (1.125, -10, -10), (10, 10, 10), 10)) == 2                 # open interval: 1.125 itself is out
    assert po.preprocess(raw[2:3], 0.5, LO, HI, 10).shape == (0, 3)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_prefilter_reproduces_fixture_and_restatement():
    G = golden()
    pf = dm.Prefilter(max_raw_points=640 * 480, max_stride=4, max_out_points=5000)
    out = pf.run(G["raw"], float(G["leaf"]), G["lo"], G["hi"])
    assert same(out, G["out"])
    assert pf.launches() == 7
    for (w, h, seed, stride) in [(640, 480, 1, 4), (640, 480, 2, 3), (320, 240, 5, 8), (33, 7, 9, 3)]:
        raw = make_depth_cloud(w, h, seed=seed, stride=stride)
        for leaf in (0.1, 0.15, 0.37):
            want = po.preprocess(raw, leaf, LO, HI, 5000)
            got = pf.run(raw, leaf, LO, HI)
            assert same(got, want), "%dx%d seed %d leaf %g: %d vs %d points" % (w, h, seed, leaf, len(got), len(want))
    pf.close()


@pytest.mark.gpu
def test_gpu_prefilter_edge_cases():
    pf = dm.Prefilter(max_raw_points=4096, max_stride=4, max_out_points=64, max_leaves=1 << 20)
    assert pf.run(np.zeros((0, 3), np.float32), LEAF, LO, HI).shape == (0, 3)                 # empty cloud
    assert pf.run(np.full((100, 3), np.nan, np.float32), LEAF, LO, HI).shape == (0, 3)        # nothing finite
    one = np.array([[0.5, -0.25, 2.0]], np.float32)
    assert pf.run(one, LEAF, LO, HI).tolist() == [[2.0, -0.5, 0.25]]                          # x = z, y = -x, z = -y
    raw = make_depth_cloud(64, 48, seed=4)
    want = po.preprocess(raw, LEAF, LO, HI, 10 ** 6)
    assert len(want) > 64
    assert same(pf.run(raw, LEAF, LO, HI), want[:64])                                         # the cut keeps the first 64 (ex:332-334)
    assert same(pf.run(raw, LEAF, LO, HI, cap=10), want[:10])
    assert same(pf.run(raw, LEAF, (0, 0, 0), (3, 3, 3)), po.preprocess(raw, LEAF, (0, 0, 0), (3, 3, 3), 64))
    far = np.array([[0, 0, 1], [500, 500, 500]], np.float32)                                   # bounding box of 1.25e11 leaves
    assert po.leaf_volume(far, LEAF) > (1 << 20) > po.leaf_volume(raw, LEAF)
    with pytest.raises(dm.DSPMapError):
        pf.run(far, LEAF, LO, HI)
    assert same(pf.run(raw, LEAF, LO, HI), want[:64])                                         # and the grid is still clean afterwards
    with pytest.raises(dm.DSPMapError):
        pf.run(np.zeros((5000, 4), np.float32), LEAF, LO, HI)                                  # larger than the staging capacity
    for _ in range(3):                                                                         # accumulators return to zero every frame
        assert same(pf.run(raw, LEAF, LO, HI), want[:64])
    pf.close()


@pytest.mark.gpu
def test_gpu_update_raw_equals_prefilter_then_update():
    from common import gpu_map, gpu_update
    cfg = dm.CONFIGS["cfg2"]
    a, b = gpu_map("cfg2", seed=5), gpu_map("cfg2", seed=5)
    pf = dm.Prefilter(max_raw_points=320 * 240, max_stride=4, max_out_points=5000)
    for f in range(4):
        raw = make_depth_cloud(320, 240, seed=20 + f, stride=4)
        pos, q, t = (0.05 * f, 0.0, 0.0), (1.0, 0.0, 0.0, 0.0), 0.1 * f
        cloud = po.preprocess(raw, LEAF, LO, HI, 5000)
        assert gpu_update(a, cloud, pos, t, q) == 1
        rc, nf = pf.update_raw(b, raw, LEAF, LO, HI, pos, t, q)
        assert rc == 1 and nf == len(cloud)
        ia, va = a.particles()
        ib, vb = b.particles()
        assert same(ia, ib) and same(va, vb), "frame %d" % f
    assert len(ia) > 1000
    for m in (a, b, pf):
        m.close()
