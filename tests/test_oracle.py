"""CPU tests of the oracle: the restatement (oracle/dsp_oracle.cpp) against the golden vectors produced by the
unmodified reference, against the reference itself where oracle/_ref is built, and known-answer cases derived from
the reference code (SURVEY.md §4 tier 2)."""
import numpy as np
import pytest

import refmap
from common import GOLDEN, check_against_golden, load_golden, make_stream
from dspmap_b200.configs import CONFIGS, derive
from oracle import OracleMap
from parity import same


@pytest.mark.parametrize("name,frames,initp", GOLDEN)
def test_restatement_reproduces_reference_golden(name, frames, initp):
    G = load_golden(name, frames, initp)
    o = OracleMap(CONFIGS[name], seed=int(G["seed"]), init_particles=initp)
    assert same(o.pdf_table(), G["pdf"]) and same(o.neighbors(), G["neighbors"])
    for f in range(frames):
        rc = o.update(G["points"][f], G["pos"][f], G["t"][f], G["quat"][f], tagged=G["tagged_%d" % f])
        assert rc == int(G["rc_%d" % f])
        check_against_golden(o, G, f, is_gpu=False)
        if f % 2 == 1:
            xyz, fut = o.occupancy(0.2)
            assert same(xyz, G["occ_xyz_%d" % f])
            ref = np.zeros(fut.size, np.float32)
            ref[G["fut_idx_%d" % f]] = G["fut_val_%d" % f]
            assert same(fut.ravel(), ref)


LIVE = [("tiny_dyn", 8, 0), ("tiny_mn", 6, 0), ("tiny_static", 6, 0), ("tiny_mn", 4, 2000), ("cfg1", 4, 0)]


def test_restatement_equals_reference_on_the_bench_stream_including_the_plane_normals():
    """cfg2 on bench.py's stream (seed 1), 8 frames.  Frame 6 holds a particle within 1e-7 of a pyramid boundary plane: it
    lands in the reference's pyramid only if the plane normals are the reference's to the last bit, i.e. if the angular
    resolution in radians is the fp32 product dsp_dynamic.h:543 forms with glibc's float M_PIf32 (dspmap_config.pi_is_double)."""
    name = "cfg2"
    if not refmap.available(name):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    cfg = CONFIGS[name]
    st = make_stream(cfg, seed=1, frames=8)
    r, o = refmap.RefMap(name, seed=7), OracleMap(cfg, seed=7)
    for f in range(8):
        r.update(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f])
        o.update(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f], tagged=r.tagged_cloud())
        for x, y in zip(r.plane_normals(), o.plane_normals()):
            assert same(x, y), "frame %d boundary-plane normals" % f
        for x, y in zip(r.pyramid_lists(), o.pyramid_lists()):
            assert same(x, y), "frame %d pyramid lists" % f
        for x, y in zip(r.particles(), o.particles()):
            assert same(x, y), "frame %d particles" % f


@pytest.mark.parametrize("name,frames,initp", LIVE)
def test_restatement_equals_reference_live(name, frames, initp):
    if not refmap.available(name):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    cfg = CONFIGS[name]
    st = make_stream(cfg, seed=3, frames=frames)
    r = refmap.RefMap(name, seed=7, init_particles=initp)
    o = OracleMap(cfg, seed=7, init_particles=initp)
    d = derive(cfg)
    assert (r.V, r.S, r.P, r.L, r.T) == (o.V, o.S, o.P, o.L, o.T) == (d["V"], d["S"], d["P"], d["L"], d["T"])
    a, b = r.gaussian_tables(100000), o.gaussian_tables(100000)
    assert same(a[0], b[0]) and same(a[1], b[1]) and same(r.pdf_table(), o.pdf_table()) and same(r.neighbors(), o.neighbors())
    for f in range(frames):
        ra = r.update(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f])
        oa = o.update(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f], tagged=r.tagged_cloud())
        assert ra == oa
        for x, y in zip(r.particles(), o.particles()):
            assert same(x, y), "frame %d particles" % f
        for x, y in zip(r.pyramid_lists(), o.pyramid_lists()):
            assert same(x, y), "frame %d pyramid lists" % f
        assert same(r.voxel_objects(), o.voxel_objects()), "frame %d voxel objects" % f
        assert np.array_equal(r.cursors()[:3], o.cursors()[:3])
        for x, y in zip(r.plane_normals(), o.plane_normals()):
            assert same(x, y), "frame %d boundary-plane normals" % f
        if f % 2 == 0:
            (rx, rf), (ox, of) = r.occupancy(0.2), o.occupancy(0.2)
            assert same(rx, ox) and same(rf, of)


def test_voxel_index_round_trip_and_open_interval():
    o = OracleMap(CONFIGS["tiny_dyn"], seed=1)
    for idx in (0, 1, o.nx, o.nx * o.ny, o.V - 1, o.V // 2 + 3):
        c = o.voxel_center(idx)
        assert o.voxel_index(*c) == idx
    half = 0.25 * 16 * 0.5
    assert o.voxel_index(half, 0, 0) == -1 and o.voxel_index(-half, 0, 0) == -1      # open interval (dsp_dynamic.h:1118-1125)
    assert o.voxel_index(np.nextafter(np.float32(half), np.float32(0)), 0, 0) >= 0


def test_pdf_table_is_the_references_expression():
    o = OracleMap(CONFIGS["tiny_dyn"], seed=1)
    t = o.pdf_table()
    assert np.isclose(t[10000], 1.0 / np.sqrt(np.pi), rtol=1e-6)          # 1/sqrt(pi): M_PI_2f32 is pi/2 (dsp_dynamic.h:81-83,1284)
    assert np.array_equal(t[10001:], t[1:10000][::-1])                     # symmetric: lets the GPU keep half of it
    assert np.isclose(t[10000 + 1000], np.exp(-0.5) / np.sqrt(np.pi), rtol=1e-5)


def test_neighbor_table_counts():
    o = OracleMap(CONFIGS["tiny_dyn"], seed=1)
    nb = o.neighbors()
    Nh, Nv = o.Nh, o.Nv
    assert nb[0, 0] == 4 and nb[Nv - 1, 0] == 4 and nb[(Nh - 1) * Nv, 0] == 4     # FOV corners (dsp_dynamic.h:1128-1147)
    assert nb[1, 0] == 6 and nb[Nv, 0] == 6                                         # edges
    assert nb[Nv + 1, 0] == 9 and list(nb[Nv + 1, 1:10]) == [0, 1, 2, Nv, Nv + 1, Nv + 2, 2 * Nv, 2 * Nv + 1, 2 * Nv + 2]
    o2 = OracleMap(CONFIGS["tiny_mn"], seed=1)
    assert o2.neighbors()[:, 0].max() == 25 and o2.neighbors()[0, 0] == 9


def test_observation_overflow_keeps_first_99_and_max_range_of_all():
    cfg = CONFIGS["tiny_dyn"]
    o = OracleMap(cfg, seed=1)
    n = 150
    pts = np.zeros((n, 3), np.float32)
    pts[:, 0] = np.linspace(0.5, 1.9, n)       # one ray straight ahead: all in one pyramid
    pts[:, 1] = 0.01
    pts[:, 2] = 0.01
    assert o.update(pts, (0, 0, 0), 0.0, (1, 0, 0, 0), tagged=np.zeros((0, 7), np.float32)) == 1
    cnt, mx, p = o.observations()
    k = int(np.argmax(cnt))
    assert cnt[k] == 99 and cnt.sum() == 99                                 # dsp_dynamic.h:281-284
    assert np.array_equal(p[k, :99, 0], pts[:99, 0])                        # input order
    assert np.isclose(mx[k], np.linalg.norm(pts[-1]), rtol=1e-6)            # max range sees the dropped points too
    assert o.counters()["n_valid_points"] == n


def test_update_rejects_bad_frames_like_the_reference():
    o = OracleMap(CONFIGS["tiny_dyn"], seed=1)
    z = np.zeros((0, 3), np.float32)
    assert o.update(z, (0, 0, 0), 0.0, (1, 0, 0, 0)) == 1
    assert o.update(z, (0, 0, 0), 0.1, (1.5, 0, 0, 0)) == 0                 # |q| > 1.001 (dsp_dynamic.h:193-196)
    assert o.update(z, (20, 0, 0), 0.2, (1, 0, 0, 0)) == 0                  # pose jump > 10 m (:203)
    assert o.update(z, (0, 0, 0), -5.0, (1, 0, 0, 0)) == 0                  # time runs backwards
    assert o.update(z, (0.1, 0, 0), 0.3, (1, 0, 0, 0)) == 1
