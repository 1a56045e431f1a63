"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI, against
  * the golden vectors produced by the unmodified reference (tests/golden),
  * the reference itself (oracle/_ref/*.so travels to the GPU box) on identical streams, frame by frame,
  * the restatement oracle for cases the compiled reference configs do not cover,
and, at BASELINE.json's full sizes, size-independent properties of the map state.
Bar: bit-exact for voxel ids, slot ids, pyramid lists, positions, velocities, weights, occupancy, mean velocity, noise
cursors and the occupied-voxel list; the future-status grid is accumulated with float atomics on the GPU, so it must have
the same support and agree within rtol 2e-6 (sums of <= a few hundred positive fp32 terms in a different order)."""
import numpy as np
import pytest

import dspmap_b200 as dm
import refmap
from common import GOLDEN, SET, check_against_golden, gpu_map, gpu_update, load_golden, make_stream
from oracle import OracleMap
from parity import compare_state, run_stream, same

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,frames,initp", GOLDEN)
def test_gpu_reproduces_reference_golden(name, frames, initp):
    G = load_golden(name, frames, initp)
    g = gpu_map(name, seed=int(G["seed"]), init_particles=initp)
    for f in range(frames):
        rc = gpu_update(g, G["points"][f], G["pos"][f], G["t"][f], G["quat"][f], tagged=G["tagged_%d" % f])
        assert rc == int(G["rc_%d" % f])
        check_against_golden(g, G, f, is_gpu=True)
        if f % 2 == 1:
            n, xyz, fut = g.getOccupancyMapWithFutureStatus(0.2)
            assert same(xyz, G["occ_xyz_%d" % f])
            ref = np.zeros(fut.size, np.float32)
            ref[G["fut_idx_%d" % f]] = G["fut_val_%d" % f]
            assert np.array_equal(fut.ravel() != 0, ref != 0) and np.allclose(fut.ravel(), ref, rtol=2e-6, atol=0)
    g.close()


STREAMS = [("tiny_dyn", 14, 0), ("tiny_static", 8, 0), ("tiny_dyn", 6, 3000), ("cfg1", 6, 0), ("cfg2", 36, 0), ("cfg3", 8, 0),
           ("cfg4", 8, 0), ("cfg5", 2, 0), ("ref_default", 3, 0)]


@pytest.mark.parametrize("name,frames,initp", STREAMS)
def test_gpu_equals_reference_on_stream(name, frames, initp):
    if not refmap.available(name):
        pytest.skip("oracle/_ref/libdspref_%s.so not present" % name)
    cfg = dm.CONFIGS[name]
    # cfg2: bench.py's own stream (seed 1), through frame 35: voxels fill up in frames 4-13 and 33, as in the timed frames
    st = make_stream(cfg, seed=1 if name == "cfg2" else 3, frames=frames)
    r = refmap.RefMap(name, seed=7, init_particles=initp)
    g = gpu_map(name, seed=7, init_particles=initp, max_points=max(cfg["points"], 1024))
    seen = {}
    bad = run_stream(r, g, st, frames, seen=seen)
    g.close()
    assert not bad, "\n".join(str(b) for b in bad)
    assert seen["overflow"] == 0
    if name == "cfg2":  # the stream reaches the saturated population bench.py times: full voxels and dropped light particles
        assert seen["n_voxel_full"] > 0 and seen["n_low_weight"] > 0 and seen["n_born"] > 0, seen


def test_gpu_builtin_velocity_estimation_path():
    """update() without an explicit newborn input: the library's own host estimator feeds the newborn kernels."""
    name = "tiny_dyn"
    st = make_stream(dm.CONFIGS[name], seed=4, frames=8)
    r = refmap.RefMap(name, seed=7) if refmap.available(name) else None
    if r is None:
        pytest.skip("reference library not present")
    g = gpu_map(name, seed=7)
    bad = run_stream(r, g, st, 8, tagged_from_ref=False)
    assert same(r.tagged_cloud(), g.getKMClusterResult())
    g.close()
    assert not bad, "\n".join(str(b) for b in bad)


def test_pyramid_overflow_is_exact_including_slot_reuse():
    """tiny_mn: the reference's own size formula gives 2 list slots per pyramid, so lists overflow every frame and the
    dropped particles free their voxel slots in the middle of the sweep (dsp_dynamic.h:1256-1259), where later arrivals of
    the same frame take them.  k_arrive replays such frames in sweep order: everything stays bit-identical."""
    name = "tiny_mn"
    if not refmap.available(name):
        pytest.skip("reference library not present")
    frames = 12
    st = make_stream(dm.CONFIGS[name], seed=3, frames=frames)
    r = refmap.RefMap(name, seed=7)
    g = gpu_map(name, seed=7)
    bad = run_stream(r, g, st, frames)
    assert g.counters()["n_pyramid_full"] > 0
    g.close()
    assert not bad, "\n".join(str(b) for b in bad)


@pytest.mark.parametrize("initp", [0, 4000])
def test_pyramid_and_voxel_overflow_together_are_exact(initp):
    """The same map seeded with constructor particles (voxels fill up as well): voxel-full and pyramid-full drops interleave."""
    name = "tiny_mn"
    if not refmap.available(name):
        pytest.skip("reference library not present")
    frames = 6
    st = make_stream(dm.CONFIGS[name], seed=5, frames=frames)
    r = refmap.RefMap(name, seed=11, init_particles=initp)
    g = gpu_map(name, seed=11, init_particles=initp)
    bad = run_stream(r, g, st, frames)
    g.close()
    assert not bad, "\n".join(str(b) for b in bad)


def test_edge_cases_empty_cloud_rejected_frames_and_stride():
    cfg = dm.CONFIGS["tiny_dyn"]
    o = OracleMap(cfg, seed=3)
    g = gpu_map("tiny_dyn", seed=3)
    st = make_stream(cfg, seed=9, frames=3)
    est = dm.VelocityEstimator(cfg, seed=3)
    z = np.zeros((0, 3), np.float32)
    for (pts, pos, t, q) in [(st["points"][0], st["pos"][0], 0.0, st["quat"][0]), (z, st["pos"][1], 0.1, st["quat"][1]),
                             (st["points"][2], st["pos"][2], 0.2, (1.5, 0, 0, 0)), (st["points"][2], (50, 0, 0), 0.3, st["quat"][2]),
                             (st["points"][2], st["pos"][2], 0.05, st["quat"][2]), (st["points"][2], st["pos"][2], 0.4, st["quat"][2])]:
        tc = est.estimate(pts, pos, t, q) if len(pts) else None
        a = o.update(pts, pos, t, q, tagged=tc)
        b = gpu_update(g, pts, pos, t, q, tagged=tc) if tc is not None else g._check(g.lib.dspmap_update_tagged(
            g.h, len(pts), 3, None if len(pts) == 0 else pts.ctypes.data_as(dm.C.POINTER(dm.C.c_float)), float(pos[0]), float(pos[1]),
            float(pos[2]), float(t), float(q[0]), float(q[1]), float(q[2]), float(q[3]), None, 0))
        assert a == b
        if a == 1:
            assert not compare_state(o, g)
    # stride 5 input: only the first three floats of each point are used (dsp_dynamic.h:247,289)
    o2, g2 = OracleMap(cfg, seed=3), gpu_map("tiny_dyn", seed=3)
    p5 = np.zeros((400, 5), np.float32)
    p5[:, :3] = st["points"][0]
    p5[:, 3:] = 7.0
    tc = dm.VelocityEstimator(cfg, seed=3).estimate(st["points"][0], st["pos"][0], 0.0, st["quat"][0])
    o2.update(p5, st["pos"][0], 0.0, st["quat"][0], tagged=tc, stride=5)
    gpu_update(g2, p5, st["pos"][0], 0.0, st["quat"][0], tagged=tc, stride=5)
    assert not compare_state(o2, g2)
    g.close()
    g2.close()


def test_readers_threshold_variants_clear_and_state_round_trip():
    cfg = dm.CONFIGS["tiny_dyn"]
    st = make_stream(cfg, seed=6, frames=5)
    o, g = OracleMap(cfg, seed=2), gpu_map("tiny_dyn", seed=2)
    est = dm.VelocityEstimator(cfg, seed=2)
    for f in range(5):
        tc = est.estimate(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f])
        o.update(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f], tagged=tc)
        gpu_update(g, st["points"][f], st["pos"][f], st["t"][f], st["quat"][f], tagged=tc)
    for thr in (0.0, 0.05, 0.2, 0.7, 100.0):
        ox, _ = o.occupancy(thr, with_future=False)
        n, gx = g.getOccupancyMap(thr)
        assert n == len(ox) and same(ox, gx)
    # the readers zero the future columns (dsp_dynamic.h:397-400,421-424)
    assert not g.voxel_objects()[:, 4:].any() and not o.voxel_objects()[:, 4:].any()
    tc = est.estimate(st["points"][4], st["pos"][4], 0.5, st["quat"][4])
    o.update(st["points"][4], st["pos"][4], 0.5, st["quat"][4], tagged=tc)
    gpu_update(g, st["points"][4], st["pos"][4], 0.5, st["quat"][4], tagged=tc)
    assert g.voxel_objects()[:, 4:].any()
    g.clearOccupancyMapPrediction()
    o.clear_prediction()
    assert not g.voxel_objects()[:, 4:].any()
    # dump -> load into a fresh map -> identical continuation
    ids, vals = g.particles()
    g2 = gpu_map("tiny_dyn", seed=2)
    g2.load_particles(ids, vals)
    c = g.cursors()
    g2.set_cursors(c[0], c[1], c[2])
    g2.set_last_pose(st["pos"][4], 0.5)
    tc = est.estimate(st["points"][3], st["pos"][3], 0.6, st["quat"][3])
    for m in (g, g2):
        gpu_update(m, st["points"][3], st["pos"][3], 0.6, st["quat"][3], tagged=tc)
    a, b = g.particles(), g2.particles()
    assert same(a[0], b[0]) and same(a[1], b[1])
    for k in range(3):
        assert g.getVoxelPositionFromIndexPublic(k * 37).tolist() == o.voxel_center(k * 37).tolist()
    assert g.getPointVoxelsIndexPublic(0.3, -0.2, 0.1) == (1, o.voxel_index(0.3, -0.2, 0.1))
    g.close()
    g2.close()


@pytest.mark.parametrize("name,frames", [("tiny_dyn", 8), ("cfg2", 8)])
def test_device_resident_update_equals_host_update(name, frames):
    """dspmap_update_device (cloud and newborn input already in HBM, nothing synchronises, the first newborn kernels run on
    the side branch) leaves the same map as dspmap_update_tagged with the same inputs — the path bench.py's `value` times."""
    import torch
    cfg = dm.CONFIGS[name]
    st = make_stream(cfg, seed=8, frames=frames)
    a, b = gpu_map(name, seed=4), gpu_map(name, seed=4)
    est = dm.VelocityEstimator(cfg, seed=4)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    b.set_stream(stream.cuda_stream)
    last = np.zeros((0, 7), np.float32)
    d_xyz = torch.zeros((b.V, 3), dtype=torch.float32, device=dev)
    d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    d_fut = torch.zeros((b.V, b.T), dtype=torch.float32, device=dev)
    for f in range(frames):
        pts, pos, t, q = st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]
        tc = est.estimate(pts, pos, t, q)
        last = tc if tc is not None else last
        assert gpu_update(a, pts, pos, t, q, tagged=last) == 1
        with torch.cuda.stream(stream):
            d_p = torch.from_numpy(np.ascontiguousarray(pts, np.float32)).to(dev)
            d_t = torch.from_numpy(last if len(last) else np.zeros((1, 7), np.float32)).to(dev)
            assert b.update_device(len(pts), d_p.data_ptr(), pos, t, q, d_t.data_ptr(), len(last)) == 1
            if f % 2:
                b.get_occupancy_device(0.2, d_xyz.data_ptr(), b.V, d_cnt.data_ptr(), d_fut.data_ptr())
        stream.synchronize()
        if f % 2:
            n, xyz, _ = a.getOccupancyMapWithFutureStatus(0.2)
            assert n == int(d_cnt.item()) and same(xyz, d_xyz[:n].cpu().numpy())
        (ia, va), (ib, vb) = a.particles(), b.particles()
        assert same(ia, ib) and same(va, vb), "frame %d" % f
        assert np.array_equal(a.cursors(), b.cursors()), "cursors, frame %d" % f
    assert len(ia) > 100
    a.close()
    b.close()


@pytest.mark.parametrize("name,frames", [("tiny_dyn", 8), ("cfg2", 6)])
def test_pipelined_reader_equals_blocking_reader(name, frames):
    """dspmap_get_occupancy_async / dspmap_wait_occupancy (copies overlapped with the next update) return what the
    reference-facing blocking reader returns, frame for frame, including the zeroing of the future columns."""
    cfg = dm.CONFIGS[name]
    st = make_stream(cfg, seed=4, frames=frames)
    a, b = gpu_map(name, seed=3), gpu_map(name, seed=3)
    est = dm.VelocityEstimator(cfg, seed=3)
    fut = np.zeros((a.V, a.T), np.float32)
    want, got, ticket = [], [], None
    for f in range(frames):
        tc = est.estimate(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f])
        for m in (a, b):
            gpu_update(m, st["points"][f], st["pos"][f], st["t"][f], st["quat"][f], tagged=tc)
        n, xyz, _ = a.getOccupancyMapWithFutureStatus(0.2, fut)
        want.append((n, xyz.copy(), fut.copy()))
        prev, ticket = ticket, b.get_occupancy_async(0.2, with_future=(f % 3 != 2))
        if prev is not None:   # frame f-1's results are collected while frame f's copies are in flight
            n2, x2, f2 = b.wait_occupancy(prev)
            got.append((n2, x2.copy(), None if f2 is None else f2.copy()))
    n2, x2, f2 = b.wait_occupancy(ticket)
    got.append((n2, x2.copy(), None if f2 is None else f2.copy()))
    assert len(got) == frames
    for f, ((n, xyz, fu), (n2, x2, f2)) in enumerate(zip(want, got)):
        assert n == n2 and same(xyz, x2), "frame %d" % f
        assert (f2 is None) == (f % 3 == 2)
        if f2 is not None:
            # fp32 atomics: identical support, values to rounding (DESIGN.md "Result contract")
            assert np.array_equal(fu != 0, f2 != 0) and np.allclose(fu, f2, rtol=2e-6, atol=0), "future status, frame %d" % f
    assert not b.voxel_objects()[:, 4:].any()
    with pytest.raises(dm.DSPMapError):
        b.wait_occupancy(5)
    a.close()
    b.close()


@pytest.mark.parametrize("name,frames", [("cfg2", 30), ("cfg5", 4)])
def test_full_size_properties(name, frames):
    """BASELINE.json sizes, no oracle in the loop: invariants every reference state satisfies."""
    cfg = dm.CONFIGS[name]
    st = make_stream(cfg, seed=1, frames=frames)
    g = gpu_map(name, seed=1, max_points=cfg["points"])
    d = dm.derive(cfg)
    for f in range(frames):
        assert gpu_update(g, st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]) == 1
        if f in (frames // 2, frames - 1):
            c = g.counters()
            ids, vals = g.particles()
            assert len(ids) == c["n_out"] and c["overflow"] == 0
            key = ids[:, 0].astype(np.int64) * 128 + ids[:, 1]
            assert np.all(np.diff(key) > 0) and ids[:, 1].max() < d["S"]            # unique (voxel, slot), sweep order
            half = 0.5 * cfg["res"] * np.array([cfg["nx"], cfg["ny"], cfg["nz"]], np.float32)
            cell = ((vals[:, 4:7] + half) / np.float32(cfg["res"])).astype(np.int64)
            vox = cell[:, 2] * cfg["ny"] * cfg["nx"] + cell[:, 1] * cfg["nx"] + cell[:, 0]
            assert np.array_equal(vox, ids[:, 0])                                   # every particle sits in its voxel
            assert np.all(vals[:, 3] == 0) and np.all(np.isin(vals[:, 0], (np.float32(1.0), np.float32(0.6))))
            cnt = np.bincount(ids[:, 0], minlength=d["V"])
            assert cnt.max() <= max(cfg["max_ppv"], 4)                              # resampled down to MAX ppv
            vo = g.voxel_objects()
            wsum = np.bincount(ids[:, 0], weights=vals[:, 7].astype(np.float64), minlength=d["V"])
            assert np.allclose(wsum, vo[:, 0], rtol=1e-4, atol=1e-7)                # resampling preserves voxel weight
            n, xyz, fut = g.getOccupancyMapWithFutureStatus(0.2)
            occ = np.nonzero(vo[:, 0] > 0.2)[0]
            assert n == len(occ)
            centres = np.stack([g.getVoxelPositionFromIndexPublic(i) for i in occ[:50]]) if n else np.zeros((0, 3))
            assert np.array_equal(xyz[:50], centres.astype(np.float32))             # ascending voxel order
            assert np.array_equal(fut, vo[:, 4:]) and fut.sum() > 0
            n2, _, fut2 = g.getOccupancyMapWithFutureStatus(0.2)
            assert n2 == n and not fut2.any()                                       # idempotent list, future cleared
    g.close()


@pytest.mark.parametrize("switch,value", [("DSPMAP_PDL", "0"), ("DSPMAP_EST_THREAD", "0"), ("DSPMAP_ASYNC_UPDATE", "0"), ("DSPMAP_EST_GPU", "1"),
                                          ("DSPMAP_NORM_POLL", "0"), ("DSPMAP_NB_POS", "2")])
def test_library_switches_do_not_change_a_bit(switch, value, monkeypatch):
    """The library's run-time switches (read by dspmap_create) select HOW a frame is executed —
    programmatic dependent launch, helper-thread / calling-thread host estimation, asynchronous update, the estimation front
    end on the device, the newborn normaliser beside / behind the C_z pass, where the early newborn kernels are enqueued —
    never what it computes: a map created with one switched must stay bit-identical to a default map on the bench workload,
    through the explicit-newborn-input path and through the library's own estimator."""
    name, frames = "cfg2", 3
    cfg = dm.CONFIGS[name]
    st = make_stream(cfg, seed=3, frames=frames + 2)
    est = dm.VelocityEstimator(cfg, seed=7, filter_res=0.1)
    monkeypatch.delenv(switch, raising=False)
    a = gpu_map(name, seed=7, max_points=cfg["points"])
    monkeypatch.setenv(switch, value)
    b = gpu_map(name, seed=7, max_points=cfg["points"])
    monkeypatch.delenv(switch, raising=False)
    bad = []
    for f in range(frames + 2):
        pts, pos, t, q = st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]
        tc = est.estimate(pts, pos, t, q) if f < frames else None   # the last two frames: built-in estimation
        assert gpu_update(a, pts, pos, t, q, tagged=tc) == gpu_update(b, pts, pos, t, q, tagged=tc) == 1
        bad += compare_state(a, b, label="%s frame %d:" % (switch, f))
        if tc is None:
            assert same(a.getKMClusterResult(), b.getKMClusterResult())
    na, xa, fa = a.getOccupancyMapWithFutureStatus(0.2)
    nb, xb, fb = b.getOccupancyMapWithFutureStatus(0.2)
    assert same(xa, xb) and np.allclose(fa, fb, rtol=2e-6, atol=0)
    a.close()
    b.close()
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("name,frames,seed", [("tiny_dyn", 10, 4), ("tiny_static", 6, 2), ("cfg2", 12, 1), ("cfg3", 4, 3)])
def test_device_velocity_estimation_equals_host_estimation(name, frames, seed, monkeypatch):
    """update() with the estimation front end on the device (dspmap_estimator.cuh: FOV filter, ground split, hash-grid
    union-find clustering, centroids, layout of the tagged cloud; Hungarian matching on the host) against the same map with
    the host implementation (the default), which test_host.py pins against the reference's own side thread: the tagged
    cloud (positions, velocities, colours, order) and the map state must be bit-identical after every frame — including a
    frame with nothing in view (the previous cloud is kept) and an empty cloud."""
    cfg = dm.CONFIGS[name]
    st = make_stream(cfg, seed=seed, frames=frames)
    monkeypatch.setenv("DSPMAP_EST_GPU", "1")
    a = gpu_map(name, seed=7, max_points=cfg["points"])
    monkeypatch.delenv("DSPMAP_EST_GPU", raising=False)
    b = gpu_map(name, seed=7, max_points=cfg["points"])
    assert a.estimator_stats()[0] == 1 and b.estimator_stats()[0] == 0
    bad, moving = [], 0
    for f in range(frames):
        pts, pos, t, q = st["points"][f].copy(), st["pos"][f], st["t"][f], st["quat"][f]
        if f == frames // 2:      # everything behind the sensor: nothing in view
            pts[:, 0] = -np.abs(pts[:, 0]) - 1.0
            pts[:, 1:] = 0.0
        if f == frames // 2 + 1:  # an empty cloud
            pts = pts[:0]
        assert gpu_update(a, pts, pos, t, q) == gpu_update(b, pts, pos, t, q) == 1
        ta, tb = a.getKMClusterResult(), b.getKMClusterResult()
        if not same(ta, tb):
            bad.append("%s frame %d: tagged clouds differ (%s vs %s)" % (name, f, ta.shape, tb.shape))
        moving += int(((tb[:, 6] > 0.01) & (tb[:, 3] > -100)).sum()) if len(tb) else 0
        bad += compare_state(b, a, label="%s frame %d:" % (name, f))
    a.close()
    b.close()
    assert not bad, "\n".join(bad)
    if name == "cfg2":
        assert moving > 0  # the streams contain moving boxes: some points carried an estimated velocity
