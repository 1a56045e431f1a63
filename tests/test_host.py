"""CPU tests of the host side: the C-ABI library loads and exports everything include/dspmap_b200.h declares, the
config derivation matches the reference's compile-time arithmetic, and the host velocity estimator reproduces the
reference's side thread.  No GPU compute calls here."""
import os
import re

import numpy as np
import pytest

import dspmap_b200 as dm
import refmap
from common import make_stream

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dspmap_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(dspmap_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 35
    L = dm.load_library()
    for s in declared:
        assert hasattr(L, s), "library does not export %s" % s
    assert set(dm.EXPORTED_SYMBOLS) <= set(declared)


def test_no_cpu_fallback():
    from conftest import HAS_GPU
    if HAS_GPU:
        pytest.skip("GPU present")
    with pytest.raises(dm.DSPMapError):
        dm.DSPMap(dm.CONFIGS["tiny_dyn"])


def test_product_does_not_touch_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "dsp-map_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(base, f)).read()
                for pat in (r"import\s+oracle", r"from\s+oracle", r"refmap", r"liboracle", r"oracle/", r"dsp_oracle", r"prefilter_oracle", r"_ref\b"):
                    assert not re.search(pat, src), "%s uses the oracle (%s)" % (f, pat)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg5", "tiny_mn"])
def test_derived_sizes_match_reference_arithmetic(name):
    d = dm.derive(dm.CONFIGS[name])
    expect = {"cfg1": (4000, 40, 504, 36), "cfg2": (174240, 48, 600, 1188), "cfg3": (174240, 48, 5400, 132),
              "cfg5": (1393920, 72, 600, 13966), "tiny_mn": (2560, 12, 84 * 54, 2)}[name]
    assert (d["V"], d["S"], d["P"], d["L"]) == expect
    if refmap.available(name) and name != "cfg5":
        r = refmap.RefMap(name, seed=1)
        assert (r.V, r.S, r.P, r.L, r.T) == (d["V"], d["S"], d["P"], d["L"], d["T"])


@pytest.mark.parametrize("name,frames", [("tiny_dyn", 10), ("tiny_static", 3), ("cfg1", 3), ("ref_default", 3), ("cfg2", 3)])
def test_host_velocity_estimator_equals_reference_side_thread(name, frames):
    if not refmap.available(name):
        pytest.skip("oracle/_ref not built")
    cfg = dm.CONFIGS[name]
    st = make_stream(cfg, seed=3, frames=frames)
    r = refmap.RefMap(name, seed=7)
    e = dm.VelocityEstimator(cfg, seed=7, filter_res=0.1)
    dyn = 0
    for f in range(frames):
        r.update(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f])
        a = r.tagged_cloud()
        b = e.estimate(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f])
        assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), "frame %d" % f
        dyn += int((a[:, 6] > 0.01).sum())
    if name == "tiny_dyn":
        assert dyn > 0  # the stream really exercises clustering + matching


def _brute_force_clusters(xyz, tol, min_size, max_size):
    """O(n^2) connected components of 'distance <= tol' in fp32, PCL's output order (size descending, then smallest member)."""
    n = len(xyz)
    d = xyz[:, None, :] - xyz[None, :, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]   # the estimator's fp32 expression
    adj = d2 <= np.float32(tol) * np.float32(tol)
    label = -np.ones(n, np.int64)
    comps = []
    for s in range(n):
        if label[s] >= 0:
            continue
        stack, members = [s], []
        label[s] = len(comps)
        while stack:
            i = stack.pop()
            members.append(i)
            for j in np.nonzero(adj[i] & (label < 0))[0]:
                label[j] = len(comps)
                stack.append(j)
        comps.append(sorted(members))
    kept = [c for c in comps if min_size <= len(c) <= max_size]
    kept.sort(key=lambda c: (-len(c), c[0]))
    out = -np.ones(n, np.int32)
    for k, c in enumerate(kept):
        out[c] = k
    return len(kept), out


@pytest.mark.parametrize("seed,n,spread", [(1, 600, 1.0), (2, 1500, 3.0), (3, 900, 0.4), (4, 5, 0.1), (5, 1, 0.1)])
def test_clustering_paths_agree_with_brute_force(seed, n, spread):
    rng = np.random.default_rng(seed)
    centres = rng.uniform(-spread, spread, (12, 3)).astype(np.float32)
    xyz = (centres[rng.integers(0, 12, n)] + rng.normal(0, 0.12, (n, 3))).astype(np.float32)
    xyz[: n // 4] = rng.uniform(-spread - 1, spread + 1, (n // 4, 3)).astype(np.float32)      # scattered outliers
    tol = 0.2
    want = _brute_force_clusters(xyz, tol, 5, 400)
    for path in (1, 2, 0):
        got = dm.euclidean_clusters(xyz, tol, 5, 400, path=path)
        assert got[0] == want[0] and np.array_equal(got[1], want[1]), "path %d" % path


def test_clustering_falls_back_to_the_hash_grid_for_huge_extents():
    xyz = np.array([[0, 0, 0], [0.05, 0, 0], [1e6, 1e6, 1e6], [1e6 + 0.05, 1e6, 1e6], [np.nan, 0, 0]], np.float32)[:4]
    with pytest.raises(dm.DSPMapError):
        dm.euclidean_clusters(xyz, 0.2, 1, 10, path=2)
    n, labels = dm.euclidean_clusters(xyz, 0.2, 1, 10, path=0)
    assert n == 2 and list(labels) == [0, 0, 1, 1]


def test_stream_generator_never_hangs_when_the_sensor_enters_an_obstacle():
    # bench.py consumes 210 frames of this stream; at frame 160 the sensor flies into a box and every return is < 0.2 m
    st = make_stream(dm.CONFIGS["cfg2"], seed=1, frames=170, points=2000)
    assert st["points"].shape == (170, 2000, 3) and np.isfinite(st["points"]).all()


def test_estimator_keeps_previous_cloud_when_nothing_in_view():
    cfg = dm.CONFIGS["tiny_dyn"]
    e = dm.VelocityEstimator(cfg, seed=1)
    behind = np.array([[-1.0, 0.0, 0.0], [-2.0, 0.1, 0.0]], np.float32)
    assert e.estimate(behind, (0, 0, 0), 0.0, (1, 0, 0, 0)) is None            # dsp_dynamic.h:1379
    ahead = np.array([[1.0, 0.0, 0.0]], np.float32)
    out = e.estimate(ahead, (0.5, 0, 0), 0.1, (1, 0, 0, 0))                   # z <= filter resolution: "ground", static
    assert out.shape == (1, 7) and np.allclose(out[0], [1.5, 0.0, 0.0, 0, 0, 0, 0])
    lone = e.estimate(ahead, (0, 0, 1.0), 0.2, (1, 0, 0, 0))                  # above ground, cluster of 1 < 5: dropped
    assert lone.shape == (0, 7)


def test_estimator_on_the_helper_thread_equals_the_calling_thread():
    """By default (DSPMAP_EST_THREAD=0 turns it off) the library hands the estimation to a persistent helper thread (host_worker.h); the hand-over must not
    change a bit, must survive being switched on and off, and must not lose a wake-up in either of its two waiting modes
    (spinning right after a job, sleeping on the condition variable after 2 ms without one)."""
    import time
    cfg = dm.CONFIGS["tiny_dyn"]
    st = make_stream(cfg, seed=3, frames=40)
    a = dm.VelocityEstimator(cfg, seed=7, filter_res=0.1)
    b = dm.VelocityEstimator(cfg, seed=7, filter_res=0.1)
    assert b.set_threaded(True) == 1   # DSPMAP_OK
    for f in range(40):
        if f == 20:
            b.set_threaded(False)
        if f == 25:
            b.set_threaded(True)
        if f in (5, 30):
            time.sleep(0.01)   # the helper has gone to sleep by now
        x = a.estimate(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f])
        y = b.estimate(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f])
        assert (x is None) == (y is None)
        if x is not None:
            assert x.shape == y.shape and np.array_equal(x.view(np.uint32), y.view(np.uint32)), "frame %d" % f
    # many short jobs back to back (the spinning hand-over) interleaved with idle gaps (the sleeping hand-over)
    pts = st["points"][0][:32]
    for k in range(3000):
        b.estimate(pts, (0, 0, 0), 100.0 + 0.1 * k, (1, 0, 0, 0))
        if k % 500 == 499:
            time.sleep(0.005)


def test_pdf_index_clamp_can_move_behind_the_conversion():
    """dsp_pdf_i (dspmap_frame.cuh) computes queryNormalPDF's table index (dsp_dynamic.h:1294-1300) as
    min(|trunc(cx*1000 + 10000) - 10000|, 9900) with a saturating conversion instead of clamping cx to +-9.9 first.  The two
    agree for every finite float; the full sweep over all 2^32 bit patterns takes five minutes in numpy (it was run once,
    0 mismatches), here every float with 9 <= |cx| <= 11 (where the clamp acts), every 4099th bit pattern elsewhere and the
    extremes are checked."""
    f32 = np.float32

    def both(cx):
        c = np.where(cx > f32(9.9), f32(9.9), np.where(cx < f32(-9.9), f32(-9.9), cx)).astype(f32)
        with np.errstate(all="ignore"):
            i1 = np.abs(np.trunc((c * f32(1000) + f32(10000)).astype(f32)).astype(np.int64) - 10000)
            t2 = (cx * f32(1000) + f32(10000)).astype(f32)
        i = np.clip(np.trunc(np.clip(np.nan_to_num(t2, nan=0.0, posinf=3e9, neginf=-3e9), -2147483648.0, 2147483647.0)).astype(np.int64),
                    -2147483648, 2147483647)
        j = ((i - 10000 + 2 ** 31) % 2 ** 32) - 2 ** 31   # 32-bit wrap-around of the subtraction
        return i1, np.minimum(np.abs(j), 9900)

    lo, hi = np.array([9.0, 11.0], f32).view(np.uint32)
    near = np.arange(lo, hi + 1, dtype=np.uint32)
    sparse = np.arange(0, 0x7f800000, 4099, dtype=np.uint32)
    extremes = np.array([0, 1, 0x007fffff, 0x00800000, 0x7f7fffff, 0x4f000000, 0x4effffff, 0x5f000000], np.uint32)
    for bits in (near, sparse, extremes):
        for sign in (0, 0x80000000):
            cx = (bits | np.uint32(sign)).view(f32)
            a, b = both(cx)
            assert np.array_equal(a, b)
