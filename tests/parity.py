"""Shared helpers of the parity tests: run an oracle (reference .so or restatement) and the CUDA library on the same
stream and compare every piece of state bit for bit."""
import numpy as np


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def same(a, b):
    return a.shape == b.shape and np.array_equal(bits(a), bits(b))


def compare_state(ref, gpu, check_future_exact=False, future_rtol=2e-6, label=""):
    """ref: RefMap or OracleMap; gpu: dspmap_b200.DSPMap. Returns a list of human-readable mismatch descriptions."""
    bad = []
    if hasattr(ref, "plane_normals") and hasattr(gpu, "plane_normals"):
        # the rotated boundary-plane normals decide the pyramid of every particle and point; a 1-ulp difference flips one particle
        # in a few hundred thousand (that is how a fp32 / fp64 pi mix-up showed, 36 frames into the cfg2 stream)
        (rh, rv), (gh, gv) = ref.plane_normals(), gpu.plane_normals()
        if not (same(rh, gh) and same(rv, gv)):
            bad.append("%s boundary-plane normals differ in %d words" % (label, int((bits(rh) != bits(gh)).sum() + (bits(rv) != bits(gv)).sum())))
    rc, rm, rp = ref.observations()
    gc, gm, gp = gpu.observations()
    if not same(rc, gc):
        bad.append("%s obs counts differ at pyramids %s" % (label, np.nonzero(rc != gc)[0][:8]))
    if not same(rm, gm):
        bad.append("%s obs max range differs at pyramids %s" % (label, np.nonzero(bits(rm) != bits(gm))[0][:8]))
    if same(rc, gc):
        m = np.arange(rp.shape[1])[None, :] < rc[:, None]
        for k, name in ((0, "x"), (1, "y"), (2, "z"), (4, "range"), (3, "Cz")):
            a, b = rp[..., k][m], gp[..., k][m]
            if not same(a, b):
                w = np.nonzero(bits(a) != bits(b))[0]
                bad.append("%s obs %s differs in %d of %d (first: ref %r gpu %r)" % (label, name, len(w), a.size, a[w[0]], b[w[0]]))
    ro, re = ref.pyramid_lists()
    go, ge = gpu.pyramid_lists()
    if not same(ro, go):
        w = np.nonzero(np.diff(ro) != np.diff(go))[0]
        bad.append("%s pyramid list lengths differ at %s (ref %s gpu %s)" % (label, w[:6], np.diff(ro)[w[:6]], np.diff(go)[w[:6]]))
    elif not same(re, ge):
        w = np.nonzero((re != ge).any(1))[0]
        bad.append("%s pyramid list entries differ in %d of %d (first idx %d ref %s gpu %s)" % (label, len(w), len(re), w[0], re[w[0]], ge[w[0]]))
    ri, rv = ref.particles()
    gi, gv = gpu.particles()
    if not same(ri, gi):
        sr = set(map(tuple, ri.tolist()))
        sg = set(map(tuple, gi.tolist()))
        bad.append("%s particle (voxel,slot) sets differ: ref %d gpu %d, only-ref %s only-gpu %s" %
                   (label, len(ri), len(gi), sorted(sr - sg)[:5], sorted(sg - sr)[:5]))
    else:
        names = ["flag", "vx", "vy", "vz", "px", "py", "pz", "w"]
        for k, nme in enumerate(names):
            if not same(rv[:, k], gv[:, k]):
                w = np.nonzero(bits(rv[:, k]) != bits(gv[:, k]))[0]
                bad.append("%s particle %s differs in %d of %d (first at %s: ref %r gpu %r)" %
                           (label, nme, len(w), len(rv), ri[w[0]], rv[w[0], k], gv[w[0], k]))
    rvo, gvo = ref.voxel_objects(), gpu.voxel_objects()
    if not same(rvo[:, :4], gvo[:, :4]):
        w = np.nonzero((bits(rvo[:, :4]) != bits(gvo[:, :4])).any(1))[0]
        bad.append("%s occupancy / mean velocity differs in %d voxels (first %d ref %s gpu %s)" % (label, len(w), w[0], rvo[w[0], :4], gvo[w[0], :4]))
    fr, fg = rvo[:, 4:], gvo[:, 4:]
    if check_future_exact:
        if not same(fr, fg):
            bad.append("%s future status not bit-identical" % label)
    else:
        # future sums are accumulated with atomics on the GPU (order differs): same support, tight tolerance
        if not np.array_equal(fr != 0, fg != 0):
            bad.append("%s future status support differs in %d cells" % (label, int(((fr != 0) != (fg != 0)).sum())))
        elif not np.allclose(fr, fg, rtol=future_rtol, atol=0):
            bad.append("%s future status differs beyond rtol %g (max rel %g)" % (label, future_rtol, float(np.max(np.abs(fr - fg) / np.maximum(np.abs(fr), 1e-30)))))
    rcu, gcu = ref.cursors(), gpu.cursors()
    if not np.array_equal(rcu[:3], gcu[:3]):
        bad.append("%s cursors differ: ref %s gpu %s" % (label, rcu[:3], gcu[:3]))
    return bad


def run_stream(ref, gpu, stream, frames, tagged_from_ref=True, reader_every=2, threshold=0.2, stop_on_first=True, seen=None):
    """Feeds the same frames to both; the newborn input of the GPU map is the reference's own tagged cloud, so the
    comparison isolates the hot path. Returns list of (frame, mismatches)."""
    out = []
    for f in range(frames):
        pts, pos, t, q = stream["points"][f][: stream["n"][f]], stream["pos"][f], stream["t"][f], stream["quat"][f]
        a = ref.update(pts, pos, t, q) if not hasattr(ref, "load_particles") else None
        if a is None:
            raise RuntimeError("run_stream expects a RefMap as reference")
        tc = ref.tagged_cloud()
        b = gpu.update(len(pts), 3, pts, float(pos[0]), float(pos[1]), float(pos[2]), float(t), float(q[0]), float(q[1]),
                       float(q[2]), float(q[3]), tagged=tc if tagged_from_ref else None)
        bad = []
        if seen is not None:  # which regimes the stream went through (per-frame counters of the CUDA path, summed)
            for k, v in gpu.counters().items():
                seen[k] = seen.get(k, 0) + v if k.startswith("n_") else max(seen.get(k, 0), v)
        if a != b:
            bad.append("return codes differ: ref %d gpu %d" % (a, b))
        bad += compare_state(ref, gpu, label="frame %d:" % f)
        if reader_every and f % reader_every == 0:
            rx, rf = ref.occupancy(threshold)
            n, gx, gf = gpu.getOccupancyMapWithFutureStatus(threshold)
            if not same(rx, gx):
                bad.append("frame %d: occupied voxel list differs (ref %d gpu %d)" % (f, len(rx), n))
            if not np.allclose(rf, gf, rtol=2e-6, atol=0):
                bad.append("frame %d: future status copy-out differs" % f)
        if bad:
            out.append((f, bad))
            if stop_on_first:
                break
    return out
