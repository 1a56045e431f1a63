#!/usr/bin/env python3
"""GPU parity diagnosis (not a pytest file): runs the reference and the CUDA library on one stream; at the first
mismatching frame replays that frame from the reference's pre-state with stage limits 1..4 on the restatement oracle
and on the GPU, so the failing stage is named in one run.

    python tests/gpu_debug.py tiny_dyn 12 [init_particles]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dsp-map_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import dspmap_b200 as dm  # noqa: E402
from dspmap_b200.streams import make_stream  # noqa: E402
from oracle import OracleMap  # noqa: E402
from parity import compare_state  # noqa: E402
from refmap import RefMap  # noqa: E402


def setters(g):
    g.setPredictionVariance(0.05, 0.05)
    g.setObservationStdDev(0.1)
    g.setNewBornParticleNumberofEachPoint(20)
    g.setNewBornParticleWeight(1e-4)
    g.setOriginalVoxelFilterResolution(0.1)


def main():
    name = sys.argv[1]
    F = int(sys.argv[2])
    initp = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    cfg = dm.CONFIGS[name]
    st = make_stream(cfg, seed=3, frames=F)
    ref = RefMap(name, seed=7, init_particles=initp)
    gpu = dm.DSPMap(cfg, seed=7, init_particle_num=initp)
    setters(gpu)
    pre = None
    for f in range(F):
        pts, pos, t, q = st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]
        pre = (ref.particles(), ref.cursors().copy())
        a = ref.update(pts, pos, t, q)
        tc = ref.tagged_cloud()
        b = gpu.update(len(pts), 3, pts, float(pos[0]), float(pos[1]), float(pos[2]), float(t), float(q[0]), float(q[1]),
                       float(q[2]), float(q[3]), tagged=tc)
        bad = compare_state(ref, gpu, label="frame %d:" % f)
        c = gpu.counters()
        print("frame", f, "rc", a, b, "OK" if not bad else "MISMATCH", {k: c[k] for k in ("n_in", "n_moved", "n_voxel_full", "n_pyramid_full", "n_fov", "n_candidates", "n_born", "n_out", "launches_frame")}, flush=True)
        if f % 2 == 0:
            rx, rf = ref.occupancy(0.2)
            n, gx, gf = gpu.getOccupancyMapWithFutureStatus(0.2)
            if not (rx.shape == gx.shape and np.array_equal(rx, gx)):
                print("  reader: occupied list differs", len(rx), n)
            if not np.allclose(rf, gf, rtol=2e-6, atol=0):
                print("  reader: future differs", float(np.abs(rf - gf).max()))
        if bad:
            for x in bad:
                print("   ", x)
            print("--- replaying frame %d by stage from the reference's pre-state" % f)
            (ids, vals), cur = pre
            for k in (1, 2, 3, 4):
                o = OracleMap(cfg, seed=7)
                g = dm.DSPMap(cfg, seed=7)
                setters(g)
                for m_ in (o, g):
                    m_.load_particles(ids, vals)
                    m_.set_cursors(cur[0], cur[1], cur[2])
                    m_.set_stage_limit(k)
                    if f > 0:
                        m_.set_last_pose(st["pos"][f - 1], st["t"][f - 1])
                o.update(pts, pos, t, q, tagged=tc)
                g.update(len(pts), 3, pts, float(pos[0]), float(pos[1]), float(pos[2]), float(t), float(q[0]), float(q[1]),
                         float(q[2]), float(q[3]), tagged=tc)
                bb = compare_state(o, g, label="stage<=%d:" % k)
                print(" stage limit", k, "OK" if not bb else "MISMATCH", "oracle", {kk: vv for kk, vv in o.counters().items() if vv}, "gpu", {kk: vv for kk, vv in g.counters().items() if vv})
                for x in bb:
                    print("     ", x)
                g.close()
                if bb:
                    break
            return 1
    print("ALL FRAMES BIT-EXACT")
    return 0


if __name__ == "__main__":
    sys.exit(main())
