#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libdspref_*.so, i.e. the reference header
compiled against the stand-in dependency headers — see oracle/build_ref.py).  Run where /root/reference exists:

    python oracle/build_ref.py && python tests/golden/make_golden.py

Each fixture holds the inputs of a short stream (sensor-frame clouds, poses, time stamps), the newborn input the
reference's side thread produced per frame, and the reference's state after every frame (particle store, occupancy /
mean-velocity / future grid, pyramid lists, binned observations, cursors) plus the reader outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "dsp-map_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from dspmap_b200.configs import CONFIGS  # noqa: E402
from dspmap_b200.streams import make_stream  # noqa: E402
from refmap import RefMap  # noqa: E402

CASES = [("tiny_dyn", 6, 0), ("tiny_static", 5, 0), ("tiny_dyn", 4, 1500)]


def main():
    for name, frames, initp in CASES:
        cfg = CONFIGS[name]
        st = make_stream(cfg, seed=11, frames=frames)
        r = RefMap(name, seed=5, init_particles=initp)
        out = dict(points=st["points"], pos=st["pos"], quat=st["quat"], t=st["t"], seed=np.int64(5), init_particles=np.int64(initp),
                   pdf=r.pdf_table(), neighbors=r.neighbors())
        for f in range(frames):
            out["rc_%d" % f] = np.int32(r.update(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]))
            out["tagged_%d" % f] = r.tagged_cloud()
            ids, vals = r.particles()
            out["ids_%d" % f], out["vals_%d" % f] = ids, vals
            vo = r.voxel_objects()
            nz = np.nonzero(np.any(vo != 0, axis=1))[0].astype(np.int32)
            out["vox_idx_%d" % f], out["vox_val_%d" % f] = nz, vo[nz]
            off, ent = r.pyramid_lists()
            out["pyr_off_%d" % f], out["pyr_ent_%d" % f] = off, ent
            cnt, mx, pts = r.observations()
            m = np.arange(pts.shape[1])[None, :] < cnt[:, None]
            out["obs_cnt_%d" % f], out["obs_max_%d" % f], out["obs_pts_%d" % f] = cnt, mx, pts[m]
            out["cursors_%d" % f] = r.cursors()[:3]
            if f % 2 == 1:
                xyz, fut = r.occupancy(0.2)
                out["occ_xyz_%d" % f] = xyz
                fz = np.nonzero(fut.ravel())[0].astype(np.int32)
                out["fut_idx_%d" % f], out["fut_val_%d" % f] = fz, fut.ravel()[fz]
        path = os.path.join(ROOT, "tests", "golden", "%s_%d_%d.npz" % (name, frames, initp))
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
