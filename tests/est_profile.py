"""Where the time of the device-side estimation front end goes (GPU box): per-kernel event times and the wall clock of
update() over a long cfg2 stream.  python tests/est_profile.py [config] [frames]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dsp-map_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
os.environ.setdefault("DSPMAP_TIMELINE", "1")
os.environ.setdefault("DSPMAP_EST_GPU", "1")
import dspmap_b200 as dm
from common import make_stream, gpu_map

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 160
cfg = dm.CONFIGS[name]
st = make_stream(cfg, seed=1, frames=frames)
m = gpu_map(name, seed=1, max_points=cfg["points"])
fut = np.zeros((m.V, m.T), np.float32)
m.pin_host_buffer(fut)
for f in range(frames if len(sys.argv) < 4 else 0):
    prof = f % 20 == 19
    if prof:
        m.profile_enable(True)
    pts, pos, t, q = st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]
    t0 = time.perf_counter()
    m.update(len(pts), 3, pts, float(pos[0]), float(pos[1]), float(pos[2]), float(t), float(q[0]), float(q[1]), float(q[2]), float(q[3]))
    t1 = time.perf_counter()
    m.getOccupancyMapWithFutureStatus(0.2, fut)
    t2 = time.perf_counter()
    if prof:
        m.synchronize()
        k = m.profile_read_kernels()
        m.profile_enable(False)
        est = {a: round(1e3 * b[0] / max(b[1], 1), 1) for a, b in k.items() if a.startswith("k_est")}
        print("frame %3d: update %.3f ms reader %.3f ms (profiled, synchronous) est kernels us: %s  stats %s" % (f, 1e3 * (t1 - t0), 1e3 * (t2 - t1), est, m.estimator_stats()[1]), flush=True)
    elif f % 20 == 10:
        print("frame %3d: update %.3f ms reader %.3f ms  timeline %s" % (f, 1e3 * (t1 - t0), 1e3 * (t2 - t1), m.timeline()), flush=True)
m.close()

# the same frames device-resident (explicit newborn input, early newborn kernels beside the observation passes): timeline
import torch
dev = torch.device("cuda", 0)
m = gpu_map(name, seed=1, max_points=cfg["points"])
est = dm.VelocityEstimator(cfg, seed=1, filter_res=0.1)
d_xyz = torch.empty((m.V, 3), dtype=torch.float32, device=dev)
d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
d_fut = torch.empty((m.V, m.T), dtype=torch.float32, device=dev)
last = np.zeros((0, 7), np.float32)
for f in range(min(frames, 100)):
    pts, pos, t, q = st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]
    tg = est.estimate(pts, pos, t, q)
    last = tg if tg is not None else last
    d_pts = torch.from_numpy(pts).to(dev)
    d_tag = torch.from_numpy(last).to(dev)
    torch.cuda.synchronize()
    m.update_device(len(pts), d_pts.data_ptr(), pos, t, q, d_tag.data_ptr(), len(last))
    m.get_occupancy_device(0.2, d_xyz.data_ptr(), m.V, d_cnt.data_ptr(), d_fut.data_ptr())
    if f % 20 == 10:
        print("device-resident frame %3d: timeline %s" % (f, m.timeline()), flush=True)
m.close()
