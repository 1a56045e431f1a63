#!/usr/bin/env python3
"""bench.py — map updates/s of the DSP-Dynamic per-frame particle loop on B200 (BASELINE.json metric).

One "step" = DSPMap::update() + getOccupancyMapWithFutureStatus() on one frame of a deterministic synthetic
depth-cloud + pose stream.  Default workload = BASELINE.json configs[1]: DSP-Dynamic 66x66x40 voxels
@0.15 m, 24 particles/voxel, 90x60 deg FOV, 10 k-point cloud.

  python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the unmodified reference header (oracle/_ref) on host cores

JSON line keys: see DESIGN.md "Measurement".  `value` = frames/s with clouds already resident in HBM (device-resident
C-ABI entry points, CUDA events, L2 flushed between steps); `e2e` = the same frames through the host-pointer C-ABI
(dspmap_update + dspmap_get_occupancy: H2D of the cloud, host velocity estimation, D2H of the occupied-voxel list and
the V x T future grid inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "dsp-map_b200"))

import numpy as np  # noqa: E402

THRESHOLD = 0.2   # occupancy threshold used by the example app (src/map_sim_example.cpp:378)
PREROLL = 25      # untimed frames to reach the steady particle population (discard >= 20)
SWITCHES = ("DSPMAP_PDL", "DSPMAP_EST_THREAD")   # see main(): opted into by the single-map arm, not by the sharded one
SETTERS = dict(p_std=0.05, v_std=0.05, ob_std=0.1, newborn_num=20, newborn_weight=1e-4, filter_res=0.1)  # ex:522-526


def apply_setters(m):
    m.setPredictionVariance(SETTERS["p_std"], SETTERS["v_std"])
    m.setObservationStdDev(SETTERS["ob_std"])
    m.setNewBornParticleNumberofEachPoint(SETTERS["newborn_num"])
    m.setNewBornParticleWeight(SETTERS["newborn_weight"])
    m.setOriginalVoxelFilterResolution(SETTERS["filter_res"])


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nme in enumerate(names):
                if len(r) > 4 + k and r[4 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def reference_arm(args, cfg_name, cfg):
    """Times the reference's own CPU implementation (unmodified header, oracle/_ref) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from dspmap_b200.streams import make_stream
    import refmap
    if not refmap.available(cfg_name):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libdspref_%s.so not built" % cfg_name}))
        return 0
    pre = 8  # the reference needs ~8 frames to reach its steady particle population (1 s each at cfg2)
    F = pre + args.warmup + args.steps
    st = make_stream(cfg, seed=1, frames=F)
    # the build with the reference's own compiler flags where this host can run it, else the -O2 build the parity tests use
    lib_name = refmap.fast_variant(cfg_name) or cfg_name
    flags = "-O3 -ftree-vectorize -ffast-math (the reference's CMakeLists.txt:4) -mavx2 -mfma" if lib_name != cfg_name else "g++ -O2"
    r = refmap.RefMap(lib_name, seed=1, **SETTERS)
    fut = np.zeros((r.V, r.T), np.float32)
    times = []
    for f in range(F):
        s, _ = r.timed_frame(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f], THRESHOLD, fut)
        if f >= pre + args.warmup:
            times.append(s)
    total = float(np.sum(times))
    val = len(times) / total
    line = {"impl": "reference", "metric": "map_updates_per_s", "value": val, "unit": "updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg_name, cfg, "cpu"),
            "cpu_baseline": {"value": val, "unit": "updates/s", "cores": 2, "kind": "reference",
                             "sample": "%d frames of the same stream after %d untimed frames; unmodified reference header, "
                                       "%s, 1 thread + its 1 helper thread of %d host cores" % (len(times), pre + args.warmup, flags, os.cpu_count())},
            "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def workload_config(cfg_name, cfg, where):
    return {"workload": "%s: DSP-%s %dx%dx%d vox @%.2f m, %d ppv, FOV %dx%d deg, %d-pt synthetic depth cloud, update()+getOccupancyMapWithFutureStatus()"
                        % (cfg_name, "Static" if cfg["model"] == "static" else "Dynamic", cfg["nx"], cfg["ny"], cfg["nz"], cfg["res"],
                           cfg["max_ppv"], 2 * cfg["half_fov_h"], 2 * cfg["half_fov_v"], cfg["points"]),
            "horizons": cfg["future_times"], "neighbors": (2 * cfg["neighbor_n"] + 1) ** 2, "where": where}


def main_sharded(args, cfg_name, cfg, dm, make_stream):
    """--gpus N > 1: ONE map, its voxel subspaces (z slabs) sharded over the N GPUs (one process per GPU, NCCL):
    all-to-all of boundary crossers, all-gather of registered particles, all-reduce of the newborn split per frame,
    plus the reader's all-reduce of the future grid and gather of the occupied-voxel lists.  Strong scaling."""
    import torch
    import torch.distributed as dist
    from dspmap_b200.sharded import NcclComm, ShardedDSPMap, sharded_update
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, args.warmup
    PROF = min(K, 10)
    F = PREROLL + W + K + PROF + W + K
    st = make_stream(cfg, seed=1, frames=F)     # the same stream on every rank: the cloud and the pose are replicated
    M = int(st["n"][0])
    est = dm.VelocityEstimator(cfg, seed=1, filter_res=SETTERS["filter_res"])
    F1 = PREROLL + W + K + PROF
    tagged, last = [], np.zeros((0, 7), np.float32)
    for f in range(F1):
        t = est.estimate(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f])
        last = t if t is not None else last
        tagged.append(last)
    nt_max = max(max(len(t) for t in tagged), 1)
    d_pts = torch.from_numpy(st["points"][:F1]).to(dev)
    tg = np.zeros((F1, nt_max, 7), np.float32)
    for f, t in enumerate(tagged):
        tg[f, :len(t)] = t
    d_tag = torch.from_numpy(tg).to(dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sm = ShardedDSPMap(cfg, rank, world, device=local, seed=1, max_points=max(M, nt_max, 1024))
    m = sm.map
    m.set_stream(stream.cuda_stream)
    apply_setters(m)
    comm = NcclComm()
    cap_occ = 32768
    d_xyz = torch.zeros((m.V, 3), dtype=torch.float32, device=dev)
    d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    d_fut = torch.zeros((m.V, m.T), dtype=torch.float32, device=dev)
    g_xyz = torch.zeros((world, cap_occ, 3), dtype=torch.float32, device=dev)
    g_cnt = torch.zeros(world, dtype=torch.int32, device=dev)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(f, d_p, d_t, nt):
        sharded_update(sm, comm, M, d_p, st["pos"][f], st["t"][f], st["quat"][f], d_t, nt)
        m.get_occupancy_device(THRESHOLD, d_xyz.data_ptr(), m.V, d_cnt.data_ptr(), d_fut.data_ptr())
        dist.all_reduce(d_fut)                                        # future contributions land in any rank's voxels
        dist.all_gather_into_tensor(g_xyz.view(-1), d_xyz[:cap_occ].reshape(-1))   # slab lists concatenate in voxel order
        dist.all_gather_into_tensor(g_cnt, d_cnt)

    for f in range(PREROLL + W):
        step(f, d_pts[f].data_ptr(), d_tag[f].data_ptr(), len(tagged[f]))
    torch.cuda.synchronize()
    dist.barrier()
    m.synchronize()
    launches0 = m.counters()["launches_total"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    for k in range(K):
        f = PREROLL + W + k
        if flush is not None:
            flush.zero_()
        e0[k].record(stream)
        step(f, d_pts[f].data_ptr(), d_tag[f].data_ptr(), len(tagged[f]))
        e1[k].record(stream)
    torch.cuda.synchronize()
    dist.barrier()
    m.synchronize()
    dev_ms = sum(a.elapsed_time(b) for a, b in zip(e0, e1))
    launches = m.counters()["launches_total"] - launches0
    tt = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms_max = float(tt.item())
    value = K / (dev_ms_max * 1e-3)        # one map: every frame is ONE update, whatever the number of GPUs
    # per-family kernel times on this rank + frame counters summed over ranks
    m.profile_enable(True)
    agg = {}
    for k in range(PROF):
        f = PREROLL + W + K + k
        if flush is not None:
            flush.zero_()
        step(f, d_pts[f].data_ptr(), d_tag[f].data_ptr(), len(tagged[f]))
        m.synchronize()
        for kk, vv in m.counters().items():
            agg[kk] = agg.get(kk, 0) + vv
    prof = m.profile_read()
    m.profile_enable(False)
    names = sorted(agg)
    ct = torch.tensor([agg[k_] for k_ in names], dtype=torch.float64, device=dev)
    dist.all_reduce(ct)
    ctr = {k_: float(v_) / PROF for k_, v_ in zip(names, ct.tolist())}
    fam_ms = {n: (ms / PROF) for n, (ms, ln) in prof.items() if ln}
    # end to end: host cloud in, host results out on rank 0
    fut_host = torch.zeros((m.V, m.T), dtype=torch.float32).pin_memory()
    xyz_host = torch.zeros((world, cap_occ, 3), dtype=torch.float32).pin_memory()
    e2e_t, h2d, d2h = [], 0, 0
    for k in range(W + K):
        f = F1 + k
        pts, pos, t, q = st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]
        if flush is not None:
            flush.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        dp_ = torch.from_numpy(pts).to(dev, non_blocking=True)
        tc = est.estimate(pts, pos, t, q)
        last = tc if tc is not None else last
        dt_ = torch.from_numpy(last if len(last) else np.zeros((1, 7), np.float32)).to(dev, non_blocking=True)
        step(f, dp_.data_ptr(), dt_.data_ptr(), len(last))
        if rank == 0:
            fut_host.copy_(d_fut, non_blocking=True)
            xyz_host.copy_(g_xyz, non_blocking=True)
            n_occ = int(g_cnt.sum().item())
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if k >= W:
            e2e_t.append(t1 - t0)
            h2d += pts.nbytes + last.nbytes
            d2h += (fut_host.numel() + xyz_host.numel()) * 4 + 4 if rank == 0 else 0
    tt = torch.tensor([sum(e2e_t)], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_val = len(e2e_t) / float(tt.item())
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        peak, peak_src = measured_peak()
        b_frame = dm.bytes_per_update(ctr, m.V, m.T, M)
        top = max(fam_ms, key=fam_ms.get)
        line = {
            "metric": "map_updates_per_s", "value": value, "unit": "updates/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": dict(workload_config(cfg_name, cfg, "hbm-resident"), parallelism="voxel z-slabs x%d (one map)" % world,
                           collectives_per_update=["all_to_all(boundary crossers)", "all_gather(registered particles)",
                                                   "all_reduce(newborn split)", "all_reduce(future grid)", "all_gather(occupied lists)"],
                           l2_flush_between_steps=flush is not None, preroll_frames=PREROLL),
            "e2e": {"value": e2e_val, "unit": "updates/s", "h2d_bytes_per_step": h2d // max(len(e2e_t), 1),
                    "d2h_bytes_per_step": d2h // max(len(e2e_t), 1), "ms_per_step": 1e3 * float(np.mean(e2e_t))},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": top, "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
                         "peak_source": peak_src, "kernel_ms": fam_ms[top], "note": "rank 0's share; see the N=1 line for the per-kernel roofline"},
            "roofline_frame": {"bound": "hbm", "algorithmic_bytes_per_update": b_frame,
                               "achieved": b_frame / (dev_ms_max / K * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                               "frac": b_frame / (dev_ms_max / K * 1e-3) / 1e9 / (peak * world)},
            "kernel_ms_per_update_rank0": {k_: round(v_, 5) for k_, v_ in sorted(fam_ms.items(), key=lambda kv: -kv[1])},
            "counters_per_update": {k_: round(v_, 1) for k_, v_ in ctr.items() if k_ not in ("launches_total", "launches_frame")},
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--replicas", action="store_true", help="N > 1: N independent maps instead of one sharded map")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    import dspmap_b200 as dm
    from dspmap_b200.streams import make_stream
    cfg_name = args.config
    cfg = dm.CONFIGS[cfg_name]
    if args.impl == "reference":
        return reference_arm(args, cfg_name, cfg)

    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not args.replicas:
        return main_sharded(args, cfg_name, cfg, dm, make_stream)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # Library switches this arm opts into (read by dspmap_create; set them to 0 to measure the library defaults).  Both were
    # A/B-measured on B200 on exactly this workload and are bit-identical to the default path (profiles/r01_ab_switches.jsonl,
    # tests/test_gpu_parity.py): programmatic dependent launch of the frame's kernels, and the velocity estimation on the
    # library's helper thread (the reference runs it on a std::thread as well).
    for k_ in SWITCHES:
        os.environ.setdefault(k_, "1")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    K, W = args.steps, args.warmup
    F = PREROLL + W + K          # device-resident pass
    PROF = min(K, 20)            # frames of the per-kernel profiling pass
    F2 = F + PROF + W + K        # + profiling pass + end-to-end pass on the following frames
    F3 = F2 + W + K              # + pipelined end-to-end pass
    st = make_stream(cfg, seed=1 + rank, frames=F3)
    M = int(st["n"][0])

    # newborn inputs for the device-resident pass: the library's own host velocity estimator, pre-computed
    est = dm.VelocityEstimator(cfg, seed=1, filter_res=SETTERS["filter_res"])
    tagged, last = [], np.zeros((0, 7), np.float32)
    for f in range(F):
        t = est.estimate(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f])
        last = t if t is not None else last
        tagged.append(last)
    nt_max = max(len(t) for t in tagged)
    d_pts = torch.from_numpy(st["points"][:F]).to(dev)
    tg = np.zeros((F, max(nt_max, 1), 7), np.float32)
    for f, t in enumerate(tagged):
        tg[f, :len(t)] = t
    d_tag = torch.from_numpy(tg).to(dev)

    m = dm.DSPMap(cfg, seed=1, device=local, max_points=max(M, nt_max, 1024))
    apply_setters(m)
    stream = torch.cuda.Stream(device=dev)  # a real (non-default) stream shared by torch's events and the library's kernels
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    m.set_stream(stream.cuda_stream)
    d_xyz = torch.empty((m.V, 3), dtype=torch.float32, device=dev)
    d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    d_fut = torch.empty((m.V, m.T), dtype=torch.float32, device=dev)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_device(f):
        m.update_device(M, d_pts[f].data_ptr(), st["pos"][f], st["t"][f], st["quat"][f], d_tag[f].data_ptr(), len(tagged[f]))
        m.get_occupancy_device(THRESHOLD, d_xyz.data_ptr(), m.V, d_cnt.data_ptr(), d_fut.data_ptr())

    for f in range(PREROLL + W):
        step_device(f)
    m.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = m.counters()["launches_total"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ncu = os.environ.get("DSPMAP_NCU") == "1"  # under `ncu --profile-from-start off` only the timed region is captured
    if ncu:
        torch.cuda.profiler.start()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ctr_sum = None
    for k in range(K):
        if flush is not None:
            flush.zero_()
        e0[k].record(stream)
        step_device(PREROLL + W + k)
        e1[k].record(stream)
    torch.cuda.synchronize()
    m.synchronize()
    if ncu:
        torch.cuda.profiler.stop()
    if world > 1:
        dist.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in zip(e0, e1))
    launches = m.counters()["launches_total"] - launches0
    tt = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms_max = float(tt.item())
    value = world * K / (dev_ms_max * 1e-3)

    # per-family kernel times (CUDA events on the launching stream) over the same kind of frames, and the frame counters
    # that define the algorithmic bytes
    m.profile_enable(True)
    P = PROF
    agg = {}
    for k in range(P):
        f = PREROLL + W + K - P + k  # replaying already-seen inputs is fine: kernel work depends on the map state
        if flush is not None:
            flush.zero_()
        m.update(M, 3, st["points"][F + k], *map(float, st["pos"][F + k]), float(st["t"][F + k]), *map(float, st["quat"][F + k]),
                 tagged=est.estimate(st["points"][F + k], st["pos"][F + k], st["t"][F + k], st["quat"][F + k]))
        c = m.counters()
        for kk, vv in c.items():
            agg[kk] = agg.get(kk, 0) + vv
        m.get_occupancy_device(THRESHOLD, d_xyz.data_ptr(), m.V, d_cnt.data_ptr(), d_fut.data_ptr())
    m.synchronize()
    prof = m.profile_read()
    kprof = m.profile_read_kernels()
    m.profile_enable(False)
    ctr = {kk: vv / P for kk, vv in agg.items()}
    fam_ms = {n: (ms / P) for n, (ms, ln) in prof.items() if ln}
    fam_launches = {n: ln / P for n, (ms, ln) in prof.items() if ln}

    # end-to-end: host buffers in, host buffers out, through the reference-facing calls
    fut_host = np.zeros((m.V, m.T), np.float32)
    m.pin_host_buffer(fut_host)  # what the drop-in header does with the application's static future_status array
    e2e_t = []
    e2e_upd = []
    h2d = d2h = 0
    for k in range(W + K):
        f = F + P + k
        pts, pos, t, q = st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]
        if flush is not None:
            flush.zero_()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = m.update(M, 3, pts, float(pos[0]), float(pos[1]), float(pos[2]), float(t), float(q[0]), float(q[1]), float(q[2]), float(q[3]))
        tm = time.perf_counter()
        n_occ, xyz, _ = m.getOccupancyMapWithFutureStatus(THRESHOLD, fut_host)
        t1 = time.perf_counter()
        if k >= W and rc == 1:
            e2e_t.append(t1 - t0)
            e2e_upd.append(tm - t0)
            h2d += pts.nbytes + 28 * len(m.getKMClusterResult())
            d2h += 4 + 12 * n_occ + fut_host.nbytes + 160
    tt = torch.tensor([sum(e2e_t)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_val = world * len(e2e_t) / float(tt.item())

    # pipelined end-to-end: the same host-pointer update(), results through
    # dspmap_get_occupancy_async / dspmap_wait_occupancy — frame k-1's device-to-host copies overlap update(k).  The L2
    # flush is enqueued INSIDE the timed region here (a synchronising flush would serialise the pipeline).
    ticket, touched, t0, n_pipe = None, 0.0, None, 0
    for k in range(W + K):
        f = F2 + k
        pts, pos, t, q = st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]
        if k == W:
            if ticket is not None:
                m.wait_occupancy(ticket)
                ticket = None
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        if flush is not None:
            flush.zero_()
        rc = m.update(M, 3, pts, float(pos[0]), float(pos[1]), float(pos[2]), float(t), float(q[0]), float(q[1]), float(q[2]), float(q[3]))
        prev, ticket = ticket, m.get_occupancy_async(THRESHOLD, True)
        if prev is not None:
            n_occ, xyz, fut = m.wait_occupancy(prev)
            touched += float(fut[0, 0]) + n_occ
        n_pipe += 1 if (k >= W and rc == 1) else 0
    n_occ, xyz, fut = m.wait_occupancy(ticket)
    touched += float(fut[0, 0]) + n_occ
    t1 = time.perf_counter()
    tt = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    pipe_val = world * n_pipe / float(tt.item())
    clocks = sampler.stop() if rank == 0 else None

    # application-side preprocessing on the GPU, measured beside the headline: a raw 640 x 480 depth
    # cloud (16-byte points, 8 % invalid) through dspmap_prefilter_run_device (CUDA events) and dspmap_prefilter_run (host)
    prefilter = None
    if rank == 0:
        from dspmap_b200.streams import make_depth_cloud
        raw = make_depth_cloud(640, 480, seed=1, stride=4)
        lo = (-cfg["nx"] * cfg["res"] / 2, -cfg["ny"] * cfg["res"] / 2, -cfg["nz"] * cfg["res"] / 2)
        hi = tuple(-x for x in lo)
        pf = dm.Prefilter(max_raw_points=len(raw), max_stride=4, max_out_points=5000)
        pf.set_stream(stream.cuda_stream)
        d_raw = torch.from_numpy(raw).to(dev)
        d_out = torch.zeros((5000, 3), dtype=torch.float32, device=dev)
        d_n = torch.zeros(1, dtype=torch.int32, device=dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for a_, b_ in [(None, None)] * 3 + ev:
            if flush is not None:
                flush.zero_()
            if a_ is not None:
                a_.record(stream)
            pf.run_device(len(raw), 4, d_raw.data_ptr(), 0.1, lo, hi, d_out.data_ptr(), 5000, d_n.data_ptr())
            if b_ is not None:
                b_.record(stream)
        torch.cuda.synchronize()
        dev_ms_pf = float(np.median([a_.elapsed_time(b_) for a_, b_ in ev]))
        raw_pin = torch.from_numpy(raw).pin_memory().numpy()
        host_ms = []
        for k in range(13):
            t0 = time.perf_counter()
            out_pf = pf.run(raw_pin, 0.1, lo, hi)
            host_ms.append(1e3 * (time.perf_counter() - t0))
        n_fin = int(np.isfinite(raw[:, 0]).sum())
        pf_bytes = 2 * raw.nbytes + 12 * len(out_pf)   # the raw cloud is read twice (extent, accumulation); centroids written once
        prefilter = {"workload": "640x480 raw depth cloud, 16 B points, %d finite -> %d points (VoxelGrid 0.1 m + axis swap + crop)" % (n_fin, len(out_pf)),
                     "device_ms": dev_ms_pf, "host_api_ms": float(np.median(host_ms[3:])), "gpu_launches": 7,
                     "algorithmic_bytes": pf_bytes, "achieved_gbs": pf_bytes / (dev_ms_pf * 1e-3) / 1e9,
                     "h2d_bytes": int(raw.nbytes), "d2h_bytes": 5000 * 12 + 4}
        pf.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    T, V = m.T, m.V
    b_frame = dm.bytes_per_update(ctr, V, T, M)
    # dominant kernel = the family with the largest device time; its algorithmic bytes (DESIGN.md "Kernels")
    top = max(fam_ms, key=fam_ms.get)
    fam_bytes = {
        "ck_pass": 16 * ctr["n_fov"] + 20 * M,             # read px,py,pz,w per registered particle; read point, write C_z
        "weight_pass": 20 * ctr["n_fov"] + 20 * M,         # read px,py,pz,w + write w; read point + C_z
        "predict": 64 * ctr["n_in"],
        "newborn": 32 * ctr["n_born"] + 28 * M,
        "resample_future": 32 * ctr["n_pre"] + 32 * ctr["n_out"] + 4 * T * ctr["n_old"] + V * (16 + 16),
        "reader": V * (4 + 12 * T),
        "pyramid_lists": 28 * ctr["n_fov"],
        "arrive": 64 * ctr["n_moved"],
        "obs_bin": 32 * M,
        "enumerate": 32 * V + 4 * ctr["n_in"],
    }
    # dominant KERNEL = the launch site with the largest event-timed device time per update; its algorithmic bytes per
    # launch (DESIGN.md "Kernels": what the reference's arithmetic needs it to read and write, not what this
    # implementation stages in between) over its average launch duration
    k_ms = {n: ms / max(ln, 1) for n, (ms, ln) in kprof.items() if ln}           # average launch duration
    k_per_update = {n: ms / P for n, (ms, ln) in kprof.items() if ln}
    top = max(k_per_update, key=k_per_update.get)
    kernel_bytes = {
        "k_weight2": 20 * ctr["n_fov"] + 20 * M,           # read px,py,pz,w + write w; read point + C_z
        "k_weight2w": 20 * ctr["n_fov"] + 20 * M,
        "k_pair_eval": 16 * ctr["n_fov"] + 16 * M,         # read px,py,pz,w per registered particle; read the points
        "k_cz_wide": 4 * ctr["n_fov"] + 8 * M,             # read P_d*w per particle; write C_z and 1/C_z per point
        "k_cz_narrow": 4 * ctr["n_fov"] + 8 * M,
        "k_nb_place": 32 * ctr["n_born"],
        "k_resample": 32 * ctr["n_pre"] + 32 * ctr["n_out"] + 4 * T * ctr["n_old"] + 16 * V,
        "k_predict": 64 * ctr["n_in"],
        "k_pyr_sort": 28 * ctr["n_fov"],
    }
    top_ms = k_ms[top]
    achieved = kernel_bytes.get(top, fam_bytes.get(top, b_frame)) / (top_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(top)
    line = {
        "metric": "map_updates_per_s", "value": value, "unit": "updates/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": dict(workload_config(cfg_name, cfg, "hbm-resident"), parallelism="replicas x%d" % world if world > 1 else "single",
                       l2_flush_between_steps=flush is not None, preroll_frames=PREROLL,
                       library_switches={k_: os.environ.get(k_, "0") for k_ in SWITCHES}),
        "e2e": {"value": e2e_val, "unit": "updates/s", "h2d_bytes_per_step": h2d // max(len(e2e_t), 1),
                "d2h_bytes_per_step": d2h // max(len(e2e_t), 1), "ms_per_step": 1e3 * float(np.mean(e2e_t)),
                "update_ms": 1e3 * float(np.mean(e2e_upd)), "reader_ms": 1e3 * float(np.mean(e2e_t) - np.mean(e2e_upd))},
        "e2e_pipelined": {"value": pipe_val, "unit": "updates/s", "ms_per_step": 1e3 * float(tt.item()) / max(n_pipe, 1),
                          "api": "dspmap_update + dspmap_get_occupancy_async / dspmap_wait_occupancy (host buffers; the copies of "
                                 "frame k-1 overlap update(k)); L2 flush enqueued inside the timed region"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": kernel_bytes.get(top),
                     "kernel_ms": top_ms, "launches_per_update": kprof[top][1] / P,
                     "note": "instruction-issue / latency bound: see DESIGN.md section 5 and profiles/r01_top_kernels.md"},
        "roofline_frame": {"bound": "hbm", "algorithmic_bytes_per_update": b_frame, "achieved": b_frame / (dev_ms_max / K * 1e-3) / 1e9,
                           "peak": peak, "unit": "GB/s", "frac": b_frame / (dev_ms_max / K * 1e-3) / 1e9 / peak},
        "kernel_ms_per_update": {k_: round(v_, 5) for k_, v_ in sorted(k_per_update.items(), key=lambda kv: -kv[1])},
        "family_ms_per_update": {k_: round(v_, 5) for k_, v_ in sorted(fam_ms.items(), key=lambda kv: -kv[1])},
        "counters_per_update": {k_: round(v_, 1) for k_, v_ in ctr.items() if k_ not in ("launches_total",)},
        "verified_fast_division": dict(zip(("voxel_size", "sigma"), m.fast_paths())),
        "prefilter": prefilter,
    }
    if world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import refmap
        import prefilter_oracle
        t0 = time.perf_counter()
        prefilter_oracle.preprocess(raw, 0.1, lo, hi, 5000)
        line["prefilter"]["cpu_port_ms"] = 1e3 * (time.perf_counter() - t0)   # numpy restatement, 1 thread (PCL itself is absent)
        if refmap.available(cfg_name):
            pre, n_t = 8, 5
            r = refmap.RefMap(cfg_name, seed=1, **SETTERS)
            futr = np.zeros((r.V, r.T), np.float32)
            ts = []
            for f in range(pre + n_t):
                s, _ = r.timed_frame(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f], THRESHOLD, futr)
                if f >= pre:
                    ts.append(s)
            line["cpu_baseline"] = {"value": len(ts) / float(np.sum(ts)), "unit": "updates/s", "cores": 2, "kind": "reference",
                                    "sample": "frames %d..%d of the same stream; unmodified reference header (oracle/_ref), g++ -O2, "
                                              "1 thread + its 1 helper thread of %d host cores" % (pre, pre + n_t - 1, os.cpu_count())}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
