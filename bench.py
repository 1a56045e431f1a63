#!/usr/bin/env python3
"""bench.py — map updates/s of the DSP-Dynamic per-frame particle loop on B200 (BASELINE.json metric).

One "step" = DSPMap::update() + getOccupancyMapWithFutureStatus() on one frame of a deterministic synthetic
depth-cloud + pose stream.  Headline workload = BASELINE.json configs[1]: DSP-Dynamic 66x66x40 voxels
@0.15 m, 24 particles/voxel, 90x60 deg FOV, 10 k-point cloud; configs[2] (5x5 neighbourhoods, 20 k points) and configs[4]
(132x132x80 voxels, 36 ppv, 30 k points) are measured the same way, more briefly, and reported under "other_configs".

  python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path (N > 1: torchrun, one rank per GPU)
  python bench.py --impl reference --steps K --warmup W    # the unmodified reference header (oracle/_ref) on host cores

N = 1: one map on one GPU.  N > 1: ONE map whose voxel subspaces (z slabs) are sharded over the N GPUs, orchestrated by the
library itself (dspmap_shard_update / dspmap_shard_get_occupancy: C++ host code, NCCL collectives), strong scaling.
Both arms time the SAME frames of the SAME stream: PREROLL untimed frames, W warm-ups, K timed frames.

JSON line keys: see DESIGN.md "Measurement".  `value` = frames/s with clouds already resident in HBM (device-resident
C-ABI entry points, CUDA events on the launching stream, L2 flushed between steps); `e2e` = the same frames through the
host-pointer C-ABI (dspmap_update + dspmap_get_occupancy: H2D of the cloud, host velocity estimation, D2H of the
occupied-voxel list and the V x T future grid inside the timed region).
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


def workload_config(cfg_name, cfg):
    """The same dictionary in both arms (the driver compares them)."""
    return {"workload": "%s: DSP-%s %dx%dx%d vox @%.2f m, %d ppv, FOV %dx%d deg, %d-pt synthetic depth cloud, update()+getOccupancyMapWithFutureStatus()"
                        % (cfg_name, "Static" if cfg["model"] == "static" else "Dynamic", cfg["nx"], cfg["ny"], cfg["nz"], cfg["res"],
                           cfg["max_ppv"], 2 * cfg["half_fov_h"], 2 * cfg["half_fov_v"], cfg["points"]),
            "horizons": cfg["future_times"], "neighbors": (2 * cfg["neighbor_n"] + 1) ** 2,
            "timed_frames": "frames %d+W .. %d+W+K-1 of the seed-1 stream (after %d untimed frames and W warm-ups)" % (PREROLL, PREROLL, PREROLL)}


def reference_arm(args, cfg_name, cfg):
    """Times the reference's own CPU implementation (unmodified header, oracle/_ref) on the host cores, on the frames the
    CUDA arm times."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from dspmap_b200.streams import make_stream
    import refmap
    if not refmap.available(cfg_name):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libdspref_%s.so not built" % cfg_name}))
        return 0
    F = PREROLL + args.warmup + args.steps
    st = make_stream(cfg, seed=1, frames=F)
    # the build with the reference's own compiler flags where this host can run it, else the -O2 build the parity tests use
    lib_name = refmap.fast_variant(cfg_name) or cfg_name
    flags = "-O3 -ftree-vectorize -ffast-math (the reference's CMakeLists.txt:4) -mavx2 -mfma" if lib_name != cfg_name else "g++ -O2"
    r = refmap.RefMap(lib_name, seed=1, **SETTERS)
    fut = np.zeros((r.V, r.T), np.float32)
    times = []
    for f in range(F):
        s, _ = r.timed_frame(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f], THRESHOLD, fut)
        if f >= PREROLL + args.warmup:
            times.append(s)
    total = float(np.sum(times))
    val = len(times) / total
    line = {"impl": "reference", "metric": "map_updates_per_s", "value": val, "unit": "updates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg_name, cfg),
            "cpu_baseline": {"value": val, "unit": "updates/s", "cores": 2, "kind": "reference", "frames": len(times),
                             "sample": "frames %d..%d of the same stream (the frames the CUDA arm times); unmodified reference header, "
                                       "%s, 1 thread + its 1 helper thread of %d host cores" % (PREROLL + args.warmup, F - 1, flags, os.cpu_count())},
            "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# one map on one GPU
# ---------------------------------------------------------------------------------------------------------------------
def precompute_tagged(dm, cfg, st, F):
    """Newborn inputs for the device-resident pass: the library's own host velocity estimator, pre-computed."""
    est = dm.VelocityEstimator(cfg, seed=1, filter_res=SETTERS["filter_res"])
    tagged, last = [], np.zeros((0, 7), np.float32)
    for f in range(F):
        t = est.estimate(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f])
        last = t if t is not None else last
        tagged.append(last)
    nt_max = max(max(len(t) for t in tagged), 1)
    tg = np.zeros((F, nt_max, 7), np.float32)
    for f, t in enumerate(tagged):
        tg[f, :len(t)] = t
    return est, tagged, nt_max, tg


def run_single(dm, make_stream, torch, cfg_name, K, W, dev, stream, flush, detailed, ncu=False):
    """Device-resident pass (value), per-kernel profile, end-to-end pass.  detailed=False: value + e2e only (other_configs)."""
    cfg = dm.CONFIGS[cfg_name]
    F = PREROLL + W + K
    PROF = min(K, 20) if detailed else 0
    st = make_stream(cfg, seed=1, frames=F + PROF)
    M = int(st["n"][0])
    est, tagged, nt_max, tg = precompute_tagged(dm, cfg, st, F)
    d_pts = torch.from_numpy(st["points"][:F]).to(dev)
    d_tag = torch.from_numpy(tg).to(dev)
    m = dm.DSPMap(cfg, seed=1, device=dev.index, max_points=max(M, nt_max, 1024))
    apply_setters(m)
    m.set_stream(stream.cuda_stream)
    d_xyz = torch.empty((m.V, 3), dtype=torch.float32, device=dev)
    d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    d_fut = torch.empty((m.V, m.T), dtype=torch.float32, device=dev)

    def step_device(f):
        m.update_device(M, d_pts[f].data_ptr(), st["pos"][f], st["t"][f], st["quat"][f], d_tag[f].data_ptr(), len(tagged[f]))
        m.get_occupancy_device(THRESHOLD, d_xyz.data_ptr(), m.V, d_cnt.data_ptr(), d_fut.data_ptr())

    for f in range(PREROLL + W):
        step_device(f)
    m.synchronize()
    torch.cuda.synchronize()
    launches0 = m.counters()["launches_total"]
    if ncu:  # under `ncu --profile-from-start off` only the timed region is captured
        torch.cuda.profiler.start()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
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
    dev_list = [a.elapsed_time(b) for a, b in zip(e0, e1)]
    dev_ms = sum(dev_list)
    out = {"cfg": cfg, "M": M, "V": m.V, "T": m.T, "dev_ms": dev_ms, "dev_list": dev_list, "launches": m.counters()["launches_total"] - launches0, "st": st}

    if detailed:  # per-kernel times (CUDA events around every launch site) and the counters that define the algorithmic bytes
        m.profile_enable(True)
        agg = {}
        for k in range(PROF):
            if flush is not None:
                flush.zero_()
            m.update(M, 3, st["points"][F + k], *map(float, st["pos"][F + k]), float(st["t"][F + k]), *map(float, st["quat"][F + k]),
                     tagged=est.estimate(st["points"][F + k], st["pos"][F + k], st["t"][F + k], st["quat"][F + k]))
            for kk, vv in m.counters().items():
                agg[kk] = agg.get(kk, 0) + vv
            m.get_occupancy_device(THRESHOLD, d_xyz.data_ptr(), m.V, d_cnt.data_ptr(), d_fut.data_ptr())
        m.synchronize()
        out["prof"], out["kprof"], out["P"] = m.profile_read(), m.profile_read_kernels(), PROF
        m.profile_enable(False)
        out["ctr"] = {kk: vv / PROF for kk, vv in agg.items()}

    m.close()

    def fresh_map():  # a new map taken through the same untimed frames: every pass times the SAME frames PREROLL+W .. PREROLL+W+K-1
        g = dm.DSPMap(cfg, seed=1, device=dev.index, max_points=max(M, nt_max, 1024))
        apply_setters(g)
        g.set_stream(stream.cuda_stream)
        buf = np.zeros((g.V, g.T), np.float32)
        g.pin_host_buffer(buf)  # what the drop-in header does with the application's static future_status array
        for f in range(PREROLL):
            pos, q = st["pos"][f], st["quat"][f]
            g.update(M, 3, st["points"][f], float(pos[0]), float(pos[1]), float(pos[2]), float(st["t"][f]), float(q[0]), float(q[1]), float(q[2]), float(q[3]))
            g.getOccupancyMapWithFutureStatus(THRESHOLD, buf)
        return g, buf

    # end-to-end: host buffers in, host buffers out, through the reference-facing calls
    m, fut_host = fresh_map()
    e2e_t, e2e_upd, h2d, d2h = [], [], 0, 0
    for k in range(W + K):
        f = PREROLL + k
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
            up, down = m.last_update_bytes()
            h2d += up
            d2h += m.last_future_bytes() + 160 + down
    out.update(e2e_t=e2e_t, e2e_upd=e2e_upd, h2d=h2d, d2h=d2h)

    if detailed:
        # pipelined end-to-end: the same host-pointer update(), results through
        # dspmap_get_occupancy_async / dspmap_wait_occupancy — frame k-1's device-to-host copies overlap update(k).  The L2
        # flush is enqueued INSIDE the timed region here (a synchronising flush would serialise the pipeline).
        m.close()
        m, fut_host = fresh_map()
        ticket, touched, t0, n_pipe = None, 0.0, None, 0
        for k in range(W + K):
            f = PREROLL + k
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
        t1 = time.perf_counter()
        out.update(pipe_s=t1 - t0, n_pipe=n_pipe, fast_paths=m.fast_paths())
    m.close()
    return out


def spread(xs, scale=1.0):
    """[min, median, max] of the per-step times: a few slow frames (a saturated voxel, a burst of boundary crossers) show up here."""
    xs = np.asarray(xs, np.float64) * scale
    return [round(float(np.min(xs)), 4), round(float(np.median(xs)), 4), round(float(np.max(xs)), 4)] if len(xs) else None


def brief_record(r, K):
    return {"value": K / (r["dev_ms"] * 1e-3), "unit": "updates/s", "ms_per_step": r["dev_ms"] / K, "steps": K,
            "ms_per_step_min_med_max": spread(r["dev_list"]), "e2e_ms_min_med_max": spread(r["e2e_t"], 1e3),
            "e2e": {"value": len(r["e2e_t"]) / float(np.sum(r["e2e_t"])), "unit": "updates/s", "ms_per_step": 1e3 * float(np.mean(r["e2e_t"])),
                    "h2d_bytes_per_step": r["h2d"] // max(len(r["e2e_t"]), 1), "d2h_bytes_per_step": r["d2h"] // max(len(r["e2e_t"]), 1)},
            "gpu_launches_per_step": r["launches"] / K}


def prefilter_record(dm, torch, cfg, dev, stream, flush):
    """Application-side preprocessing on the GPU, measured beside the headline."""
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
    rec = {"workload": "640x480 raw depth cloud, 16 B points, %d finite -> %d points (VoxelGrid 0.1 m + axis swap + crop)" % (n_fin, len(out_pf)),
           "device_ms": dev_ms_pf, "host_api_ms": float(np.median(host_ms[3:])), "gpu_launches": 7,
           "algorithmic_bytes": pf_bytes, "achieved_gbs": pf_bytes / (dev_ms_pf * 1e-3) / 1e9,
           "h2d_bytes": int(raw.nbytes), "d2h_bytes": 5000 * 12 + 4}
    pf.close()
    return rec, raw, lo, hi


def main_single(args, cfg_name, dm, make_stream):
    import torch
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)  # a real (non-default) stream shared by torch's events and the library's kernels
    torch.cuda.set_stream(stream)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    K, W = args.steps, args.warmup
    cfg = dm.CONFIGS[cfg_name]
    sampler = ClockSampler(0)
    sampler.start()
    r = run_single(dm, make_stream, torch, cfg_name, K, W, dev, stream, flush, detailed=True, ncu=os.environ.get("DSPMAP_NCU") == "1")
    clocks = sampler.stop()
    others = {}
    if not args.only_headline:
        Kb = max(3, min(K, 10))
        for other in ("cfg3", "cfg5"):
            if other != cfg_name:
                others[other] = dict(brief_record(run_single(dm, make_stream, torch, other, Kb, W, dev, stream, flush, detailed=False), Kb),
                                     config=workload_config(other, dm.CONFIGS[other]))
    prefilter, raw, lo, hi = prefilter_record(dm, torch, cfg, dev, stream, flush)

    peak, peak_src = measured_peak()
    T, V, M, P, ctr = r["T"], r["V"], r["M"], r["P"], r["ctr"]
    b_frame = dm.bytes_per_update(ctr, V, T, M)
    fam_ms = {n: (ms / P) for n, (ms, ln) in r["prof"].items() if ln}
    # dominant KERNEL = the launch site with the largest event-timed device time per update; its algorithmic bytes per
    # launch (DESIGN.md "Kernels": what the reference's arithmetic needs it to read and write, not what this
    # implementation stages in between) over its average launch duration
    kprof = r["kprof"]
    k_ms = {n: ms / max(ln, 1) for n, (ms, ln) in kprof.items() if ln}
    k_per_update = {n: ms / P for n, (ms, ln) in kprof.items() if ln}
    top = max(k_per_update, key=k_per_update.get)
    kernel_bytes = {
        "k_weight2": 20 * ctr["n_fov"] + 20 * M,           # read px,py,pz,w + write w; read point + C_z
        "k_weight2w": 20 * ctr["n_fov"] + 20 * M,
        "k_pair_eval": 16 * ctr["n_fov"] + 16 * M,         # read px,py,pz,w per registered particle; read the points
        "k_cz_wide": 4 * ctr["n_fov"] + 8 * M,             # read P_d*w per particle; write C_z and 1/C_z per point
        "k_nb_place": 32 * ctr["n_born"],
        "k_resample": 32 * ctr["n_pre"] + 32 * ctr["n_out"] + 4 * T * ctr["n_old"] + 16 * V,
        "k_predict": 64 * ctr["n_in"],
        "k_pyr_sort": 28 * ctr["n_fov"],
    }
    top_ms = k_ms[top]
    achieved = kernel_bytes.get(top, b_frame) / (top_ms * 1e-3) / 1e9
    traffic = frame_traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic, frame_traffic = tj.get(top), tj.get("frame")
    dev_ms, e2e_t, e2e_upd = r["dev_ms"], r["e2e_t"], r["e2e_upd"]
    line = {
        "metric": "map_updates_per_s", "value": K / (dev_ms * 1e-3), "unit": "updates/s", "n_gpus": 1, "steps": K, "warmup": W,
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(cfg_name, cfg),
        "run": {"where": "hbm-resident", "parallelism": "single", "l2_flush_between_steps": flush is not None, "preroll_frames": PREROLL,
                "library_defaults": "PDL, helper-thread velocity estimation, asynchronous update, sparse future copy-out"},
        "e2e": {"value": len(e2e_t) / float(np.sum(e2e_t)), "unit": "updates/s", "h2d_bytes_per_step": r["h2d"] // max(len(e2e_t), 1),
                "d2h_bytes_per_step": r["d2h"] // max(len(e2e_t), 1), "ms_per_step": 1e3 * float(np.mean(e2e_t)),
                "update_ms": 1e3 * float(np.mean(e2e_upd)), "reader_ms": 1e3 * float(np.mean(e2e_t) - np.mean(e2e_upd))},
        "e2e_pipelined": {"value": r["n_pipe"] / r["pipe_s"], "unit": "updates/s", "ms_per_step": 1e3 * r["pipe_s"] / max(r["n_pipe"], 1),
                          "api": "dspmap_update + dspmap_get_occupancy_async / dspmap_wait_occupancy (host buffers; the copies of "
                                 "frame k-1 overlap update(k)); L2 flush enqueued inside the timed region"},
        "gpu_launches": int(r["launches"]), "launches_per_update": r["launches"] / K,
        "ms_per_step_min_med_max": spread(r["dev_list"]), "e2e_ms_min_med_max": spread(e2e_t, 1e3),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": kernel_bytes.get(top),
                     "kernel_ms": top_ms, "launches_per_update": kprof[top][1] / P,
                     "note": "instruction-issue / latency bound: see DESIGN.md section 5 and profiles/r02_top_kernels.md"},
        "roofline_frame": {"bound": "hbm", "algorithmic_bytes_per_update": b_frame, "achieved": b_frame / (dev_ms / K * 1e-3) / 1e9,
                           "peak": peak, "unit": "GB/s", "frac": b_frame / (dev_ms / K * 1e-3) / 1e9 / peak, "traffic": frame_traffic},
        "kernel_ms_per_update": {k_: round(v_, 5) for k_, v_ in sorted(k_per_update.items(), key=lambda kv: -kv[1])},
        "family_ms_per_update": {k_: round(v_, 5) for k_, v_ in sorted(fam_ms.items(), key=lambda kv: -kv[1])},
        "counters_per_update": {k_: round(v_, 1) for k_, v_ in ctr.items() if k_ not in ("launches_total",)},
        "verified_fast_division": dict(zip(("voxel_size", "sigma"), r["fast_paths"])),
        "other_configs": others,
        "prefilter": prefilter,
    }
    bm = os.path.join(ROOT, "profiles", "build_manifest.json")
    if os.path.exists(bm):  # flags / compiler of the last build, and the hash of the library this process actually loaded
        import hashlib
        man = json.load(open(bm))
        line["build"] = {"nvcc": man["nvcc"], "flags": " ".join(man["flags"]), "sha256_manifest": man["sha256"],
                         "sha256_loaded": hashlib.sha256(open(dm.LIB_PATH, "rb").read()).hexdigest()}
    if not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import refmap
        import prefilter_oracle
        t0 = time.perf_counter()
        prefilter_oracle.preprocess(raw, 0.1, lo, hi, 5000)
        line["prefilter"]["cpu_port_ms"] = 1e3 * (time.perf_counter() - t0)   # numpy restatement, 1 thread (PCL itself is absent)
        if refmap.available(cfg_name):
            st = r["st"]
            pre, n_t = 12, 5   # a BOUNDED sample (about 7 s of CPU work); `--impl reference` times the CUDA arm's own frames
            rm = refmap.RefMap(cfg_name, seed=1, **SETTERS)
            futr = np.zeros((rm.V, rm.T), np.float32)
            ts = []
            for f in range(pre + n_t):
                s, _ = rm.timed_frame(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f], THRESHOLD, futr)
                if f >= pre:
                    ts.append(s)
            line["cpu_baseline"] = {"value": len(ts) / float(np.sum(ts)), "unit": "updates/s", "cores": 2, "kind": "reference", "frames": n_t,
                                    "sample": "frames %d..%d of the same stream; unmodified reference header (oracle/_ref), g++ -O2, "
                                              "1 thread + its 1 helper thread of %d host cores" % (pre, pre + n_t - 1, os.cpu_count())}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# one map sharded over N GPUs
# ---------------------------------------------------------------------------------------------------------------------
def run_sharded(dm, make_stream, torch, dist, cfg_name, K, W, rank, world, dev, stream, flush, nccl_id, detailed):
    cfg = dm.CONFIGS[cfg_name]
    PROF = min(K, 10) if detailed else 0
    CHECK = 6 if detailed else 0
    F1 = PREROLL + W + K + PROF
    F = F1 + W + K
    st = make_stream(cfg, seed=1, frames=F)     # the same stream on every rank: the cloud and the pose are replicated
    M = int(st["n"][0])
    est, tagged, nt_max, tg = precompute_tagged(dm, cfg, st, F1)
    d_pts = torch.from_numpy(st["points"][:F1]).to(dev)
    d_tag = torch.from_numpy(tg).to(dev)
    m = dm.DSPMap(cfg, seed=1, device=dev.index, max_points=max(M, nt_max, 1024))
    apply_setters(m)
    m.set_stream(stream.cuda_stream)
    m.shard_init(rank, world, nccl_id)
    d_xyz = torch.zeros((m.V, 3), dtype=torch.float32, device=dev)
    d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    d_fut = torch.zeros((m.V, m.T), dtype=torch.float32, device=dev)

    def step(f, d_p, d_t, nt):
        m.shard_update(M, d_p, st["pos"][f], st["t"][f], st["quat"][f], d_t, nt)
        m.shard_get_occupancy(THRESHOLD, d_xyz.data_ptr(), m.V, d_cnt.data_ptr(), d_fut.data_ptr())

    # correctness inside the run: rank 0 drives an unsharded map through the first frames and compares the results every rank
    # received from the sharded reader (occupied-voxel list bit for bit, future grid to the atomics' rounding)
    equal = None
    one = None
    if CHECK and rank == 0:
        one = dm.DSPMap(cfg, seed=1, device=dev.index, max_points=max(M, nt_max, 1024))
        apply_setters(one)
        one.set_stream(stream.cuda_stream)
        equal = True
    for f in range(PREROLL + W):
        step(f, d_pts[f].data_ptr(), d_tag[f].data_ptr(), len(tagged[f]))
        if f < CHECK:
            m.synchronize()
            n = int(d_cnt.item())
            if one is not None:
                one.update(M, 3, st["points"][f], *map(float, st["pos"][f]), float(st["t"][f]), *map(float, st["quat"][f]), tagged=tagged[f])
                n1, xyz1, fut1 = one.getOccupancyMapWithFutureStatus(THRESHOLD)
                equal = bool(equal and n == n1 and np.array_equal(d_xyz[:n].cpu().numpy().view(np.uint32), xyz1.view(np.uint32)) and
                             np.array_equal(d_fut.cpu().numpy() != 0, fut1 != 0) and np.allclose(d_fut.cpu().numpy(), fut1, rtol=4e-6, atol=0))
    if one is not None:
        one.close()
    torch.cuda.synchronize()
    dist.barrier()
    m.synchronize()
    launches0 = m.counters()["launches_total"]
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
    out = {"cfg": cfg, "M": M, "V": m.V, "T": m.T, "dev_ms": float(tt.item()), "launches": launches, "equal": equal, "info": m.shard_info()}

    if detailed:  # per-kernel and per-collective times on this rank + frame counters summed over ranks
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
        out["prof"], out["kprof"], out["P"] = m.profile_read(), m.profile_read_kernels(), PROF
        m.profile_enable(False)
        names = sorted(agg)
        ct = torch.tensor([agg[k_] for k_ in names], dtype=torch.float64, device=dev)
        dist.all_reduce(ct)
        out["ctr"] = {k_: float(v_) / PROF for k_, v_ in zip(names, ct.tolist())}

    # end to end: host cloud in (every rank: the cloud is replicated), host results out on rank 0
    fut_host = torch.zeros((m.V, m.T), dtype=torch.float32).pin_memory()
    xyz_host = torch.zeros((m.V, 3), dtype=torch.float32).pin_memory()
    last = tagged[-1]
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
        n_occ = 0
        if rank == 0:
            n_occ = int(d_cnt.item())
            fut_host.copy_(d_fut, non_blocking=True)
            xyz_host[:n_occ].copy_(d_xyz[:n_occ], non_blocking=True)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if k >= W:
            e2e_t.append(t1 - t0)
            h2d += pts.nbytes + last.nbytes
            d2h += (fut_host.numel() + 3 * n_occ) * 4 + 4 if rank == 0 else 0
    tt = torch.tensor([sum(e2e_t)], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    out.update(e2e_s=float(tt.item()), e2e_n=len(e2e_t), e2e_mean=float(np.mean(e2e_t)), h2d=h2d, d2h=d2h)
    m.close()
    return out


def main_sharded(args, cfg_name, dm, make_stream):
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    K, W = args.steps, args.warmup

    def fresh_id():  # every sharded map gets its own communicator: rank 0 draws the id, NCCL broadcasts the 128 bytes
        t = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(dm.shard_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=0)
        torch.cuda.synchronize()
        return bytes(t.cpu().numpy().tobytes())

    cfg = dm.CONFIGS[cfg_name]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    r = run_sharded(dm, make_stream, torch, dist, cfg_name, K, W, rank, world, dev, stream, flush, fresh_id(), detailed=True)
    clocks = sampler.stop() if rank == 0 else None
    others = {}
    if not args.only_headline:
        Kb = max(3, min(K, 10))
        for other in ("cfg5",):
            if other != cfg_name:
                o = run_sharded(dm, make_stream, torch, dist, other, Kb, W, rank, world, dev, stream, flush, fresh_id(), detailed=False)
                others[other] = {"value": Kb / (o["dev_ms"] * 1e-3), "unit": "updates/s", "ms_per_step": o["dev_ms"] / Kb, "steps": Kb,
                                 "e2e": {"value": o["e2e_n"] / o["e2e_s"], "unit": "updates/s", "ms_per_step": 1e3 * o["e2e_mean"]},
                                 "gather_records_per_rank": o["info"]["gather_records"], "config": workload_config(other, dm.CONFIGS[other])}
    if rank == 0:
        peak, peak_src = measured_peak()
        P, ctr = r["P"], r["ctr"]
        b_frame = dm.bytes_per_update(ctr, r["V"], r["T"], r["M"])
        k_per_update = {n: ms / P for n, (ms, ln) in r["kprof"].items() if ln}
        coll = {n: 1e3 * v for n, v in k_per_update.items() if n.startswith("coll_")}
        kern = {n: v for n, v in k_per_update.items() if not n.startswith("coll_")}
        top = max(kern, key=kern.get)
        ms_step = r["dev_ms"] / K
        line = {
            "metric": "map_updates_per_s", "value": K / (r["dev_ms"] * 1e-3), "unit": "updates/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg_name, cfg),
            "run": {"where": "hbm-resident", "parallelism": "ONE map, voxel z-slabs x%d, orchestrated by the library (dspmap_shard_update: C++ + NCCL)" % world,
                    "collectives_per_update": sorted(coll), "host_syncs_per_update": 1, "gather_records_per_rank": r["info"]["gather_records"],
                    "l2_flush_between_steps": flush is not None, "preroll_frames": PREROLL},
            "sharded_equals_single": r["equal"],
            "e2e": {"value": r["e2e_n"] / r["e2e_s"], "unit": "updates/s", "h2d_bytes_per_step": r["h2d"] // max(r["e2e_n"], 1),
                    "d2h_bytes_per_step": r["d2h"] // max(r["e2e_n"], 1), "ms_per_step": 1e3 * r["e2e_mean"]},
            "gpu_launches": int(r["launches"]), "launches_per_update": r["launches"] / K, "clocks": clocks,
            "collectives_us_per_update_rank0": {k_: round(v_, 1) for k_, v_ in sorted(coll.items(), key=lambda kv: -kv[1])},
            "limiting_collective": max(coll, key=coll.get) if coll else None,
            "roofline": {"bound": "hbm", "kernel": top, "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
                         "peak_source": peak_src, "kernel_ms": kern[top], "note": "rank 0's share; see the N=1 line for the per-kernel roofline"},
            "roofline_frame": {"bound": "hbm", "algorithmic_bytes_per_update": b_frame, "achieved": b_frame / (ms_step * 1e-3) / 1e9,
                               "peak": peak * world, "unit": "GB/s", "frac": b_frame / (ms_step * 1e-3) / 1e9 / (peak * world)},
            "kernel_ms_per_update_rank0": {k_: round(v_, 5) for k_, v_ in sorted(kern.items(), key=lambda kv: -kv[1])},
            "counters_per_update": {k_: round(v_, 1) for k_, v_ in ctr.items() if k_ not in ("launches_total", "launches_frame")},
            "other_configs": others,
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
    return 0


def main_replicas(args, cfg_name, dm, make_stream):
    """--replicas with N > 1: N independent maps, one per GPU (weak scaling), each measured like N = 1."""
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    K, W = args.steps, args.warmup
    dist.barrier()
    r = run_single(dm, make_stream, torch, cfg_name, K, W, dev, stream, flush, detailed=False)
    tt = torch.tensor([r["dev_ms"], float(np.sum(r["e2e_t"]))], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        line = {"metric": "map_updates_per_s", "value": world * K / (float(tt[0]) * 1e-3), "unit": "updates/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": float(tt[0]) / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(cfg_name, dm.CONFIGS[cfg_name]),
                "run": {"where": "hbm-resident", "parallelism": "replicas x%d (independent maps)" % world},
                "e2e": {"value": world * len(r["e2e_t"]) / float(tt[1]), "unit": "updates/s", "h2d_bytes_per_step": r["h2d"] // max(len(r["e2e_t"]), 1),
                        "d2h_bytes_per_step": r["d2h"] // max(len(r["e2e_t"]), 1)},
                "gpu_launches": int(r["launches"])}
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
    ap.add_argument("--only-headline", action="store_true", help="skip the brief cfg3 / cfg5 records")
    ap.add_argument("--replicas", action="store_true", help="N > 1: N independent maps instead of one sharded map")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    import dspmap_b200 as dm
    from dspmap_b200.streams import make_stream
    cfg_name = args.config
    if args.impl == "reference":
        return reference_arm(args, cfg_name, dm.CONFIGS[cfg_name])
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return (main_replicas if args.replicas else main_sharded)(args, cfg_name, dm, make_stream)
    return main_single(args, cfg_name, dm, make_stream)


if __name__ == "__main__":
    sys.exit(main())
