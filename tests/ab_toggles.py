#!/usr/bin/env python3
"""A/B harness for the library's experiment switches (environment variables read by dspmap_create).

For every switch set given on the command line (e.g. `DSPMAP_PDL=0` or `DSPMAP_PDL=0,DSPMAP_EST_THREAD=0`: the remaining switches are defaults that NAME=0 turns off) it creates a
baseline map (no switches) and a switched map with the same seeds, feeds both the same stream and requires the complete
map state to be bit-identical after every frame (tests/parity.py: compare_state; the atomically accumulated future grid
within rtol 2e-6); then it times the device-resident frame (update_device + get_occupancy_device, CUDA events, L2 flushed
between steps) and the host-pointer frame (update + getOccupancyMapWithFutureStatus, wall clock) of both.

    python tests/ab_toggles.py [--cfg cfg2] [--frames 10] [--steps 40] [--also cfg3:2] DSPMAP_PDL=1 ...

Every switch set runs in its own process with a timeout (--timeout), so a kernel that hangs or faults under one switch
costs that set only; the default path is timed once, by the first child.

Results are appended line by line to gpurun_out/ab_toggles.jsonl (flushed as it goes, so a run cut short keeps what it had).
This is a measurement tool, not part of the test-suite (pytest does not collect it)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dsp-map_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

SWITCHES = ("DSPMAP_PDL", "DSPMAP_EST_THREAD", "DSPMAP_ASYNC_UPDATE", "DSPMAP_NORM_POLL", "DSPMAP_EST_GPU", "DSPMAP_NB_POS")  # run-time switches


def make_map(dm, gpu_map, name, env, **kw):
    for k in SWITCHES:
        os.environ.pop(k, None)
    os.environ.update(env)
    g = gpu_map(name, seed=7, **kw)
    for k in env:
        os.environ.pop(k, None)
    return g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="cfg2")
    ap.add_argument("--frames", type=int, default=10, help="parity frames (host-pointer path)")
    ap.add_argument("--steps", type=int, default=40, help="timed frames per path")
    ap.add_argument("--preroll", type=int, default=25, help="untimed frames before the timed ones")
    ap.add_argument("--also", default="", help="extra parity-only configs, name:frames[,name:frames]")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ab_toggles.jsonl"))
    ap.add_argument("--timeout", type=float, default=90.0, help="seconds per switch set (each set runs in its own process)")
    ap.add_argument("--inproc", action="store_true", help="run all sets in this process (no protection against a hung kernel)")
    ap.add_argument("--baseline", default="", help="(internal) baseline device_ms,host_ms measured by the driver process")
    ap.add_argument("sets", nargs="+")
    args = ap.parse_args()

    if not args.inproc:
        # one process per switch set, so that a kernel that hangs or faults under one switch costs that set only
        import subprocess
        common = [sys.executable, os.path.abspath(__file__), "--inproc", "--cfg", args.cfg, "--frames", str(args.frames), "--steps", str(args.steps),
                  "--preroll", str(args.preroll), "--out", args.out]
        if args.also:
            common += ["--also", args.also]
        base = ""
        if args.steps > 0:
            r = subprocess.run(common + ["--baseline", "measure", "BASELINE"], capture_output=True, text=True, timeout=args.timeout)
            for line in r.stdout.splitlines():
                if line.startswith("BASELINE "):
                    base = line.split()[1]
            print("baseline (device ms, host-api ms):", base or "failed: " + r.stderr[-300:], flush=True)
        rc = 0
        for spec in args.sets:
            try:
                r = subprocess.run(common + (["--baseline", base] if base else []) + [spec], capture_output=True, text=True, timeout=args.timeout)
                sys.stdout.write(r.stdout)
                if r.returncode != 0:
                    rc = 1
                    rec = {"switches": spec, "cfg": args.cfg, "stage": "error", "returncode": r.returncode, "stderr": r.stderr[-600:]}
                    open(args.out, "a").write(json.dumps(rec) + "\n")
                    print(json.dumps(rec), flush=True)
            except subprocess.TimeoutExpired:
                rc = 1
                rec = {"switches": spec, "cfg": args.cfg, "stage": "timeout", "seconds": args.timeout}
                open(args.out, "a").write(json.dumps(rec) + "\n")
                print(json.dumps(rec), flush=True)
        return rc

    import torch
    import dspmap_b200 as dm
    from common import gpu_map, gpu_update, make_stream
    from parity import compare_state, same

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    out = open(args.out, "a")

    def emit(rec):
        out.write(json.dumps(rec) + "\n")
        out.flush()
        print(json.dumps(rec), flush=True)

    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def parity(name, frames, env):
        cfg = dm.CONFIGS[name]
        st = make_stream(cfg, seed=3, frames=frames + 2)
        est = dm.VelocityEstimator(cfg, seed=7, filter_res=0.1)
        a = make_map(dm, gpu_map, name, {}, max_points=cfg["points"])
        b = make_map(dm, gpu_map, name, env, max_points=cfg["points"])
        bad = []
        fpin = np.full((b.V, b.T), 7.0, np.float32)   # the switched map reads into ONE registered buffer, like the drop-in header
        b.pin_host_buffer(fpin)
        for f in range(frames):
            pts, pos, t, q = st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]
            tc = est.estimate(pts, pos, t, q)
            ra, rb = gpu_update(a, pts, pos, t, q, tagged=tc), gpu_update(b, pts, pos, t, q, tagged=tc)
            if ra != rb:
                bad.append("frame %d: return codes %d / %d" % (f, ra, rb))
            bad += compare_state(a, b, label="frame %d:" % f)
            if f % 3 != 1:   # two frames out of three, so that rows of an older call have to be cleared too
                na, xa, fa = a.getOccupancyMapWithFutureStatus(0.2)
                nb, xb, fb = b.getOccupancyMapWithFutureStatus(0.2, fpin)
                if not same(xa, xb) or not np.array_equal(fa != 0, fb != 0) or not np.allclose(fa, fb, rtol=2e-6, atol=0):
                    bad.append("frame %d: reader output differs" % f)
            if bad:
                break
        # the library's own estimator path (no explicit newborn input) on two more frames
        if not bad:
            for f in range(frames, frames + 2):
                pts, pos, t, q = st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]
                ra, rb = gpu_update(a, pts, pos, t, q), gpu_update(b, pts, pos, t, q)
                if ra != rb or not same(a.getKMClusterResult(), b.getKMClusterResult()):
                    bad.append("estimator path frame %d: rc %d / %d or tagged cloud differs" % (f, ra, rb))
                bad += compare_state(a, b, label="estimator path frame %d:" % f)
        a.close()
        b.close()
        return bad

    tcache = {}

    def timing_inputs(name, steps):
        if (name, steps) in tcache:
            return tcache[(name, steps)]
        cfg = dm.CONFIGS[name]
        pre, W = args.preroll, 3
        F = pre + W + steps
        st = make_stream(cfg, seed=1, frames=2 * F)
        M = int(st["n"][0])
        est = dm.VelocityEstimator(cfg, seed=1, filter_res=0.1)
        tagged, last = [], np.zeros((0, 7), np.float32)
        for f in range(F):
            t = est.estimate(st["points"][f], st["pos"][f], st["t"][f], st["quat"][f])
            last = t if t is not None else last
            tagged.append(last)
        nt_max = max(max(len(t) for t in tagged), 1)
        d_pts = torch.from_numpy(st["points"][:F]).to(dev)
        tg = np.zeros((F, nt_max, 7), np.float32)
        for f, t in enumerate(tagged):
            tg[f, :len(t)] = t
        d_tag = torch.from_numpy(tg).to(dev)
        tcache[(name, steps)] = (pre, W, F, st, M, tagged, nt_max, d_pts, d_tag)
        return tcache[(name, steps)]

    def timing(name, env, steps):
        pre, W, F, st, M, tagged, nt_max, d_pts, d_tag = timing_inputs(name, steps)
        m = make_map(dm, gpu_map, name, env, max_points=max(M, nt_max, 1024))
        m.set_stream(stream.cuda_stream)
        d_xyz = torch.empty((m.V, 3), dtype=torch.float32, device=dev)
        d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        d_fut = torch.empty((m.V, m.T), dtype=torch.float32, device=dev)

        def step(f):
            m.update_device(M, d_pts[f].data_ptr(), st["pos"][f], st["t"][f], st["quat"][f], d_tag[f].data_ptr(), len(tagged[f]))
            m.get_occupancy_device(0.2, d_xyz.data_ptr(), m.V, d_cnt.data_ptr(), d_fut.data_ptr())

        for f in range(pre + W):
            step(f)
        m.synchronize()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for k, (e0, e1) in enumerate(ev):
            flush.zero_()
            e0.record(stream)
            step(pre + W + k)
            e1.record(stream)
        torch.cuda.synchronize()
        m.synchronize()
        dev_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        fut_host = np.zeros((m.V, m.T), np.float32)
        m.pin_host_buffer(fut_host)
        ts = []
        for k in range(W + steps):
            f = F + k
            pts, pos, t, q = st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            gpu_update(m, pts, pos, t, q)
            m.getOccupancyMapWithFutureStatus(0.2, fut_host)
            if k >= W:
                ts.append(time.perf_counter() - t0)
        # per-kernel device time (CUDA events around every launch site) on the frames of the stream that are left
        m.profile_enable(True)
        NP = min(10, pre)
        for k in range(NP):
            f = 2 * F - pre + k
            flush.zero_()
            gpu_update(m, st["points"][f], st["pos"][f], st["t"][f], st["quat"][f])
            m.get_occupancy_device(0.2, d_xyz.data_ptr(), m.V, d_cnt.data_ptr(), d_fut.data_ptr())
        m.synchronize()
        kern = {n: round(1e3 * ms / NP, 2) for n, (ms, ln) in m.profile_read_kernels().items() if ln}
        m.profile_enable(False)
        m.close()
        return dev_ms, 1e3 * float(np.mean(ts)), kern

    base = None
    if args.baseline == "measure":  # driver's first child: time the default path once for all sets
        b_dev, b_host, b_kern = timing(args.cfg, {}, args.steps)
        json.dump({"dev": b_dev, "host": b_host, "kernels": b_kern}, open(args.out + ".baseline", "w"))
        print("BASELINE %r,%r" % (b_dev, b_host), flush=True)
        return 0
    if args.baseline:
        bj = json.load(open(args.out + ".baseline"))
        base = (bj["dev"], bj["host"], bj["kernels"])
    for spec in args.sets:
        env = dict(kv.split("=", 1) for kv in spec.split(",") if kv)
        rec = {"switches": env, "cfg": args.cfg}
        t0 = time.time()
        bad = parity(args.cfg, args.frames, env)
        rec["parity_frames"] = args.frames
        rec["parity"] = "bit-identical" if not bad else bad[:6]
        emit(dict(rec, stage="parity", seconds=round(time.time() - t0, 1)))
        for extra in [e for e in args.also.split(",") if e]:
            nme, fr = extra.split(":")
            t0 = time.time()
            bad2 = parity(nme, int(fr), env)
            emit({"switches": env, "cfg": nme, "stage": "parity", "parity_frames": int(fr),
                  "parity": "bit-identical" if not bad2 else bad2[:6], "seconds": round(time.time() - t0, 1)})
        if args.steps > 0:
            if base is None:
                base = timing(args.cfg, {}, args.steps)
            base_dev, base_host, base_kern = base
            sw_dev, sw_host, sw_kern = timing(args.cfg, env, args.steps)
            # kernels whose device time per frame moved by more than 1 us (replaying already-seen frames: comparable state)
            moved = {n: [base_kern.get(n), sw_kern.get(n)] for n in sorted(set(base_kern) | set(sw_kern))
                     if abs((base_kern.get(n) or 0.0) - (sw_kern.get(n) or 0.0)) > 1.0}
            emit({"switches": env, "cfg": args.cfg, "stage": "timing", "steps": args.steps, "kernel_us_per_frame_baseline_vs_switched": moved,
                  "device_ms_per_frame": {"baseline": base_dev, "switched": sw_dev},
                  "host_api_ms_per_frame": {"baseline": base_host, "switched": sw_host},
                  "updates_per_s_device": {"baseline": 1e3 / base_dev, "switched": 1e3 / sw_dev},
                  "updates_per_s_host_api": {"baseline": 1e3 / base_host, "switched": 1e3 / sw_host}})
    return 0


if __name__ == "__main__":
    sys.exit(main())
