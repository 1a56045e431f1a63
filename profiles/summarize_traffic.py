#!/usr/bin/env python3
"""DRAM / L2 traffic of one map update from an ncu pass that leaves the caches alone between kernels
(`ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
lts__t_bytes.sum --csv`, captured over `frames` updates of bench.py's timed region; tests/gpu_profile_pass.sh):
    python profiles/summarize_traffic.py traffic.csv frames [traffic.json]
Prints per-kernel bytes per update and, with a third argument, writes the JSON bench.py reads into roofline.traffic (the top
kernel, per launch) and roofline_frame.traffic ("frame": all of the library's kernels of one update; torch's L2-flush fill
kernel is listed but not counted).  With the caches left alone a consumer finds its producer's output in the 126 MB L2:
these are the bytes that actually reach HBM inside a frame, not the cold-cache figure of a default ncu pass."""
import collections
import csv
import json
import sys


def main():
    path, frames = sys.argv[1], int(sys.argv[2])
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit, metric = row["Metric Unit"], row["Metric Name"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(unit, 1)
        a = per.setdefault(name, collections.defaultdict(float))
        a[metric] += v * scale
        if metric == "gpu__time_duration.sum":
            a["n"] += 1
    out, frame = {}, collections.defaultdict(float)
    print("%-34s %6s %9s %10s %10s %10s" % ("kernel", "n/upd", "us/upd", "dram rd", "dram wr", "L2 bytes"))
    for k, a in sorted(per.items(), key=lambda kv: -(kv[1]["dram__bytes_read.sum"] + kv[1]["dram__bytes_write.sum"])):
        rd, wr, l2, us, n = (a["dram__bytes_read.sum"] / frames, a["dram__bytes_write.sum"] / frames, a["lts__t_bytes.sum"] / frames,
                             a["gpu__time_duration.sum"] / frames, a["n"] / frames)
        print("%-34s %6.1f %9.1f %10.0f %10.0f %10.0f" % (k[:34], n, us, rd, wr, l2))
        ours = k.startswith("k_")
        if ours:
            out[k.split("<")[0]] = (rd + wr) / max(n, 1e-9)  # per launch
            frame["dram"] += rd + wr
            frame["l2"] += l2
            frame["us"] += us
    print("library kernels per update: DRAM %.2f MB, L2 %.2f MB, %.1f us under ncu" % (frame["dram"] / 1e6, frame["l2"] / 1e6, frame["us"]))
    if len(sys.argv) > 3:
        out["frame"] = frame["dram"]
        out["frame_l2_bytes"] = frame["l2"]
        out["source"] = "ncu --cache-control none over %d updates of bench.py's timed region (profiles/summarize_traffic.py)" % frames
        json.dump(out, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
