#!/usr/bin/env python3
"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`):
    python profiles/summarize_launches.py launches.csv [frames]
Times under ncu are cold-cache and serialised: compare shares, not absolutes (see profiles/README.md)."""
import collections
import csv
import sys


def main():
    path, frames = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else v * 1000 if row["Metric Unit"] == "ms" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("total us/frame (all kernels) %.1f" % (tot / frames))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-32s %7.1f us/frame %5.1f%%  n/frame %.1f" % (k[:32], v[1] / frames, 100 * v[1] / tot, v[0] / frames))


if __name__ == "__main__":
    main()
