#!/usr/bin/env python3
"""Top stall sites of one kernel from an ncu report captured with --set full --import-source on.

    python profiles/hot_lines.py REPORT.ncu-rep KERNEL_REGEX [N]

Reads `ncu -i REPORT --page source --csv --kernel-name regex:KERNEL_REGEX` (SASS view: one row per instruction), prints the
kernel's total samples / executed warp instructions and the N instructions with the most stall samples with their
dominant stall reasons.  With -lineinfo the CUDA-C view (`--page source` default when sources were imported) can be
read the same way."""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                         capture_output=True, text=True).stdout.splitlines()
    rows = list(csv.reader(out))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[hdr_i - 1][1] if hdr_i else "")
    h = rows[hdr_i]
    col = {name: k for k, name in enumerate(h)}
    stall = [k for k, name in enumerate(h) if name.startswith("stall_") and "Not Issued" not in name]
    body, seen = [], set()
    for r in rows[hdr_i + 1:]:  # a report can hold several views of the kernel: keep the first row per address
        if len(r) == len(h) and r[0] != "Address" and r[0] not in seen:
            seen.add(r[0])
            body.append(r)

    def num(x):
        try:
            return float(x)
        except ValueError:
            return 0.0
    tot_s = sum(num(r[col["# Samples"]]) for r in body)
    tot_i = sum(num(r[col["Instructions Executed"]]) for r in body)
    print("instructions (SASS): %d   samples: %d   executed warp instructions: %d" % (len(body), tot_s, tot_i))
    agg = {}
    for r in body:
        for k in stall:
            agg[h[k]] = agg.get(h[k], 0) + num(r[k])
    print("stall mix:", ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(tot_s, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
    ops = {}
    for r in body:
        t = r[col["Source"]].split()
        op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
        o = ops.setdefault(op, [0.0, 0.0])
        o[0] += num(r[col["# Samples"]])
        o[1] += num(r[col["Instructions Executed"]])
    print("by opcode:", ", ".join("%s %.1f%% (%d)" % (k, 100 * v[0] / max(tot_s, 1), v[1]) for k, v in sorted(ops.items(), key=lambda kv: -kv[1][0])[:12]))
    body.sort(key=lambda r: -num(r[col["# Samples"]]))
    for r in body[:n]:
        s = num(r[col["# Samples"]])
        why = sorted(((num(r[k]), h[k][6:]) for k in stall), reverse=True)[:2]
        print("%5.1f%%  exec %9d  %-58s %s" % (100 * s / max(tot_s, 1), num(r[col["Instructions Executed"]]), r[col["Source"]][:58],
                                               " ".join("%s:%d" % (nm, v) for v, nm in why if v)))


if __name__ == "__main__":
    main()
