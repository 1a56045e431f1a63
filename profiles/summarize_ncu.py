#!/usr/bin/env python3
"""Markdown summary of an ncu report captured with --set full --import-source on (reads it with `ncu -i`, no GPU needed).

    python profiles/summarize_ncu.py REPORT.ncu-rep [REPORT2.ncu-rep ...] > profiles/rNN_top_kernels.md

Per kernel: duration, DRAM bytes, L2 hit rate, occupancy, issue utilisation, executed warp instructions, registers, launch
geometry, average cycles between two issues of a warp, the stall mix and the hottest SASS instructions (profiles/hot_lines.py)."""
import csv
import subprocess
import sys

METRICS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
           ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"), ("smsp__inst_executed.sum", "warp instructions"),
           ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
           ("launch__waves_per_multiprocessor", "waves per SM"), ("smsp__average_warp_latency_per_inst_issued.ratio", "cycles between issues of a warp")]


def main():
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()
        rows = list(csv.reader(raw))
        h, units = rows[0], rows[1]
        print("## %s\n" % rep)
        seen = set()
        for r in rows[2:]:
            name = r[h.index("Kernel Name")]
            short = name.split("(")[0].replace("void ", "")
            if short in seen:
                continue
            seen.add(short)
            print("### %s\n" % short)
            for m, label in METRICS:
                if m in h:
                    print("- %s: %s %s" % (label, r[h.index(m)], units[h.index(m)]))
            hot = subprocess.run([sys.executable, __file__.replace("summarize_ncu.py", "hot_lines.py"), rep, "^" + short.split("<")[0] + "$", "8"],
                                 capture_output=True, text=True).stdout.splitlines()
            print("\n```")
            for line in hot[1:]:
                print(line[:150])
            print("```\n")


if __name__ == "__main__":
    main()
