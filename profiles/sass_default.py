#!/usr/bin/env python3
"""SASS evidence for the default path:  python profiles/sass_default.py [lib.so] > profiles/r02_sass_default.txt
Per kernel of the library: instruction count, registers, and how often the mnemonics that matter appear (UBLKCP = cp.async.bulk
/ TMA 1-D bulk copy, SYNCS = mbarrier, LDGSTS = cp.async, ATOM*/RED = atomics, BAR = block barriers, SHFL / VOTE / MATCH /
REDUX = warp collectives, MUFU.RCP / FCHK = the IEEE division sequence, ACQBULK / griddepcontrol for programmatic
dependent launch); then the full SASS of k_pair_eval, the kernel that stages its operands with bulk copies."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "dsp-map_b200/lib/libdspmap_b200.so"
KEYS = ["UBLKCP", "SYNCS", "LDGSTS", "ATOMG", "ATOMS", "RED", "BAR", "SHFL", "VOTE", "MATCH", "REDUX", "MUFU.RCP", "FCHK", "ACQBULK", "LDG", "STG", "LDS", "STS", "FADD", "FMUL", "FFMA"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = dict(re.findall(r"Function (\S+):\n\s*REG:(\d+)", res))
blocks = re.split(r"\n\s*Function : ", sass)[1:]
print("library: %s\n" % lib)
print("%-44s %6s %5s  %s" % ("kernel", "instr", "regs", "mnemonics"))
full = {}
for b in blocks:
    name = b.split("\n", 1)[0].strip()
    short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
    ins = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(.*?);", b, re.M)
    cnt = collections.Counter()
    for i in ins:
        op = re.sub(r"^@!?U?P\d+\s+", "", i).split()[0]
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                cnt[k] += 1
    print("%-44s %6d %5s  %s" % (short[:44], len(ins), regs.get(name, "?"), " ".join("%s:%d" % kv for kv in cnt.items() if kv[0] not in ("LDG", "STG", "LDS", "STS", "FADD", "FMUL", "FFMA") or True)))
    full[short] = b
print("\n\n==== full SASS of k_pair_eval (UBLKCP = cp.async.bulk global -> shared, SYNCS = mbarrier arrive / try_wait) ====\n")
print("Function : " + full.get("k_pair_eval", "(not found)"))
