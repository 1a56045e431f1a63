#!/usr/bin/env python3
"""Builds dsp-map_b200/lib/libdspmap_b200.so: hand-written sm_100a kernels + the C-ABI (include/dspmap_b200.h).

nvcc cross-compiles without a GPU. -fmad=false, IEEE division / square root and no flush-to-zero are REQUIRED: the
kernels reproduce the reference's fp32 results bit for bit.  The built .so is git-ignored but travels
to the GPU box with gpurun."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib")
SO = os.path.join(OUT, "libdspmap_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-shared", "-cudart", "static"]


def sources():
    return [os.path.join(SRC, f) for f in ("dspmap.cu", "prefilter.cu", "velocity_estimator.cpp")]


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(HERE, "..", "include", "dspmap_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out: a compile-time variant (-DNAME=value ...) written to another path, for A/B runs on the GPU box."""
    if out is None and not force and not stale():
        return SO
    os.makedirs(OUT, exist_ok=True)
    target = out or SO
    cmd = [NVCC] + FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + sources() + ["-o", target]
    subprocess.check_call(cmd)
    if out is None:
        write_manifest()
    return target


def write_manifest():
    """profiles/build_manifest.json: what exactly was compiled (bench.py copies it into its JSON line)."""
    import hashlib
    import json
    ver = subprocess.run([NVCC, "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2:]
    h = hashlib.sha256(open(SO, "rb").read()).hexdigest()
    srcs = {os.path.basename(f): hashlib.sha256(open(os.path.join(SRC, f), "rb").read()).hexdigest()[:16] for f in sorted(os.listdir(SRC))}
    man = {"library": "dsp-map_b200/lib/libdspmap_b200.so", "sha256": h, "nvcc": " / ".join(ver), "flags": FLAGS, "sources_sha256_16": srcs}
    path = os.path.join(HERE, "..", "profiles", "build_manifest.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(man, f, indent=1)
        f.write("\n")


if __name__ == "__main__":
    # build.py [--force] [-v] [--out PATH -DNAME=value ...]
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=out))
