#!/usr/bin/env python3
"""Builds oracle/_ref/libdspref_<cfg>.so from the UNMODIFIED reference headers under /root/reference.

TEST INFRASTRUCTURE ONLY. Nothing from the reference is copied into the repository: the per-config header
is produced at build time by the same line substitutions the reference's own tuner performs
(script/set_map_parameters.py:392-452 rewrites the `#define` / `const int` lines) and written to
oracle/_ref/gen/ (git-ignored). The .so files are git-ignored too but travel to the GPU box with gpurun.

    python oracle/build_ref.py            # all configs
    python oracle/build_ref.py cfg2 cfg5  # selected
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dsp-map_b200"))
from dspmap_b200.configs import CONFIGS  # noqa: E402

REF_INCLUDE = "/root/reference/include"
OUT = os.path.join(ROOT, "oracle", "_ref")


def _fmt(x):
    s = repr(float(x))
    return s + "f"


def patch_header(src, cfg):
    subs = [
        (r"#define MAP_LENGTH_VOXEL_NUM \d+", "#define MAP_LENGTH_VOXEL_NUM %d" % cfg["nx"]),
        (r"#define MAP_WIDTH_VOXEL_NUM \d+", "#define MAP_WIDTH_VOXEL_NUM %d" % cfg["ny"]),
        (r"#define MAP_HEIGHT_VOXEL_NUM \d+", "#define MAP_HEIGHT_VOXEL_NUM %d" % cfg["nz"]),
        (r"#define VOXEL_RESOLUTION [0-9.]+", "#define VOXEL_RESOLUTION %r" % cfg["res"]),
        (r"#define ANGLE_RESOLUTION \d+", "#define ANGLE_RESOLUTION %d" % cfg["angle_res"]),
        (r"#define MAX_PARTICLE_NUM_VOXEL \d+", "#define MAX_PARTICLE_NUM_VOXEL %d" % cfg["max_ppv"]),
        (r"const int half_fov_h = \d+;", "const int half_fov_h = %d;" % cfg["half_fov_h"]),
        (r"const int half_fov_v = \d+;", "const int half_fov_v = %d;" % cfg["half_fov_v"]),
        (r"#define PREDICTION_TIMES \d+", "#define PREDICTION_TIMES %d" % len(cfg["future_times"])),
        (r"prediction_future_time\[PREDICTION_TIMES\] = \{[^}]*\}",
         "prediction_future_time[PREDICTION_TIMES] = {%s}" % ", ".join(_fmt(t) for t in cfg["future_times"])),
    ]
    if "PYRAMID_NEIGHBOR_N" in src:
        subs.append((r"#define PYRAMID_NEIGHBOR_N \d+", "#define PYRAMID_NEIGHBOR_N %d" % cfg["neighbor_n"]))
    for pat, rep in subs:
        src, n = re.subn(pat, rep, src, count=1)
        if n != 1:
            raise RuntimeError("substitution %r matched %d times" % (pat, n))
    return src


# The reference's own CMakeLists.txt:4 compiles with -O3 -ftree-vectorize -ffast-math -march=native.  The parity libraries
# cannot use those flags (-ffast-math changes results, -march=native does not travel to another box); for TIMING the
# reference arm of bench.py also gets a "<cfg>_fast" library with the reference's flags, -march=native replaced by the
# portable -mavx2 -mfma (bench.py checks /proc/cpuinfo before loading it).
FAST_FLAGS = ["-O3", "-ftree-vectorize", "-ffast-math", "-mavx2", "-mfma"]
FAST_CONFIGS = ("cfg2",)


def build(name, fast=False):
    cfg = CONFIGS[name]
    os.makedirs(os.path.join(OUT, "gen", name), exist_ok=True)
    with open(os.path.join(REF_INCLUDE, cfg["header"])) as f:
        src = f.read()
    gen = os.path.join(OUT, "gen", name, "dsp_ref.h")
    with open(gen, "w") as f:
        f.write(patch_header(src, cfg))
    so = os.path.join(OUT, "libdspref_%s%s.so" % (name, "_fast" if fast else ""))
    V = cfg["nx"] * cfg["ny"] * cfg["nz"]
    cmd = ["g++", "-std=c++14"] + (FAST_FLAGS if fast else ["-O2", "-ffp-contract=off"]) + ["-fPIC", "-shared", "-Wl,-Bsymbolic", "-w",
           # function-local statics of the reference's inline methods must stay private to each loaded copy
           "-fno-gnu-unique",
           "-I", os.path.join(ROOT, "oracle", "shim"), '-DREF_HEADER="%s"' % gen]
    if cfg["model"] == "static":
        cmd.append("-DREF_STATIC")
    if V * cfg["max_ppv"] * 2 * 36 > 1.5e9:  # file-static particle store > 2 GB needs the medium code model
        cmd.append("-mcmodel=medium")
    cmd += [os.path.join(ROOT, "oracle", "ref_driver.cpp"), "-o", so, "-lpthread"]
    try:
        subprocess.check_call(cmd)
    finally:
        os.remove(gen)  # the substituted header is a build intermediate only
    return so


if __name__ == "__main__":
    if not os.path.isdir(REF_INCLUDE):
        print("reference tree not present; keeping prebuilt oracle/_ref/*.so")
        sys.exit(0)
    names = sys.argv[1:] or list(CONFIGS)
    for n in names:
        print("built", build(n))
        if n in FAST_CONFIGS:
            print("built", build(n, fast=True))
