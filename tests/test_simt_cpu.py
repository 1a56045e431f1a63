"""CPU check of kernels by running the kernels' OWN source on the host: tests/simt/extract.py slices kernels out of
dsp-map_b200/csrc/dspmap_frame.cuh (and, into namespace legacy, out of tests/simt/legacy_frame_r01.cuh — the round-1 kernels
as they were verified on a B200), tests/simt/simt_host.h supplies threadIdx, __shared__, warp collectives and atomics (one
OS thread per CUDA thread), and tests/simt/check_*.cpp feeds the current kernel and the GPU-verified kernel it replaced the
same random inputs and requires bit-identical outputs.  No GPU involved."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMT = os.path.join(ROOT, "tests", "simt")
BUILD = os.path.join(ROOT, "tests", "_build", "simt")
CUH = os.path.join(ROOT, "dsp-map_b200", "csrc", "dspmap_frame.cuh")
LEGACY = os.path.join(SIMT, "legacy_frame_r01.cuh")


def build_and_run(check, inc, kernels, legacy=(), timeout=600, defines=(), tag="", sources=(), flags=()):
    os.makedirs(BUILD, exist_ok=True)
    subprocess.check_call([sys.executable, os.path.join(SIMT, "extract.py"), CUH, os.path.join(BUILD, inc)] + list(kernels))
    if legacy:
        subprocess.check_call([sys.executable, os.path.join(SIMT, "extract.py"), "--namespace", "legacy", LEGACY,
                               os.path.join(BUILD, "legacy_" + inc)] + list(legacy))
    exe = os.path.join(BUILD, check + tag)
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-I/usr/local/cuda/include", "-I" + SIMT, "-I" + BUILD,
                           "-I" + os.path.join(ROOT, "dsp-map_b200", "csrc")] + ["-D" + d for d in defines] +
                          list(flags) + [os.path.join(SIMT, check + ".cpp")] + list(sources) + ["-o", exe, "-lpthread"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_resampling_kernel_equals_the_round1_kernel():
    """k_resample (a warp walks RS_VPW voxels, one lane per voxel, over a shared-memory tile) against the warp-per-voxel kernel
    of round 1: particles, masks, occupancy / mean velocity, future grid and counters, bit for bit, on random voxels (empty,
    sparse, above MAX, full, heavy-tailed weights, weights below 1e-3, all four flag values)."""
    out = build_and_run("check_resample", "resample.inc", ["rs_warp_bytes", "k_resample"], legacy=["k_resample"], defines=("GRID=3",))
    assert out.count("identical") == 6 and "DIFFERENT" not in out


def test_newborn_placement_kernel_equals_the_round1_kernel():
    """k_nb_place (one thread per candidate: rank by counting inside the voxel's grouped segment, rank-th free slot of the
    mask snapshot) against round 1's warp-per-voxel minimum extraction: same candidates in the same slots, same masks, same
    count, on random voxels (full, nearly full, 1 .. 330 candidates per voxel)."""
    out = build_and_run("check_nb_place", "nb_place.inc", ["nb_place_born", "k_nb_place"], legacy=["k_nb_place"])
    assert out.count("identical") == 4 and "DIFFERENT" not in out


def test_device_estimation_front_end_equals_the_host_estimator():
    """dspmap_estimator.cuh (FOV filter, ground split, hash-grid union-find clustering, PCL's cluster order, centroids, layout
    of the tagged cloud) + the host matching, against VelocityEstimator::estimate on the same random scenes (blobs of 1..600
    points, some above 1.5 m, a wall, ground points around the threshold, points behind the sensor, a frame with nothing in
    view), dynamic and static model: the tagged clouds — positions, velocities, colours, order — bit for bit, 14 frames."""
    out = build_and_run("check_estimator", "est_scan.inc", ["block_exclusive_scan"], flags=("-ffp-contract=off",),
                        sources=(os.path.join(ROOT, "dsp-map_b200", "csrc", "velocity_estimator.cpp"),))
    assert out.count("identical") == 14 and "DIFFERENT" not in out
