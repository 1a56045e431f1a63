"""The drop-in headers (include/dsp_dynamic.h, dsp_dynamic_multiple_neighbors.h, dsp_static.h) compile against the
application-side code pattern of src/map_sim_example.cpp and, on a GPU, produce the reference's outputs."""
import os
import struct
import subprocess

import numpy as np
import pytest

import dspmap_b200 as dm
import refmap
from common import make_stream

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")


def build(header):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "dropin_" + header.replace(".h", ""))
    cmd = ["g++", "-std=c++14", "-O1", "-w", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "shim"),
           '-DDSPMAP_HEADER="%s"' % header, os.path.join(ROOT, "tests", "dropin_main.cpp"), "-o", exe,
           "-L", os.path.join(ROOT, "dsp-map_b200", "lib"), "-ldspmap_b200", "-Wl,-rpath," + os.path.join(ROOT, "dsp-map_b200", "lib"),
           "-lpthread", "-ldl", "-lrt"]
    subprocess.check_call(cmd)
    return exe


@pytest.mark.parametrize("header", ["dsp_dynamic.h", "dsp_dynamic_multiple_neighbors.h", "dsp_static.h"])
def test_dropin_headers_compile_and_link(header):
    dm.load_library()
    assert os.path.exists(build(header))


REF_APP = "/root/reference/src/map_sim_example.cpp"


@pytest.mark.skipif(not os.path.exists(REF_APP), reason="the reference tree is only present in the build container")
def test_reference_application_compiles_unchanged():
    """north_star: "drops into map_sim_example.cpp unchanged".  The reference's application, byte for byte as it lies in
    /root/reference (global `DSPMap my_map;` ex:39, the map macros ex:52-57 / 400-403, VOXEL_NUM / PREDICTION_TIMES ex:371,
    bare queue / vector / string / endl ex:42-44 / 342 / 443, update ex:345-349, the readers ex:378, the setters ex:522-528,
    getVoxelPositionFromIndexPublic ex:409), type-checks against include/dsp_dynamic.h.  ROS, PCL's filters and the parts of
    Eigen only the application uses are declarations-only stand-ins (tests/shim_ros)."""
    src = open(REF_APP).read()
    for needle in ('#include "dsp_dynamic.h"', "DSPMap my_map;", "my_map.update(", "my_map.getOccupancyMapWithFutureStatus(",
                   "static float future_status[VOXEL_NUM][PREDICTION_TIMES];", "my_map.setPredictionVariance(",
                   "DSPMap::setOriginalVoxelFilterResolution(", "my_map.setParticleRecordFlag(", "my_map.getVoxelPositionFromIndexPublic("):
        assert needle in src, needle
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "tests", "shim_ros"), "-I", os.path.join(ROOT, "oracle", "shim"), REF_APP],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


@pytest.mark.gpu
def test_dropin_application_matches_reference_bit_for_bit():
    """dsp_dynamic.h as shipped (66x66x40, 9 ppv) == config 'ref_default'.  The header seeds from the wall clock like the
    reference (dsp_dynamic.h:586, 1151); DSPMAP_TABLE_SEED / DSPMAP_UNIFORM_SEED pin both generators, so the unchanged
    application pattern (tests/dropin_main.cpp), with the library's OWN velocity estimation, must print exactly what the
    reference computes with the same seed."""
    name = "ref_default"
    if not refmap.available(name):
        pytest.skip("reference library not present")
    exe = build("dsp_dynamic.h")
    cfg = dm.CONFIGS[name]
    F, seed = 5, 11
    st = make_stream(cfg, seed=2, frames=F)
    path = os.path.join(BUILD, "stream.bin")
    from dspmap_b200.streams import write_stream
    write_stream(path, st)
    env = dict(os.environ, DSPMAP_TABLE_SEED=str(seed), DSPMAP_UNIFORM_SEED=str(seed))
    out = subprocess.check_output([exe, path], text=True, env=env)
    lines = [l for l in out.splitlines() if l.startswith("frame ")]
    assert len(lines) == F
    r = refmap.RefMap(name, seed=seed)
    for k in range(F):
        r.update(st["points"][k], st["pos"][k], st["t"][k], st["quat"][k])
        xyz, fut = r.occupancy(0.2)
        tok = lines[k].split()
        assert int(tok[3]) == len(xyz) and int(tok[5]) == len(xyz)      # occupied voxels; the cloud got exactly those
        assert int(tok[9]) == len(r.tagged_cloud())                      # getKMClusterResult size
        first = xyz[0] if len(xyz) else np.zeros(3, np.float32)
        assert [float(x) for x in tok[11:14]] == [float("%.4f" % v) for v in first]
        # the future grid is accumulated with fp32 atomics on the GPU: its total agrees to rounding
        assert abs(float(tok[7]) - float(fut.astype(np.float64).sum())) <= 2e-6 * float(fut.astype(np.float64).sum()) + 1e-6


def test_replay_tool_compiles():
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "dspmap_replay")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-w", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "shim"),
                           os.path.join(ROOT, "dsp-map_b200", "tools", "dspmap_replay.cpp"), "-o", exe,
                           "-L", os.path.join(ROOT, "dsp-map_b200", "lib"), "-ldspmap_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "dsp-map_b200", "lib"), "-lpthread", "-ldl", "-lrt"])
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_replay_tool_outputs_equal_the_reference(tmp_path):
    """the ROS-free replay driver (drop-in header, built-in velocity estimation, seeds pinned by
    environment) against the reference run on the same recorded stream: occupied-voxel files bit for bit, future-status
    files to the atomics' rounding, and the particle CSV (setParticleRecordFlag; columns flag,vx,vy,vz,px,py,pz,weight,voxel,
    dsp_dynamic.h:339-344) BYTE for byte, file name included."""
    name = "ref_default"
    if not refmap.available(name):
        pytest.skip("reference library not present")
    test_replay_tool_compiles()
    from dspmap_b200.streams import write_stream
    cfg = dm.CONFIGS[name]
    F, seed, csv_at = 6, 9, 4
    st = make_stream(cfg, seed=2, frames=F)
    path = str(tmp_path / "replay_stream.bin")
    write_stream(path, st)
    gdir, rdir = tmp_path / "gpu", tmp_path / "ref"
    gdir.mkdir()
    rdir.mkdir()
    env = dict(os.environ, DSPMAP_TABLE_SEED=str(seed), DSPMAP_UNIFORM_SEED=str(seed))
    out = subprocess.check_output([os.path.join(BUILD, "dspmap_replay"), path, "--out", "rp", "--future", "--csv-at-frame", str(csv_at)],
                                  text=True, env=env, cwd=str(gdir))
    assert out.count("occupied voxels") == F
    r = refmap.RefMap(name, seed=seed)
    V, T = cfg["nx"] * cfg["ny"] * cfg["nz"], len(cfg["future_times"])
    for k in range(F):
        if k == csv_at:
            r.set_particle_record_flag(-1, 1.0, str(rdir))
        assert r.update(st["points"][k], st["pos"][k], st["t"][k], st["quat"][k]) == 1
        if k == csv_at:
            r.set_particle_record_flag(0, 1.0, str(rdir))
        xyz, fut = r.occupancy(0.2)
        occ = np.fromfile(str(gdir / ("rp_frame%04d.occ" % k)), np.float32).reshape(-1, 3)
        assert occ.shape == xyz.shape and np.array_equal(occ.view(np.uint32), xyz.view(np.uint32)), "frame %d" % k
        gf = np.fromfile(str(gdir / ("rp_frame%04d.fut" % k)), np.float32).reshape(V, T)
        assert np.array_equal(gf != 0, fut != 0) and np.allclose(gf, fut, rtol=2e-6, atol=0), "future status, frame %d" % k
    gcsv = sorted(f for f in os.listdir(str(gdir)) if f.endswith(".csv"))
    rcsv = sorted(f for f in os.listdir(str(rdir)) if f.endswith(".csv"))
    assert gcsv == rcsv and len(gcsv) == 1, (gcsv, rcsv)
    a, b = open(str(gdir / gcsv[0]), "rb").read(), open(str(rdir / rcsv[0]), "rb").read()
    assert len(a) > 10000 and a == b
    # column order: flag, vx, vy, vz, px, py, pz, weight, voxel
    row = a.split(b"\n")[0].split(b",")
    assert len(row) == 9 and float(row[0]) in (1.0, 0.6) and float(row[3]) == 0.0 and 0 <= int(row[8]) < V


@pytest.mark.gpu
def test_particle_csv_equals_the_state_dump(tmp_path):
    """The CSV written by update() after setParticleRecordFlag holds dspmap_dump_particles' records in sweep order, in the
    reference's column order, formatted like the reference's ofstream (6 significant digits)."""
    from common import gpu_map, gpu_update
    cfg = dm.CONFIGS["tiny_dyn"]
    st = make_stream(cfg, seed=5, frames=3)
    g = gpu_map("tiny_dyn", seed=3)
    g.setParticleRecordFlag(-1, 1.0, folder=str(tmp_path))
    for k in range(3):
        assert gpu_update(g, st["points"][k], st["pos"][k], st["t"][k], st["quat"][k]) == 1
    ids, vals = g.particles()
    files = sorted(os.listdir(str(tmp_path)), key=lambda f: int(f.split("_")[3]))
    assert [f.split("_")[3] for f in files] == ["1", "2", "3"]
    rows = [l.split(",") for l in open(str(tmp_path / files[-1])).read().splitlines()]
    assert len(rows) == len(ids) > 50
    for (v, s), rec, row in zip(ids, vals, rows):
        assert int(row[8]) == v and row[:8] == ["%g" % x for x in rec]
    g.close()


