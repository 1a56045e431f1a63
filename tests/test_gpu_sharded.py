"""GPU tests of voxel-subspace sharding: N shards of one map (LocalCluster: the library's sharded code path, collectives
done as tensor copies on one GPU) must reproduce an unsharded map bit for bit, frame after frame."""
import numpy as np
import pytest

import dspmap_b200 as dm
from common import SET, gpu_map, gpu_update, make_stream
from dspmap_b200.sharded import LocalCluster
from parity import same

pytestmark = pytest.mark.gpu


def setters(g):
    g.setPredictionVariance(SET["p_std"], SET["v_std"])
    g.setObservationStdDev(SET["ob_std"])
    g.setNewBornParticleNumberofEachPoint(SET["newborn_num"])
    g.setNewBornParticleWeight(SET["newborn_weight"])


@pytest.mark.parametrize("name,nranks,frames", [("tiny_dyn", 2, 10), ("tiny_dyn", 3, 8), ("tiny_static", 2, 6), ("cfg1", 2, 5), ("cfg2", 4, 4), ("cfg5", 8, 2)])
def test_sharded_equals_single_gpu(name, nranks, frames):
    cfg = dm.CONFIGS[name]
    st = make_stream(cfg, seed=8, frames=frames)
    est = dm.VelocityEstimator(cfg, seed=4, filter_res=0.1)
    one = gpu_map(name, seed=4, max_points=cfg["points"])
    cl = LocalCluster(cfg, nranks, seed=4, max_points=cfg["points"], setters=setters)
    crossed = 0
    for f in range(frames):
        pts, pos, t, q = st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]
        tc = est.estimate(pts, pos, t, q)
        assert gpu_update(one, pts, pos, t, q, tagged=tc) == 1
        assert cl.update(pts, pos, t, q, tc) == 1
        ids, vals = one.particles()
        sids, svals = cl.particles()
        assert same(ids, sids), "frame %d: particle (voxel, slot) sets" % f
        assert same(vals, svals), "frame %d: particle fields" % f
        vo, svo = one.voxel_objects(), cl.voxel_objects()
        assert same(vo[:, :4], svo[:, :4]), "frame %d: occupancy / mean velocity" % f
        assert np.array_equal(vo[:, 4:] != 0, svo[:, 4:] != 0) and np.allclose(vo[:, 4:], svo[:, 4:], rtol=4e-6, atol=0)
        c1 = one.cursors()
        for c in cl.cursors():
            assert np.array_equal(c[:3], c1[:3]), "frame %d: cursors" % f
        cs = cl.counters()
        assert sum(c["n_out"] for c in cs) == one.counters()["n_out"] and all(c["overflow"] == 0 for c in cs)
        # every shard only holds particles of its own voxel subspace
        for s in cl.shards:
            pid, _ = s.map.particles()
            assert len(pid) == 0 or (pid[:, 0].min() >= s.v_lo and pid[:, 0].max() < s.v_hi)
        if f % 2 == 1:
            n, xyz, fut = one.getOccupancyMapWithFutureStatus(0.2)
            sxyz, sfut = cl.occupancy(0.2)
            assert same(xyz, sxyz) and np.allclose(fut, sfut, rtol=4e-6, atol=0)
    one.close()
    cl.close()


@pytest.mark.parametrize("name,nranks,frames", [("tiny_dyn", 2, 8), ("tiny_dyn", 3, 6), ("cfg2", 4, 5), ("cfg2", 8, 3)])
def test_library_orchestrated_sharded_map_equals_single_gpu(name, nranks, frames):
    """The C++ orchestrator (dspmap_shard_update_local: six phases + collectives issued by the library, no host synchronisation
    inside a frame, the gather of frame k sized from frame k-2) against an unsharded map: particle state after every frame,
    and the sharded reader's occupied list / future grid on EVERY rank."""
    cfg = dm.CONFIGS[name]
    st = make_stream(cfg, seed=8, frames=frames)
    est = dm.VelocityEstimator(cfg, seed=4, filter_res=0.1)
    one = gpu_map(name, seed=4, max_points=cfg["points"])
    cl = dm.LocalShardedMap(cfg, nranks, seed=4, max_points=cfg["points"], setters=setters)
    for f in range(frames):
        pts, pos, t, q = st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]
        tc = est.estimate(pts, pos, t, q)
        assert gpu_update(one, pts, pos, t, q, tagged=tc) == 1
        assert cl.update(pts, pos, t, q, tc) == 1
        ids, vals = one.particles()
        sids, svals = cl.particles()
        assert same(ids, sids), "frame %d: particle (voxel, slot) sets" % f
        assert same(vals, svals), "frame %d: particle fields" % f
        vo, svo = one.voxel_objects(), cl.voxel_objects()
        assert same(vo[:, :4], svo[:, :4]), "frame %d: occupancy / mean velocity" % f
        assert all(m.counters()["overflow"] == 0 for m in cl.maps)
        n, xyz, fut = one.getOccupancyMapWithFutureStatus(0.2)
        for r, (k, sxyz, sfut) in enumerate(cl.occupancy(0.2)):
            assert k == n and same(xyz, sxyz), "frame %d rank %d: occupied-voxel list" % (f, r)
            assert np.array_equal(fut != 0, sfut != 0) and np.allclose(fut, sfut, rtol=4e-6, atol=0), "frame %d rank %d: future grid" % (f, r)
    info = cl.maps[0].shard_info()
    assert info["frames"] == frames and (frames < 3 or info["gather_records"] < info["cap_g"])   # the gather moves what the exchange headers bound, not the slab capacity
    one.close()
    cl.close()


def test_library_orchestrated_sharded_map_over_nccl():
    """The same orchestrator over NCCL, one process per GPU (tests/shard_nccl_cpp_check.py under torchrun): needs a box with at
    least two GPUs, skipped otherwise."""
    import json
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU visible: the NCCL back end needs two (the local back end above runs the same orchestration code)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)), "--master-addr", "127.0.0.1",
                        "--master-port", "29513", os.path.join(root, "tests", "shard_nccl_cpp_check.py"), "cfg2", "5"],
                       capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and line and json.loads(line[-1])["bit_identical"], r.stdout[-2000:] + r.stderr[-2000:]
