#!/usr/bin/env python3
"""Multi-GPU parity check of the LIBRARY-ORCHESTRATED sharded map (run with torchrun, one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/shard_nccl_cpp_check.py cfg2 6

Every rank drives its shard with dspmap_shard_update / dspmap_shard_get_occupancy (the phases and the NCCL collectives are
issued by the C++ library; torch.distributed only carries the 128-byte NCCL id and the verdict).  Rank 0 additionally runs
an unsharded map on the same inputs and compares: every rank's occupied-voxel list and future grid, and the union of the
shards' particles, after every frame.  Prints one JSON line; exit code 0 only if everything is bit-identical."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dsp-map_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dspmap_b200 as dm  # noqa: E402
from dspmap_b200.streams import make_stream  # noqa: E402


def setters(g):
    g.setPredictionVariance(0.05, 0.05)
    g.setObservationStdDev(0.1)
    g.setNewBornParticleNumberofEachPoint(20)
    g.setNewBornParticleWeight(1e-4)


def main():
    name, frames = sys.argv[1], int(sys.argv[2])
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("gloo")   # only host-side plumbing goes through torch: the map's collectives are the library's own
    cfg = dm.CONFIGS[name]
    st = make_stream(cfg, seed=8, frames=frames)
    est = dm.VelocityEstimator(cfg, seed=4, filter_res=0.1)
    box = [dm.shard_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    m = dm.DSPMap(cfg, seed=4, device=local, max_points=cfg["points"])
    setters(m)
    m.shard_init(rank, world, box[0])
    one = None
    if rank == 0:
        one = dm.DSPMap(cfg, seed=4, device=local, max_points=cfg["points"])
        setters(one)
    d_xyz = torch.zeros((m.V, 3), dtype=torch.float32, device=dev)
    d_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    d_fut = torch.zeros((m.V, max(m.T, 1)), dtype=torch.float32, device=dev)
    bad = []
    for f in range(frames):
        pts, pos, t, q = st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]
        tc = est.estimate(pts, pos, t, q)
        d_pts = torch.from_numpy(np.ascontiguousarray(pts, np.float32)).to(dev)
        d_tag = torch.from_numpy(np.ascontiguousarray(tc if len(tc) else np.zeros((1, 7), np.float32), np.float32)).to(dev)
        torch.cuda.synchronize()
        assert m.shard_update(len(pts), d_pts.data_ptr(), pos, t, q, d_tag.data_ptr(), len(tc)) == 1
        m.shard_get_occupancy(0.2, d_xyz.data_ptr(), m.V, d_cnt.data_ptr(), d_fut.data_ptr())
        m.synchronize()
        n = int(d_cnt.item())
        mine = (n, d_xyz[:n].cpu().numpy(), d_fut.cpu().numpy(), m.particles(), m.counters()["overflow"])
        gathered = [None] * world
        dist.gather_object(mine, gathered if rank == 0 else None, dst=0)
        if rank == 0:
            one.update(len(pts), 3, pts, float(pos[0]), float(pos[1]), float(pos[2]), float(t), float(q[0]), float(q[1]), float(q[2]), float(q[3]), tagged=tc)
            ids, vals = one.particles()
            n1, xyz1, fut1 = one.getOccupancyMapWithFutureStatus(0.2)
            sid = np.concatenate([g[3][0] for g in gathered])
            sva = np.concatenate([g[3][1] for g in gathered])
            order = np.argsort(sid[:, 0].astype(np.int64) * 128 + sid[:, 1], kind="stable")
            if not (sid[order].shape == ids.shape and np.array_equal(sid[order], ids) and np.array_equal(sva[order].view(np.uint32), vals.view(np.uint32))):
                bad.append("frame %d: particles of the shards differ from the single map" % f)
            for r, g in enumerate(gathered):
                if g[4]:
                    bad.append("frame %d rank %d: capacity code %d" % (f, r, g[4]))
                if g[0] != n1 or not np.array_equal(g[1].view(np.uint32), xyz1.view(np.uint32)):
                    bad.append("frame %d rank %d: occupied list differs (%d vs %d)" % (f, r, g[0], n1))
                if not (np.array_equal(g[2] != 0, fut1 != 0) and np.allclose(g[2], fut1, rtol=4e-6, atol=0)):
                    bad.append("frame %d rank %d: future grid differs" % (f, r))
    verdict = [bad]
    dist.broadcast_object_list(verdict, src=0)
    if rank == 0:
        print(json.dumps({"check": "library-orchestrated sharded map vs single GPU", "config": name, "frames": frames, "ranks": world,
                          "bit_identical": not bad, "mismatches": bad[:6], "shard_info": m.shard_info()}))
    m.close()
    if one is not None:
        one.close()
    dist.destroy_process_group()
    return 1 if verdict[0] else 0


if __name__ == "__main__":
    sys.exit(main())
