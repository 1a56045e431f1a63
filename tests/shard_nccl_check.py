#!/usr/bin/env python3
"""Multi-GPU parity check (run with torchrun, one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/shard_nccl_check.py cfg2 6

Every rank runs its shard of the map (NCCL collectives between the phases); rank 0 additionally runs an unsharded map on
the same inputs and compares the gathered shards with it bit for bit after every frame."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dsp-map_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dspmap_b200 as dm  # noqa: E402
from dspmap_b200.sharded import NcclComm, ShardedDSPMap, sharded_update  # noqa: E402
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
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    cfg = dm.CONFIGS[name]
    st = make_stream(cfg, seed=8, frames=frames)
    est = dm.VelocityEstimator(cfg, seed=4, filter_res=0.1)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sm = ShardedDSPMap(cfg, rank, world, device=local, seed=4, max_points=cfg["points"])
    sm.map.set_stream(stream.cuda_stream)
    setters(sm.map)
    comm = NcclComm()
    one = None
    if rank == 0:
        one = dm.DSPMap(cfg, seed=4, device=local, max_points=cfg["points"])
        one.set_stream(stream.cuda_stream)
        setters(one)
    ok = True
    for f in range(frames):
        pts, pos, t, q = st["points"][f], st["pos"][f], st["t"][f], st["quat"][f]
        tc = est.estimate(pts, pos, t, q)
        d_pts = torch.from_numpy(pts).to(dev)
        d_tag = torch.from_numpy(tc).to(dev)
        sharded_update(sm, comm, len(pts), d_pts.data_ptr(), pos, t, q, d_tag.data_ptr(), len(tc))
        sm.map.synchronize()
        torch.cuda.synchronize()
        ids, vals = sm.map.particles()
        n = torch.tensor([len(ids)], device=dev)
        ns = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(ns, n)
        cap = int(max(x.item() for x in ns))
        pi = torch.zeros((cap, 2), dtype=torch.int32, device=dev)
        pv = torch.zeros((cap, 8), dtype=torch.float32, device=dev)
        pi[:len(ids)] = torch.from_numpy(ids).to(dev)
        pv[:len(ids)] = torch.from_numpy(vals).to(dev)
        gi = [torch.zeros_like(pi) for _ in range(world)]
        gv = [torch.zeros_like(pv) for _ in range(world)]
        dist.all_gather(gi, pi)
        dist.all_gather(gv, pv)
        if rank == 0:
            one.update(len(pts), 3, pts, float(pos[0]), float(pos[1]), float(pos[2]), float(t), float(q[0]), float(q[1]), float(q[2]), float(q[3]), tagged=tc)
            oid, oval = one.particles()
            sid = np.concatenate([gi[r][:int(ns[r].item())].cpu().numpy() for r in range(world)])
            sval = np.concatenate([gv[r][:int(ns[r].item())].cpu().numpy() for r in range(world)])
            same = sid.shape == oid.shape and np.array_equal(sid, oid) and np.array_equal(sval.view(np.uint32), oval.view(np.uint32))
            print("frame", f, "particles", len(oid), "per rank", [int(x.item()) for x in ns], "bit-identical" if same else "MISMATCH", flush=True)
            ok = ok and same
    if rank == 0:
        print("SHARDED == SINGLE GPU" if ok else "FAILED")
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
