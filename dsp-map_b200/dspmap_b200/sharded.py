"""Voxel-subspace sharding of one DSP map over several GPUs (include/dspmap_b200.h "sharding").

The map's z layers are cut into `nranks` slabs; rank r owns the particles whose voxel lies in its slab.  One frame is four
library phases with three collectives between them:

    phase 0  -> all-to-all   boundary-crossing movers (fixed-size slabs, no host round trip)
    phase 1  -> all-gather   particles registered in FOV pyramids (every rank builds identical global pyramid lists)
    phase 2  -> all-reduce   per-point newborn split (computed by the owner of the point's voxel)
    phase 3

`NcclComm` runs them with torch.distributed (one process per GPU, NCCL over NVLink); `LocalCluster` drives N handles in
one process on one GPU and performs the same data movement with tensor copies — the library code path is identical, which
is how the sharded path is tested on a single-GPU box.  Results are bit-identical to an unsharded handle.
"""
import numpy as np
import torch

from . import DSPMap, derive

XREC, GREC, HDR = 12, 8, 4


def slab_plan(nz, nranks):
    """z-layer ranges [z0, z1) per rank: equal slabs of ceil(nz / nranks) layers (same arithmetic as the library)."""
    zpr = (nz + nranks - 1) // nranks
    return [(min(nz, r * zpr), min(nz, (r + 1) * zpr)) for r in range(nranks)]


def default_caps(cfg, nranks):
    d = derive(cfg)
    cap_live = min(d["V"] * d["S"], 8 << 20)
    cap_g = max(1024, min(cap_live // nranks, 2 << 20))
    cap_x = max(1024, min(cap_g // 4, 1 << 18))
    return cap_x, cap_g


class ShardBuffers:
    def __init__(self, nranks, cap_x, cap_g, max_points, device):
        self.xs = HDR + cap_x * XREC
        self.gs = HDR + cap_g * GREC
        self.xsend = torch.zeros(nranks * self.xs, dtype=torch.float32, device=device)
        self.xrecv = torch.zeros(nranks * self.xs, dtype=torch.float32, device=device)
        self.gsend = torch.zeros(self.gs, dtype=torch.float32, device=device)
        self.grecv = torch.zeros(nranks * self.gs, dtype=torch.float32, device=device)
        self.nst = torch.zeros(max_points, dtype=torch.int32, device=device)


class ShardedDSPMap:
    """One rank's handle of a sharded map."""

    def __init__(self, cfg, rank, nranks, device=0, seed=1, max_points=0, cap_x=None, cap_g=None, **kw):
        self.cfg, self.rank, self.nranks = cfg, rank, nranks
        dx, dg = default_caps(cfg, nranks)
        self.cap_x, self.cap_g = cap_x or dx, cap_g or dg
        self.map = DSPMap(cfg, seed=seed, device=device, max_points=max_points, **kw)
        mp = max_points or 65536
        self.buf = ShardBuffers(nranks, self.cap_x, self.cap_g, mp, torch.device("cuda", device))
        self.map.shard_config(rank, nranks, self.buf.xsend.data_ptr(), self.buf.xrecv.data_ptr(), self.cap_x,
                              self.buf.gsend.data_ptr(), self.buf.grecv.data_ptr(), self.cap_g, self.buf.nst.data_ptr())
        self.z0, self.z1 = slab_plan(cfg["nz"], nranks)[rank]
        self.v_lo, self.v_hi = self.z0 * cfg["nx"] * cfg["ny"], self.z1 * cfg["nx"] * cfg["ny"]

    def phase(self, k, n, d_pts, pos, t, quat, d_tagged, n_tagged):
        return self.map.shard_phase(k, n, d_pts, pos, t, quat, d_tagged, n_tagged)

    def load_particles(self, ids, vals):
        """Keeps only the particles of this rank's voxel subspace."""
        ids = np.asarray(ids)
        own = (ids[:, 0] >= self.v_lo) & (ids[:, 0] < self.v_hi) if len(ids) else np.zeros(0, bool)
        self.map.load_particles(ids[own], np.asarray(vals)[own])

    def close(self):
        self.map.close()


class NcclComm:
    """The three collectives of a sharded frame over torch.distributed (backend nccl; gloo works too for tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.gloo = dist.get_backend(group) == "gloo"

    def all_to_all(self, recv, send, nranks):
        if self.gloo:  # gloo has no all_to_all: gather every rank's send buffer and pick the slab addressed to this rank
            parts = [torch.empty_like(send) for _ in range(nranks)]
            self.dist.all_gather(parts, send, group=self.group)
            me, n = self.dist.get_rank(self.group), send.numel() // nranks
            for s in range(nranks):
                recv[s * n:(s + 1) * n] = parts[s][me * n:(me + 1) * n]
        else:
            self.dist.all_to_all_single(recv, send, group=self.group)

    def all_gather(self, recv, send, nranks):
        if self.gloo:
            parts = [torch.empty_like(send) for _ in range(nranks)]
            self.dist.all_gather(parts, send, group=self.group)
            recv.copy_(torch.cat(parts))
        else:
            self.dist.all_gather_into_tensor(recv, send, group=self.group)

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)


def sharded_update(sm, comm, n, d_pts, pos, t, quat, d_tagged, n_tagged):
    """One frame on this rank (all ranks call it with the same cloud, pose and newborn input)."""
    rc = sm.phase(0, n, d_pts, pos, t, quat, d_tagged, n_tagged)
    if rc != 1:
        return rc
    comm.all_to_all(sm.buf.xrecv, sm.buf.xsend, sm.nranks)
    sm.phase(1, n, d_pts, pos, t, quat, d_tagged, n_tagged)
    comm.all_gather(sm.buf.grecv, sm.buf.gsend, sm.nranks)
    sm.phase(2, n, d_pts, pos, t, quat, d_tagged, n_tagged)
    if n_tagged > 0:
        comm.all_reduce_sum(sm.buf.nst[:n_tagged])
    sm.phase(3, n, d_pts, pos, t, quat, d_tagged, n_tagged)
    return 1


class LocalCluster:
    """N shards of one map in ONE process on ONE GPU; the collectives are tensor copies (SURVEY.md §4 tier 5)."""

    def __init__(self, cfg, nranks, device=0, seed=1, max_points=0, setters=None, **kw):
        self.cfg, self.nranks = cfg, nranks
        self.shards = [ShardedDSPMap(cfg, r, nranks, device=device, seed=seed, max_points=max_points, **kw) for r in range(nranks)]
        stream = torch.cuda.current_stream(device)
        for s in self.shards:
            if stream.cuda_stream != 0:
                s.map.set_stream(stream.cuda_stream)
            if setters:
                setters(s.map)
        self.dev = torch.device("cuda", device)

    def _sync(self):
        for s in self.shards:
            s.map.synchronize()
        torch.cuda.synchronize()

    def update(self, pts, pos, t, quat, tagged):
        """pts [n,3], tagged [m,7]: host arrays; the same inputs go to every shard."""
        d_pts = torch.from_numpy(np.ascontiguousarray(pts, np.float32)).to(self.dev)
        tg = np.ascontiguousarray(tagged, np.float32).reshape(-1, 7)
        d_tag = torch.from_numpy(tg if len(tg) else np.zeros((1, 7), np.float32)).to(self.dev)
        n, nt, N = len(pts), len(tg), self.nranks
        rcs = [s.phase(0, n, d_pts.data_ptr(), pos, t, quat, d_tag.data_ptr(), nt) for s in self.shards]
        if any(rc != 1 for rc in rcs):
            return rcs[0]
        self._sync()
        xs = self.shards[0].buf.xs
        for r, dst in enumerate(self.shards):   # all-to-all: slab r of every sender goes to rank r
            for s, src in enumerate(self.shards):
                dst.buf.xrecv[s * xs:(s + 1) * xs].copy_(src.buf.xsend[r * xs:(r + 1) * xs])
        self._sync()
        for s in self.shards:
            s.phase(1, n, d_pts.data_ptr(), pos, t, quat, d_tag.data_ptr(), nt)
        self._sync()
        g = torch.cat([s.buf.gsend for s in self.shards])   # all-gather
        for s in self.shards:
            s.buf.grecv.copy_(g)
        self._sync()
        for s in self.shards:
            s.phase(2, n, d_pts.data_ptr(), pos, t, quat, d_tag.data_ptr(), nt)
        self._sync()
        if nt:
            tot = torch.stack([s.buf.nst[:nt] for s in self.shards]).sum(0, dtype=torch.int32)   # all-reduce
            for s in self.shards:
                s.buf.nst[:nt].copy_(tot)
        self._sync()
        for s in self.shards:
            s.phase(3, n, d_pts.data_ptr(), pos, t, quat, d_tag.data_ptr(), nt)
        self._sync()
        return 1

    # ---- assembled state (for parity checks against an unsharded map) ------------------------------------------
    def particles(self):
        parts = [s.map.particles() for s in self.shards]
        ids = np.concatenate([p[0] for p in parts])
        vals = np.concatenate([p[1] for p in parts])
        order = np.argsort(ids[:, 0].astype(np.int64) * 128 + ids[:, 1], kind="stable")
        return ids[order], vals[order]

    def voxel_objects(self):
        out = None
        for s in self.shards:
            vo = s.map.voxel_objects()
            if out is None:
                out = np.zeros_like(vo)
            out[s.v_lo:s.v_hi, :4] = vo[s.v_lo:s.v_hi, :4]
            out[:, 4:] += vo[:, 4:]
        return out

    def cursors(self):
        return [s.map.cursors() for s in self.shards]

    def counters(self):
        return [s.map.counters() for s in self.shards]

    def occupancy(self, threshold):
        xyz, fut = [], None
        for s in self.shards:
            n, x, f = s.map.getOccupancyMapWithFutureStatus(threshold)
            xyz.append(x)
            fut = f if fut is None else fut + f
        return np.concatenate(xyz), fut

    def close(self):
        for s in self.shards:
            s.close()
