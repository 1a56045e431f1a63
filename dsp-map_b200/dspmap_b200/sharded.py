"""Voxel-subspace sharding of one DSP map over several GPUs (include/dspmap_b200.h "sharding").

The map's z layers are cut into `nranks` slabs; rank r owns the particles whose voxel lies in its slab.  One frame is six
library phases with five collectives between them:

    phase 0  -> all-to-all   boundary-crossing movers (fixed-size slabs, no host round trip)
    phase 1  -> all-gather   particles registered in FOV pyramids: 16-byte headers first (one small host sync to size the
                             payload), then only the used part of the slabs; every rank builds identical pyramid lists
    phase 2  -> all-reduce   C_z, computed for point pyramids i % nranks == rank   (zero-initialised, one writer: exact)
    phase 3  -> all-reduce   new weights of particle chunks c % nranks == rank (whoever owns the particles)
    phase 4  -> all-reduce   owners apply their weights; newborn split of the points whose voxel the rank owns
    phase 5

`NcclComm` runs them with torch.distributed (one process per GPU, NCCL over NVLink); `LocalCluster` drives N handles in
one process on one GPU and performs the same data movement with tensor copies — the library code path is identical, which
is how the sharded path is tested on a single-GPU box.  Results are bit-identical to an unsharded handle.
"""
import numpy as np
import torch

from . import DSPMap, derive

XREC, GREC, HDR = 12, 8, 4
GATHER_ROUND = 1024


def slab_plan(nz, nranks):
    """z-layer ranges [z0, z1) per rank: equal slabs of ceil(nz / nranks) layers (same arithmetic as the library)."""
    zpr = (nz + nranks - 1) // nranks
    return [(min(nz, r * zpr), min(nz, (r + 1) * zpr)) for r in range(nranks)]


def default_caps(cfg, nranks):
    d = derive(cfg)
    cap_live = min(d["V"] * d["S"], 8 << 20)
    cap_g = max(GATHER_ROUND, min(cap_live // nranks, 1 << 20) // GATHER_ROUND * GATHER_ROUND)
    cap_x = 8192   # boundary crossers per (source, destination) pair and frame
    return cap_x, cap_g


def gather_size(counts, cap_g):
    """Records per rank moved by the all-gather: the largest count, rounded up, at most the slab capacity."""
    g = (max(max(counts), 1) + GATHER_ROUND - 1) // GATHER_ROUND * GATHER_ROUND
    return min(g, cap_g)


class ShardBuffers:
    def __init__(self, nranks, cap_x, cap_g, max_points, P, obs_max, device):
        self.xs = HDR + cap_x * XREC
        self.gs = HDR + cap_g * GREC
        z = lambda n, dt=torch.float32: torch.zeros(n, dtype=dt, device=device)  # noqa: E731
        self.xsend, self.xrecv = z(nranks * self.xs), z(nranks * self.xs)
        self.gsend, self.grecv = z(self.gs), z(nranks * self.gs)
        self.hdr = z(nranks * HDR)
        self.czinv = z(P * obs_max + max_points)
        self.shared = z(max_points + nranks * cap_g)


class ShardedDSPMap:
    """One rank's handle of a sharded map."""

    def __init__(self, cfg, rank, nranks, device=0, seed=1, max_points=0, cap_x=None, cap_g=None, **kw):
        self.cfg, self.rank, self.nranks = cfg, rank, nranks
        dx, dg = default_caps(cfg, nranks)
        self.cap_x, self.cap_g = cap_x or dx, cap_g or dg
        self.max_points = max_points or 65536
        self.map = DSPMap(cfg, seed=seed, device=device, max_points=self.max_points, **kw)
        m = self.map
        self.buf = b = ShardBuffers(nranks, self.cap_x, self.cap_g, self.max_points, m.P, m.obs_max, torch.device("cuda", device))
        m.shard_config(rank, nranks, b.xsend.data_ptr(), b.xrecv.data_ptr(), self.cap_x, b.gsend.data_ptr(), b.grecv.data_ptr(),
                       self.cap_g, b.czinv.data_ptr(), b.shared.data_ptr())
        self.z0, self.z1 = slab_plan(cfg["nz"], nranks)[rank]
        self.v_lo, self.v_hi = self.z0 * cfg["nx"] * cfg["ny"], self.z1 * cfg["nx"] * cfg["ny"]
        self.P_obs = m.P * m.obs_max

    def bind_current_stream(self):
        """Run the map's frames on torch's current stream of its device (where torch.distributed enqueues the collectives)."""
        dev = int(self.map.config.device)
        h = torch.cuda.current_stream(dev).cuda_stream
        if h != getattr(self, "_bound_stream", None):
            if h != 0:
                self.map.set_stream(h)
            # (0 = torch's legacy default stream: the library keeps its own stream; sharded_update then synchronises around every collective)
            self._bound_stream = h

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
    """The collectives of a sharded frame over torch.distributed (backend nccl; gloo works too, for the CPU tests)."""

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
    """One frame on this rank (all ranks call it with the same cloud, pose and newborn input).

    The phases only enqueue kernels, and torch.distributed orders its collectives against torch's CURRENT stream: the map is
    (re)bound to that stream here, so a collective never reads a slab the phase in front of it has not written yet (the
    library's side branches are joined into that stream by events).  On torch's legacy default stream, which the library
    cannot adopt, every hand-over between a phase and a collective is a full synchronisation instead."""
    sm.bind_current_stream()
    legacy = sm._bound_stream == 0

    def coll(fn, *a):
        if legacy:
            sm.map.synchronize()
        fn(*a)
        if legacy:
            torch.cuda.current_stream(int(sm.map.config.device)).synchronize()

    b, N = sm.buf, sm.nranks
    rc = sm.phase(0, n, d_pts, pos, t, quat, d_tagged, n_tagged)
    if rc != 1:
        return rc
    coll(comm.all_to_all, b.xrecv, b.xsend, N)
    sm.phase(1, n, d_pts, pos, t, quat, d_tagged, n_tagged)
    coll(comm.all_gather, b.hdr, b.gsend[:HDR], N)                # counts of registered particles per rank
    counts = b.hdr.view(torch.int32)[::HDR].tolist()              # the one host synchronisation of a sharded frame
    g = gather_size(counts, sm.cap_g)
    sm.map.shard_gather_records(g)
    gs = HDR + g * GREC
    coll(comm.all_gather, b.grecv[:N * gs], b.gsend[:gs], N)
    sm.phase(2, n, d_pts, pos, t, quat, d_tagged, n_tagged)
    coll(comm.all_reduce_sum, b.czinv[:sm.P_obs + n])
    sm.phase(3, n, d_pts, pos, t, quat, d_tagged, n_tagged)
    coll(comm.all_reduce_sum, b.shared[sm.max_points:sm.max_points + sum(counts)])   # new weights
    sm.phase(4, n, d_pts, pos, t, quat, d_tagged, n_tagged)
    if n_tagged > 0:
        coll(comm.all_reduce_sum, b.shared[:n_tagged])                               # newborn split
    sm.phase(5, n, d_pts, pos, t, quat, d_tagged, n_tagged)
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

    def _each(self, k, args):
        for s in self.shards:
            s.phase(k, *args)
        self._sync()

    def update(self, pts, pos, t, quat, tagged):
        """pts [n,3], tagged [m,7]: host arrays; the same inputs go to every shard."""
        d_pts = torch.from_numpy(np.ascontiguousarray(pts, np.float32)).to(self.dev)
        tg = np.ascontiguousarray(tagged, np.float32).reshape(-1, 7)
        d_tag = torch.from_numpy(tg if len(tg) else np.zeros((1, 7), np.float32)).to(self.dev)
        n, nt, N, sh = len(pts), len(tg), self.nranks, self.shards
        args = (n, d_pts.data_ptr(), pos, t, quat, d_tag.data_ptr(), nt)
        rcs = [s.phase(0, *args) for s in sh]
        if any(rc != 1 for rc in rcs):
            return rcs[0]
        self._sync()
        xs = sh[0].buf.xs
        for r, dst in enumerate(sh):   # all-to-all: slab r of every sender goes to rank r
            for s, src in enumerate(sh):
                dst.buf.xrecv[s * xs:(s + 1) * xs].copy_(src.buf.xsend[r * xs:(r + 1) * xs])
        self._sync()
        self._each(1, args)
        counts = [int(s.buf.gsend[:1].view(torch.int32).item()) for s in sh]
        g = gather_size(counts, sh[0].cap_g)
        gs = HDR + g * GREC
        gathered = torch.cat([s.buf.gsend[:gs] for s in sh])   # all-gather of the used part of the slabs
        for s in sh:
            s.map.shard_gather_records(g)
            s.buf.grecv[:N * gs].copy_(gathered)
        self._sync()
        self._each(2, args)
        k = sh[0].P_obs + n
        tot = torch.stack([s.buf.czinv[:k] for s in sh]).sum(0)   # all-reduce: one non-zero contributor per element
        for s in sh:
            s.buf.czinv[:k].copy_(tot)
        self._sync()
        self._each(3, args)
        k0, k1 = sh[0].max_points, sh[0].max_points + sum(counts)
        tot = torch.stack([s.buf.shared[k0:k1] for s in sh]).sum(0)
        for s in sh:
            s.buf.shared[k0:k1].copy_(tot)
        self._sync()
        self._each(4, args)
        if nt:
            tot = torch.stack([s.buf.shared[:nt] for s in sh]).sum(0)
            for s in sh:
                s.buf.shared[:nt].copy_(tot)
        self._sync()
        self._each(5, args)
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
