"""CPU tests of the multi-GPU host logic: the slab plan matches the library's owner arithmetic, and the three collectives
of a sharded frame route fixed-size slabs correctly (torch.distributed, gloo, world_size 2 — no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dspmap_b200 import CONFIGS
from dspmap_b200.sharded import GREC, HDR, XREC, NcclComm, default_caps, gather_size, slab_plan


def test_slab_plan_covers_the_map_once():
    for nz in (10, 40, 80, 7):
        for n in (1, 2, 3, 4, 8):
            plan = slab_plan(nz, n)
            assert plan[0][0] == 0 and plan[-1][1] == nz or plan[-1][1] == plan[-1][0]
            covered = sum(b - a for a, b in plan)
            assert covered == nz and all(a <= b for a, b in plan)
            zpr = (nz + n - 1) // n
            for z in range(nz):  # the library's owner rule (dsp_owner): min(n-1, z / ceil(nz/n))
                r = min(n - 1, z // zpr)
                assert plan[r][0] <= z < plan[r][1]
    cx, cg = default_caps(CONFIGS["cfg5"], 8)
    assert cx >= 1024 and cg * 8 <= 8 << 20 and cg % 1024 == 0
    assert gather_size([5, 3000, 70], cg) == 3072 and gather_size([0, 0], cg) == 1024 and gather_size([10 ** 9], cg) == cg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = NcclComm()
    cap_x, cap_g = 5, 6
    xs, gs = HDR + cap_x * XREC, HDR + cap_g * GREC
    # all-to-all: slab d of rank s carries the value 100*s + d in its header and 1000*s + d in its records
    send = torch.zeros(world * xs)
    for d in range(world):
        send[d * xs] = 100 * rank + d
        send[d * xs + HDR:(d + 1) * xs] = 1000 * rank + d
    recv = torch.zeros(world * xs)
    comm.all_to_all(recv, send, world)
    ok = all(recv[s * xs] == 100 * s + rank and torch.all(recv[s * xs + HDR:(s + 1) * xs] == 1000 * s + rank) for s in range(world))
    # all-gather
    g = torch.full((gs,), float(rank + 1))
    gr = torch.zeros(world * gs)
    comm.all_gather(gr, g, world)
    ok = ok and all(torch.all(gr[s * gs:(s + 1) * gs] == s + 1) for s in range(world))
    # all-reduce of a zero-initialised buffer with exactly one writer per element reproduces the writers' bits
    vals = torch.tensor([0.1 * (i + 1) for i in range(10)], dtype=torch.float32)
    nst = torch.where(torch.arange(10) % world == rank, vals, torch.zeros(10))
    comm.all_reduce_sum(nst)
    ok = ok and bool(torch.equal(nst, vals))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_collectives_route_slabs_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
