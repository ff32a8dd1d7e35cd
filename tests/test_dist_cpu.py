"""Host-side logic of the multi-GPU path (lisa_b200/dist.py) on CPU: subframe partitioning and the one
reduce step, world_size 2 over gloo.  The GPU-side equivalence (N ranks == 1 rank over the same subframe
set) is in test_gpu_multi.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_partition_covers_disjointly():
    from lisa_b200.dist import subframes_for_rank
    for first in (0, 5):
        for count in (1, 2, 7, 8, 64, 125, 4096):
            for world in (1, 2, 4, 8):
                parts = [subframes_for_rank(first, count, r, world) for r in range(world)]
                got = sorted(f for (f0, n) in parts for f in range(f0, f0 + n))
                assert got == list(range(first, first + count))
                sizes = [n for _, n in parts]
                assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lisa_b200.dist import reduce_accum, subframes_for_rank
    # each rank "renders" its subframes: accumulator = (sum of subframe means, subframe count) per pixel
    npix, first, count = 37, 3, 9
    f0, n = subframes_for_rank(first, count, rank, world)
    acc = torch.zeros(npix * 4, dtype=torch.float32)
    a = acc.view(npix, 4)
    for f in range(f0, f0 + n):
        g = torch.Generator().manual_seed(f)
        a[:, :3] += torch.rand(npix, 3, generator=g)
        a[:, 3] += 1
    reduce_accum(acc, 0)
    if rank == 0:
        q.put(acc.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduce_equals_single_rank():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120).reshape(-1, 4)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    exp = np.zeros((37, 4), np.float32)
    for f in range(3, 12):
        g = torch.Generator().manual_seed(f)
        exp[:, :3] += torch.rand(37, 3, generator=g).numpy()
        exp[:, 3] += 1
    np.testing.assert_allclose(got, exp, rtol=1e-6)
    assert (got[:, 3] == 9).all()
