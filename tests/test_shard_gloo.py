"""world_size-2 gloo test of the multi-GPU host logic (no GPU needed): the library's contiguous chunk ranges
(CSCB200_ChunkRange, the rule CSCB200_EncodeChunks fans out by and bench.py partitions by), nothing lost or
duplicated, ordered concatenation on rank 0, max-over-ranks timing reduction."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from csc_b200 import api as shard  # noqa: E402  (the product's own partition rule, CSCB200_ChunkRange)


def _owner(c, n, w):
    return c * w // n


def test_chunk_ranges_partition_exactly():
    for n in (0, 1, 2, 7, 8, 9, 1000, 4096):
        for w in (1, 2, 4, 8):
            got = []
            for r in range(w):
                lo, hi = shard.chunk_range(n, r, w)
                assert 0 <= lo <= hi <= n
                got += list(range(lo, hi))
                assert all(_owner(c, n, w) == r for c in range(lo, hi))
            assert got == list(range(n))


def _worker(rank, world, port, n_chunks, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.chunk_range(n_chunks, rank, world)
    # stand-in for the per-rank compressed sizes + elapsed time
    sizes = torch.zeros(n_chunks, dtype=torch.int64)
    sizes[lo:hi] = torch.arange(lo, hi) * 3 + 1
    dist.all_reduce(sizes, op=dist.ReduceOp.SUM)
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        q.put((sizes.tolist(), float(t.item())))
    dist.destroy_process_group()


def test_two_rank_gloo_reduction():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 37
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    sizes, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sizes == [c * 3 + 1 for c in range(n)]
    assert tmax == 11.0
