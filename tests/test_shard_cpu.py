"""Multi-GPU protocol of the batched path, on CPU with gloo (world_size 2): the pair index is
partitioned contiguously, there is no collective in the solve, and the only communication is the
reporting reduction bench.py does (sum of work, max of time)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opticalflow2d_b200 import shard_pairs


def test_shards_partition_the_batch():
    for total in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_pairs(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_native_shard_ranges_match_the_python_partition():
    """of2d_shard_range (include/of2d_host.h) is the partition of2d_batch_create_multi uses inside one process: it must be
    the one the one-process-per-GPU launch uses (shard_pairs).  Pure host arithmetic: no device needed."""
    import ctypes as C
    import opticalflow2d_b200 as of
    lib = of.host(32)
    for total in (0, 1, 7, 512, 4096, 4097):
        for world in (1, 2, 3, 8):
            for r in range(world):
                lo, hi = C.c_int(), C.c_int()
                assert lib.of2d_shard_range(total, world, r, C.byref(lo), C.byref(hi)) == 0
                assert (lo.value, hi.value) == shard_pairs(total, world, r)
    assert lib.of2d_shard_range(8, 2, 2, None, None) != 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_pairs(total, world, rank)
    pairs = torch.tensor([float(hi - lo)], dtype=torch.float64)
    seconds = torch.tensor([1.0 + 0.5 * rank], dtype=torch.float64)      # stand-in for the device time of the shard
    dist.all_reduce(pairs, op=dist.ReduceOp.SUM)
    dist.all_reduce(seconds, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put((float(pairs), float(seconds)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reporting_reduction():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 4096, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    pairs, seconds = out.get(timeout=10)
    assert pairs == 4096.0 and seconds == 1.5
