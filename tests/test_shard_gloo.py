"""Multi-process (gloo, world_size 2, CPU) test of the N>1 host logic: one process per GPU, images sharded by rank with
shard_by_size, NO data-path collective; torch.distributed is used exactly as bench.py uses it - a barrier and a MAX
all-reduce of the timings.  The decode itself needs a GPU, so each rank here stands in for it by hashing its shard's
bytes; what is checked is that the shards are disjoint, cover the list, are balanced by compressed bytes, and that
the per-rank results combine to the same aggregate the single-process run gives."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sizes, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from pim_jpeg_decoder_b200.decoder import shard_by_size
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_by_size(sizes, world)[rank]
    dist.barrier()
    # stand-in for the per-rank decode: units processed + a fake elapsed time that differs per rank
    units = torch.tensor([float(sum(sizes[i] for i in mine))], dtype=torch.float64)
    elapsed = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)          # bench.py: max over ranks
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)                  # harness-only (the test's check), not a data-path collective
    total = units.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put((gathered, float(elapsed.item()), float(total.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_ranks_shard_disjoint_and_balanced():
    import torch.multiprocessing as mp
    rng = np.random.default_rng(5)
    sizes = [int(x) for x in rng.integers(5_000, 3_000_000, size=257)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sizes, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, elapsed, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a, b = gathered
    assert not set(a) & set(b) and sorted(a + b) == list(range(len(sizes)))
    assert elapsed == 2.0                                    # MAX over ranks, not rank 0's own
    assert total == float(sum(sizes))                        # whole-job aggregate = all ranks' units
    ba, bb = sum(sizes[i] for i in a), sum(sizes[i] for i in b)
    assert abs(ba - bb) <= max(sizes)                        # round-robin over the size-sorted list balances bytes


def test_shard_by_size_properties():
    sys.path.insert(0, ROOT)
    from pim_jpeg_decoder_b200.decoder import shard_by_size
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 8, 100):
            sizes = [(i * 7919) % 1000 for i in range(n)]
            shards = shard_by_size(sizes, world)
            assert len(shards) == world and sorted(i for s in shards for i in s) == list(range(n))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
            for s in shards:                                 # ascending by size inside a shard, like the reference's sort
                assert [sizes[i] for i in s] == sorted(sizes[i] for i in s)


def test_lpt_stream_dealing_properties():
    """bench.py --workload config5 --stream N: ONE stream (a pool cycled) dealt over the ranks by compressed size,
    longest first (LPT): shards are disjoint, cover the stream, and their byte loads differ by less than one image."""
    sys.path.insert(0, ROOT)
    from pim_jpeg_decoder_b200.decoder import lpt_shards
    rng = np.random.default_rng(5)
    pool = [int(x) for x in rng.choice([75_000, 120_000, 20_000, 300_000, 800_000, 3_200_000], size=64, p=[.55, .15, .1, .1, .07, .03])]
    for world in (1, 2, 4, 8):
        for n in (0, 1, 63, 4096):
            costs = [pool[i % len(pool)] for i in range(n)]
            shards = lpt_shards(costs, world)
            assert len(shards) == world and sorted(i for s in shards for i in s) == list(range(n))
            loads = [sum(costs[i] for i in s) for s in shards]
            if n >= world:
                assert max(loads) - min(loads) <= max(costs)
            for s in shards:
                assert s == sorted(s)
