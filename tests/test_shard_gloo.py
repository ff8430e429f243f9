"""CPU, world_size 2 over gloo: the N>1 path of bench.py (unit sharding + max-over-ranks timing)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from fuif_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.units_for_rank(3, rank)
    strong = shard.split_batch(7, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, strong))
    ms, units = shard.reduce_step_time(10.0 + 5.0 * rank, len(mine))
    if rank == 0:
        out.put((gathered, ms, units))
    dist.destroy_process_group()


def test_units_are_disjoint_and_timing_is_max_over_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    gathered, ms, units = q.get(timeout=120)
    [p.join(60) for p in procs]
    weak = [u for g in gathered for u in g[0]]
    strong = [u for g in gathered for u in g[1]]
    assert sorted(weak) == list(range(6))          # 3 units per rank, no overlap
    assert sorted(strong) == list(range(7))        # the fixed batch is covered exactly once
    assert ms == 15.0                              # max over ranks, not rank 0's own time
    assert units == 6


def test_single_process_is_identity():
    sys.path.insert(0, ROOT)
    from fuif_b200 import shard
    assert shard.reduce_step_time(3.5, 4) == (3.5, 4)
    assert shard.units_for_rank(2, 3) == [6, 7]
