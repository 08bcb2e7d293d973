"""Region sharding (N > 1 path): partition logic and the max/sum reductions over a world_size-2 gloo group."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rabbitvar_b200.shard import bai_tile_weights, concat_in_order, contiguous_blocks, reduce_step_metrics  # noqa: E402


def test_blocks_cover_in_order_and_balance():
    w = [10] * 100
    for parts in (1, 2, 3, 4, 8):
        b = contiguous_blocks(w, parts)
        assert b[0][0] == 0 and b[-1][1] == 100
        assert all(b[i][1] == b[i + 1][0] for i in range(parts - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1
    # skewed weights: a heavy head must not starve later blocks
    w = [1000] + [1] * 9
    b = contiguous_blocks(w, 4)
    assert b[0] == (0, 1) and all(hi > lo for lo, hi in b) and b[-1][1] == 10
    assert contiguous_blocks([], 3) == [(0, 0)] * 3
    assert contiguous_blocks([5, 5], 4)[:2] == [(0, 1), (1, 2)]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tiles = list(range(20))
    weights = [3 if t < 5 else 1 for t in tiles]
    lo, hi = contiguous_blocks(weights, world)[rank]
    my_units = sum(weights[lo:hi]) * 1000
    my_ms = 10.0 + rank  # rank 1 is slower: the job time is the max
    t, u = reduce_step_metrics(my_ms, my_units, dist)
    text = concat_in_order([f"tile{t}\n" for t in tiles[lo:hi]], dist)
    dist.barrier()
    if rank == 0:
        q.put((t, u, text, (lo, hi)))
    dist.destroy_process_group()


def test_two_rank_gloo_reduction_and_ordered_concat():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    t, u, text, blk0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t == 11.0                      # max over ranks
    assert u == (5 * 3 + 15) * 1000      # sum over ranks == whole job
    assert text == "".join(f"tile{i}\n" for i in range(20))  # rank order == tile order
    assert blk0[0] == 0


def test_bai_weights_follow_the_read_density(built):
    """bai_tile_weights: compressed bytes per tile from the BAI linear index — the weights bench.py cuts the tile list
    with (contiguous_blocks).  A region without reads weighs (almost) nothing, equal tiles of a uniform BAM weigh alike."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    d = cases.generate("c2_somatic_bed_k0")
    tiles = [(int(l.split()[1]), int(l.split()[2])) for l in open(os.path.join(d, "tiles.bed"))]
    w = bai_tile_weights(os.path.join(d, "T.bam.bai"), 0, tiles + [(100, 900)])
    assert len(w) == len(tiles) + 1 and all(x > 0 for x in w)
    assert w[-1] < 0.2 * max(w[:-1])                       # 800 bp against 10 kb tiles
    blocks = contiguous_blocks(w[:-1], 2)
    assert blocks[0][1] == blocks[1][0] and abs((blocks[0][1] - blocks[0][0]) - (blocks[1][1] - blocks[1][0])) <= 1
