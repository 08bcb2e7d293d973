"""Region sharding across GPUs (one process per GPU, no data-path collective).

The reference parallelises over regions with an OpenMP dynamic loop (src/modes/simpleMode.cpp:320) and one
process per NUMA socket (run_rabbitvar.py:72-108).  Here the ordered tile list is cut into contiguous blocks
balanced by expected work (reads per tile), block g goes to GPU g, and the host concatenates the per-block
output in tile order.  torch.distributed is used only for the barrier and for the max-over-ranks timing.
"""


def contiguous_blocks(weights, n_parts):
    """Cut range(len(weights)) into n_parts contiguous [lo, hi) blocks with near-equal weight sums.

    Deterministic; every block is non-empty while there are at least n_parts items; blocks cover the
    whole range in order (tile order is the output order)."""
    n = len(weights)
    n_parts = max(1, int(n_parts))
    if n == 0:
        return [(0, 0)] * n_parts
    total = float(sum(weights))
    blocks, lo, acc = [], 0, 0.0
    for part in range(n_parts):
        remaining_parts = n_parts - part
        if part == n_parts - 1:
            hi = n
        else:
            target = (total - acc) / remaining_parts
            hi, run = lo, 0.0
            max_hi = min(n, max(lo + 1, n - (remaining_parts - 1)))  # leave one item for every later block
            while hi < max_hi and (hi == lo or run + weights[hi] / 2.0 <= target):
                run += weights[hi]
                hi += 1
            acc += run
        blocks.append((lo, hi))
        lo = hi
    return blocks


def reduce_step_metrics(step_ms, units, dist=None):
    """max over ranks of the step time, sum over ranks of the units processed (whole-job throughput)."""
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return float(step_ms), float(units)
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(step_ms)], dtype=torch.float64, device=dev)
    u = torch.tensor([float(units)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


def concat_in_order(parts, dist=None):
    """Gathers per-rank text blocks on rank 0 in rank (= tile) order. Returns the joined text on rank 0."""
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return "".join(parts)
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, "".join(parts))
    return "".join(out)
