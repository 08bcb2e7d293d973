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


def bai_tile_weights(bai_path, tid, tiles):
    """Expected work of every tile = its share of the compressed BAM bytes of the 16 kb index windows it touches
    (BAI linear index, SAM spec 5.2) — the 'balanced by expected reads' weight of SURVEY 8(e).
    tiles: [(start1, end1)] 1-based inclusive.  Falls back to the tile lengths when the index has no entry."""
    import struct
    with open(bai_path, "rb") as f:
        data = f.read()
    if data[:4] != b"BAI\x01":
        raise ValueError("not a BAI file: " + bai_path)
    off = 4
    (n_ref,) = struct.unpack_from("<i", data, off)
    off += 4
    ioffset = []
    for r in range(n_ref):
        (n_bin,) = struct.unpack_from("<i", data, off)
        off += 4
        for _ in range(n_bin):
            _, n_chunk = struct.unpack_from("<Ii", data, off)
            off += 8 + 16 * n_chunk
        (n_intv,) = struct.unpack_from("<i", data, off)
        off += 4
        if r == tid:
            ioffset = list(struct.unpack_from("<%dQ" % n_intv, data, off))
        off += 8 * n_intv
    if not ioffset:
        return [float(e - s + 1) for s, e in tiles]
    for i in range(1, len(ioffset)):  # empty windows inherit the previous offset
        if ioffset[i] == 0:
            ioffset[i] = ioffset[i - 1]

    def coff(w):
        return ioffset[min(max(w, 0), len(ioffset) - 1)] >> 16

    # bytes of window w: up to the next window's first offset; the last window gets the mean of the others
    n_w = len(ioffset)
    wbytes = [max(0, coff(w + 1) - coff(w)) for w in range(n_w)]
    filled = [b for b in wbytes[:-1] if b > 0]
    if n_w:
        wbytes[-1] = sum(filled) / len(filled) if filled else 0.0
    out = []
    for s, e in tiles:
        w0, w1 = max(s - 1, 0) >> 14, max(e - 1, 0) >> 14
        total = 0.0
        for w in range(w0, w1 + 1):
            if w >= n_w:
                break
            lo, hi = max(s - 1, w << 14), min(e, (w + 1) << 14)  # the tile's share of the window
            total += wbytes[w] * max(0, hi - lo) / 16384.0
        out.append(total if total > 0 else float(e - s + 1) * 1e-3)
    return out
