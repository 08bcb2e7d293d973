"""H2D rate of the registered (cudaHostRegister) read batch vs torch-pinned memory, alone and in 16 concurrent streams.
   python tools/pcie_probe.py   (GPU box)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import rabbitvar_b200 as rv

work = os.path.join(ROOT, "_work", "bench_cfg2_5002600")
bench.make_dataset(work, 5002600, level=0)
tiles = bench.read_tiles(work)
bt = rv.HostBatch(os.path.join(work, "T.bam"), "chrS2", tiles[0][0], tiles[-1][1])
bt.pin()
pool = torch.from_numpy(bt.pool_numpy())
n = pool.numel()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
print("registered batch pool bytes", n, "is_pinned", pool.is_pinned())
def timed(f, label):
    for _ in range(3):
        torch.cuda.synchronize(); t = time.perf_counter(); f(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"{label}: {n / dt / 1e9:.1f} GB/s ({dt * 1e3:.1f} ms)")
timed(lambda: d.copy_(pool, non_blocking=True), "cudaHostRegister'd vector, one copy")
h = torch.empty(n, dtype=torch.uint8).pin_memory(); h.copy_(pool)
timed(lambda: d.copy_(h, non_blocking=True), "cudaHostAlloc (torch pin_memory), one copy")
streams = [torch.cuda.Stream() for _ in range(16)]
piece = 27 << 20
def multi(src):
    k = 0
    for o in range(0, n, piece):
        with torch.cuda.stream(streams[k % 16]):
            d[o:o + piece].copy_(src[o:o + piece], non_blocking=True)
        k += 1
timed(lambda: multi(pool), "registered, 27 MB pieces round-robin on 16 streams")
timed(lambda: multi(h), "pin_memory, 27 MB pieces round-robin on 16 streams")
