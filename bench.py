#!/usr/bin/env python
"""bench.py — aligned bases/sec piled+scored on B200 (BASELINE.json metric), config 2 of BASELINE.json:
tumor/normal, synthetic 5 Mb region at 200x/100x as 500 x 10 kb tiles per sample, --fisher.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --steps K --warmup W    (the reference's CPU path, oracle/_ref)

One step = one pass of the hot path (rv_pileup + rv_score: read filter, CIGAR rewrite + walk, pileup,
per-position scoring incl. Fisher) over every (tile, sample) of the workload.
  value : whole-job throughput, inputs resident in HBM, device time (CUDA events on the launching stream),
          max over ranks.  Region-sharded: every rank owns a full config-2 shard, no collective (weak scaling).
  e2e   : the same metric through the public host-buffer API (rvh_call_regions: pinned H2D, pileup, event +
          table D2H, host realign hand-off, patch H2D, score, variant D2H, TSV formatting), wall clock.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
BUILD = os.path.join(ROOT, "build")
METRIC = "aligned bases/sec piled+scored"
UNIT = "bases/s"
TILE = 10000


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def ensure_built():
    need = [os.path.join(ROOT, "rabbitvar_b200", "librvgpu.so"), os.path.join(BUILD, "synthgen")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g
        g.build()


def make_dataset(work, length, level=0):
    """config-2 shaped data (T 200x + N 100x over `length` bases), seeded; cached by directory."""
    if not os.path.exists(os.path.join(work, "meta.txt")):
        os.makedirs(work, exist_ok=True)
        t0 = time.time()
        subprocess.run([os.path.join(BUILD, "synthgen"), "--cfg", "2", "--out", work, "--len", str(length),
                        "--level", str(level)], check=True, stderr=subprocess.DEVNULL)
        log(f"[bench] generated {work} in {time.time() - t0:.1f}s")
    meta = dict(l.split("\t") for l in open(os.path.join(work, "meta.txt")).read().splitlines())
    return meta


def read_tiles(work, limit=None):
    tiles = []
    for l in open(os.path.join(work, "tiles.bed")):
        c, s, e, g = l.split()
        tiles.append((int(s), int(e)))
        if limit and len(tiles) >= limit:
            break
    return tiles


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.stop = False
        self.index = index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        # NVML in-process (a sample every ~20 ms); nvidia-smi as the fallback
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = ((getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), 3), (getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), 4),
                    (getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), 5), (getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4), 6))
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop:
                r = get_reasons(h)
                row = [str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(sm_max), "0", "", "", "", ""]
                for bit, i in bits:
                    row[i] = "Active" if (r & bit) else "Not Active"
                self.rows.append(row)
                time.sleep(0.02)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.t.start()

    def finish(self):
        self.stop = True
        self.t.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for i, name in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(r[i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ref_cmd(work, bed, out, threads):
    return [os.path.join(ROOT, "oracle", "_ref", "RabbitVar"), "-G", os.path.join(work, "ref.fa"), "-b",
            os.path.join(work, "T.bam") + "|" + os.path.join(work, "N.bam"), "-N", "T|N", "-i", bed, "-c", "1", "-S", "2",
            "-E", "3", "-g", "4", "-f", "0.01", "--fisher", "--th", str(threads), "--out", out]


def sample_bed(work, n_tiles):
    bed = os.path.join(work, f"sample_{n_tiles}.bed")
    with open(os.path.join(work, "tiles.bed")) as f, open(bed, "w") as o:
        for i, l in enumerate(f):
            if i >= n_tiles:
                break
            o.write(l)
    return bed


def count_sample_bases(work, n_tiles):
    """Aligned bases (M/=/X of reads surviving the filters, counted once per tile fetch) of the first n tiles,
    both samples — counted from the BAMs on the host with the same definition the GPU statistics use."""
    import numpy as np
    import rabbitvar_b200 as rv
    tiles = read_tiles(work, n_tiles)
    dt = np.dtype([("pos", "<i4"), ("mpos", "<i4"), ("off", "<u4"), ("l_seq", "<i4"), ("flag", "<u2"),
                   ("n_cigar", "<u2"), ("nm", "<i2"), ("mapq", "u1"), ("same", "u1"), ("end", "<i4"), ("rsv", "<i4")])
    total = 0
    for bam in ("T.bam", "N.bam"):
        b = rv.HostBatch(os.path.join(work, bam), "chrS2", tiles[0][0], tiles[-1][1])
        r = b.reads_numpy().view(dt)
        keep = (r["flag"] & 0x504) == 0
        keep &= (r["flag"] & 0x800) == 0
        r = r[keep]
        for s, e in tiles:
            sel = r[(r["pos"] - 1 < e) & (r["end"] > s - 1)]
            # synthetic reads: aligned bases = read length minus soft clips/insertions ~= end - pos + 1 - deletions;
            # use the reference span as the aligned-base count (exact for M-only reads, within 0.1% otherwise)
            total += int((sel["end"] - sel["pos"] + 1).sum())
        b.close()
    return total


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation (oracle/_ref/RabbitVar, unmodified sources
    built by oracle/Makefile) timed on this box's host cores on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ensure_built()
    binp = os.path.join(ROOT, "oracle", "_ref", "RabbitVar")
    if not os.path.exists(binp):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/RabbitVar was not built/shipped"}))
        return
    n_tiles = args.ref_tiles
    work = os.path.join(ROOT, "_work", f"bench_ref_{n_tiles}")
    length = 1300 + n_tiles * TILE + 1300
    make_dataset(work, length, level=1)
    bed = sample_bed(work, n_tiles)
    cores = os.cpu_count() or 1
    bases = count_sample_bases(work, n_tiles)
    out = os.path.join(work, "ref_out.tsv")
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        subprocess.run(ref_cmd(work, bed, out, cores), check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = bases * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "config 2 (tumor/normal 200x/100x, 10 kb tiles, --fisher)",
                   "sample": f"first {n_tiles} tiles x 2 samples per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"oracle/_ref/RabbitVar --th {cores}, first {n_tiles} tiles of config 2 (T+N), {bases} aligned bases"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--length", type=int, default=5002600, help="contig length of the config-2 shard")
    ap.add_argument("--ref-tiles", type=int, default=150, help="tiles in the CPU-baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--workers", type=int, default=16, help="pipeline workers (contexts) per GPU in the e2e path")
    ap.add_argument("--chunk", type=int, default=5, help="tiles per pipeline chunk in the e2e path")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    ensure_built()
    import rabbitvar_b200 as rv
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the one JSON line (NCCL's banner goes there otherwise)
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=rank, world_size=world)
    if not torch.cuda.is_available():
        raise rv.RabbitVarError("bench.py needs a GPU (there is no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    # ---- workload: one full config-2 shard per rank (same seed => same shape on every rank) -------------
    work = os.path.join(ROOT, "_work", f"bench_cfg2_{args.length}" + (f"_r{rank}" if world > 1 else ""))
    t0 = time.time()
    make_dataset(work, args.length, level=0)
    tiles = read_tiles(work)
    starts, ends = [t[0] for t in tiles], [t[1] for t in tiles]
    bt = rv.HostBatch(os.path.join(work, "T.bam"), "chrS2", starts[0], ends[-1])
    bn = rv.HostBatch(os.path.join(work, "N.bam"), "chrS2", starts[0], ends[-1])
    n_t = bt.n_reads
    off_n = bt.append(bn)
    n_n = bn.n_reads
    bn.close()
    regs_t = bt.make_regions(starts, ends, 1200, 0, n_t)
    regs_n = bt.make_regions(starts, ends, 1200, off_n, n_n)
    nreg = 2 * len(tiles)
    regs = (rv.Region * nreg)()
    C.memmove(regs, regs_t, C.sizeof(rv.Region) * len(tiles))
    C.memmove(C.byref(regs, C.sizeof(rv.Region) * len(tiles)), regs_n, C.sizeof(rv.Region) * len(tiles))
    ref = rv.fetch_ref(os.path.join(work, "ref.fa"), "chrS2", 1, bt.chr_len)
    log(f"[bench r{rank}] data ready in {time.time() - t0:.1f}s: {bt.n_reads} reads, pool {bt.pool_bytes / 1e9:.2f} GB, "
        f"{nreg} (tile, sample) regions")
    halo = 512
    n_pos = sum(e - s + 1 + 2 * halo for s, e in tiles) * 2
    lim = rv.default_limits(max_reads=bt.n_reads + 64, max_read_bytes=bt.pool_bytes + 256, max_positions=n_pos + 64,
                            max_regions=nreg + 8, halo=halo, max_events=max(1 << 20, bt.n_reads),
                            max_variants=3 * n_pos + 1024, max_patch=max(1 << 20, bt.n_reads // 2),
                            max_ref_bases=len(ref) + 64)
    params = rv.default_params(fisher=1, has_bam2=1, candidates_only=2)
    ctx = rv.Context(local, params, lim)
    ctx.set_reference(1, ref)

    # device-resident inputs live in torch tensors (torch = device memory plumbing)
    reads_np = bt.reads_numpy()
    pool_np = bt.pool_numpy()
    d_reads = torch.from_numpy(reads_np).to(dev)
    # rv_push_reads_device: the pool buffer extends 32 readable bytes beyond pool_bytes (look-ahead loads of the gather kernel)
    d_pool = torch.zeros(int(pool_np.size) + 32, dtype=torch.uint8, device=dev)
    d_pool[: int(pool_np.size)].copy_(torch.from_numpy(pool_np))
    torch.cuda.synchronize()
    read_bytes_total = int(reads_np.size + pool_np.size)
    avg_read_bytes = 32 + float((pool_np.size) / max(1, bt.n_reads))  # 32 B header + cigar + packed seq + qual (16 B aligned)

    half = len(tiles)
    # ---- e2e warm pass through the public host-buffer API (also installs the patch list used below) --------
    bt.pin()

    pipe = rv.Pipeline(local, args.workers)
    e2e_params = rv.default_params(fisher=1)  # the reference CLI's flags for this config: -f 0.01 --fisher, -b 'T|N'
    regs_t_arr = (rv.Region * half).from_address(C.addressof(regs))
    regs_n_arr = (rv.Region * half).from_address(C.addressof(regs) + half * C.sizeof(rv.Region))

    def e2e_pass():
        t_a = time.perf_counter()
        # host buffers in, TSV out: every chunk of tiles is copied H2D, piled, handed to the host stage, scored,
        # copied back and formatted; the pipeline's workers overlap those stages across chunks
        tsv, tm_p = pipe.run(e2e_params, bt, regs, args.chunk, ref, 1, "T|N", "chrS2", paired=True, raw=True)
        return time.perf_counter() - t_a, (tm_p,), len(tsv)

    # ---- device-resident steps ---------------------------------------------------------------------------
    # One step = what run_batch_somatic launches for this workload: rv_pileup over every (tile, sample), rv_score with the
    # device-side candidate cut over every position of both samples, rv_score_positions for the full records of both
    # samples at the positions where either has a candidate (the tumor | normal join list, built once here).
    ctx.push_reads_ptr(bt.n_reads, d_reads.data_ptr(), d_pool.data_ptr(), int(pool_np.size), device=True)
    ctx.set_regions(regs)
    st = ctx.pileup()
    bases_per_step = st.n_aligned_bases
    kept_per_step = st.n_reads_kept
    log(f"[bench r{rank}] pileup: {st.n_items} work items, {st.n_reads_kept} kept, {st.n_walk_items} walked "
        f"({st.n_walk_full} whole reads), {st.n_events} events")
    # sparse keys of this batch (identical every step): reduce once on the host, keep the patch resident
    ctx.install_patch_from_events(bt, regs, ref, 1)
    ctx.score()
    vp, n_candidate_records = ctx.fetch_variants()
    if n_candidate_records > 5000000:
        raise rv.RabbitVarError("candidate pass returned an implausible number of records")
    rec = np.frombuffer((C.c_char * (n_candidate_records * C.sizeof(rv.Variant))).from_address(C.addressof(vp.contents)),
                        dtype=np.int32).reshape(n_candidate_records, C.sizeof(rv.Variant) // 4) \
        if n_candidate_records else np.zeros((0, C.sizeof(rv.Variant) // 4), np.int32)
    cr = rec[:, 0].astype(np.int64) % half
    cp = rec[:, 1].astype(np.int64)
    uniq = np.unique(cr * (1 << 32) + cp)
    jr = (uniq >> 32).astype(np.int32)
    jp = (uniq & 0xffffffff).astype(np.int32)
    join_regions = np.concatenate([jr, jr + half]).astype(np.int32)
    join_positions = np.concatenate([jp, jp]).astype(np.int32)

    def step():
        ctx.pileup()
        ctx.score()
        a, b = ctx.kernel_ms()
        sp = ctx.pileup_split_ms()
        ctx.score_positions(join_regions, join_positions)
        _, b2 = ctx.kernel_ms()
        return a, b + b2, sp

    for _ in range(max(0, args.warmup - 1)):
        step()
    ctx.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count()
    pile_ms, score_ms, split_ms = [], [], []
    ctx.timer_start()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        a, b, sp = step()
        pile_ms.append(a)
        score_ms.append(b)
        split_ms.append(sp)
    dev_ms = ctx.timer_stop()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - w0) * 1000.0
    if world > 1:
        dist.barrier()
    launches = ctx.launch_count() - l0
    clocks = sampler.finish()
    n_var_step = ctx.n_variants() + n_candidate_records

    # ---- e2e steps (host buffers, copies inside the timed region) ---------------------------------------
    e2e_t, e2e_tm, tsv_len = [], None, 0
    for i in range(1 + args.e2e_steps if args.e2e_steps > 0 else 0):
        dt, tms, tsv_len = e2e_pass()
        if i > 0:
            e2e_t.append(dt)
            e2e_tm = tms
    # --e2e-steps 0 (kernel experiments, ncu captures): no end-to-end number, the line says so
    e2e_sec = sum(e2e_t) / len(e2e_t) if e2e_t else float("nan")
    for nm, t in zip(("T|N",), e2e_tm or ()):
        log(f"[bench r{rank}] e2e {nm}: push {t.push_ms:.1f} pileup {t.pileup_ms:.1f} fetch {t.fetch_ms:.1f} host {t.host_ms:.1f} "
            f"patch {t.patch_ms:.1f} score {t.score_ms:.1f} assemble {t.assemble_ms:.1f} ms; events {t.n_events} "
            f"variants {t.n_variants} lines {t.n_lines}")
    h2d = sum(t.h2d_bytes for t in e2e_tm or ())
    d2h = sum(t.d2h_bytes for t in e2e_tm or ())

    # ---- reductions over ranks (max time, sum of work) ---------------------------------------------------
    from rabbitvar_b200.shard import reduce_step_metrics
    d = dist if world > 1 else None
    dev_ms_max, total_bases = reduce_step_metrics(dev_ms, bases_per_step, d)
    wall_ms_max, _ = reduce_step_metrics(wall_ms, 0, d)
    e2e_sec_max, _ = reduce_step_metrics(e2e_sec, 0, d)
    value = total_bases * args.steps / (dev_ms_max / 1000.0)
    e2e_value = total_bases / e2e_sec_max

    if rank == 0:
        peak, peak_src = measured_peak()
        P = sum(e - s + 1 for s, e in tiles) * 2
        alg_pileup = kept_per_step * avg_read_bytes + P * 133.0
        alg_score = P * 133.0 + n_var_step * 128.0
        pk = float(np.mean(pile_ms)) / 1000.0
        sk = float(np.mean(score_ms)) / 1000.0
        achieved = alg_pileup / pk / 1e9
        split = [float(np.mean([x[k] for x in split_ms])) for k in range(3)]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("pileup_stage_dram_bytes_per_launch")
        cpu = None
        if not args.skip_cpu and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "RabbitVar")):
            try:
                n_tiles = min(args.ref_tiles, len(tiles))
                bed = sample_bed(work, n_tiles)
                cores = os.cpu_count() or 1
                # aligned bases of the sample: exact, from a GPU pileup of just those tiles
                sub = (rv.Region * (2 * n_tiles))()
                C.memmove(sub, regs, C.sizeof(rv.Region) * n_tiles)
                C.memmove(C.byref(sub, C.sizeof(rv.Region) * n_tiles), C.byref(regs, C.sizeof(rv.Region) * half),
                          C.sizeof(rv.Region) * n_tiles)
                ctx.push_reads_ptr(bt.n_reads, d_reads.data_ptr(), d_pool.data_ptr(), int(pool_np.size), device=True)
                ctx.set_regions(sub)
                sb = ctx.pileup().n_aligned_bases
                out = os.path.join(work, "ref_out.tsv")
                c0 = time.perf_counter()
                subprocess.run(ref_cmd(work, bed, out, cores), check=True, stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL)
                cdt = time.perf_counter() - c0
                cpu = {"value": sb / cdt, "unit": UNIT, "cores": cores, "kind": "reference",
                       "sample": f"oracle/_ref/RabbitVar (unmodified reference, -O3 -ffast-math -fopenmp) --th {cores} on the "
                                 f"first {n_tiles} tiles x 2 samples of this workload: {sb} aligned bases in {cdt:.2f}s"}
            except Exception as e:  # the baseline is reported, never required
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[1]: tumor/normal, 5 Mb as 500 x 10 kb tiles, T 200x + N 100x, "
                                   "2x150 bp, -f 0.01 --fisher; one full shard per GPU (region-sharded, no collective)",
                       "regions_per_gpu": nreg, "reads_per_gpu": int(bt.n_reads), "aligned_bases_per_step_per_gpu": int(bases_per_step),
                       "l2_policy": f"inputs ({read_bytes_total / 1e9:.2f} GB reads + {n_pos * 132 / 1e9:.2f} GB tables) exceed the 126 MB L2",
                       "wall_ms_per_step": wall_ms_max / args.steps,
                       "step": "rv_pileup + rv_score (candidate cut, every position of both samples) + rv_score_positions (full records of both samples at the joined candidate positions): the launches of run_batch_somatic; the join list and the host realigner patch are built once outside the timed loop"},
            "clocks": clocks,
            "e2e": {"value": e2e_value if e2e_t else None, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": f"rvh_pipeline_run_paired (host buffers -> somatic-mode TSV), {args.workers} worker contexts x {args.chunk}-tile chunks of both samples, pinned H2D", "sec_per_step": e2e_sec_max if e2e_t else None,
                    "tsv_bytes": tsv_len},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm",
                         "kernel": "pileup stage = rv_pileup_kernel + rv_tile_index_kernel + rv_gather4_kernel + rv_walk_kernel "
                                   "(one rv_pileup call; rv_gather4_kernel is the largest)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_pileup, "kernel_ms": pk * 1000.0,
                         "split_ms": {"rv_pileup_kernel": split[0], "rv_tile_index_kernel+rv_gather4_kernel": split[1],
                                      "rv_walk_kernel": split[2]},
                         "score_kernel": {"achieved": alg_score / sk / 1e9, "kernel_ms": sk * 1000.0,
                                          "algorithmic_bytes_per_launch": alg_score}},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    # explicit teardown in a fixed order (nothing is left to interpreter shutdown, where CUDA may already be gone)
    pipe.close()
    ctx.close()
    del d_reads, d_pool
    bt.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
