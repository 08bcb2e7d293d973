#!/usr/bin/env python
"""Full-size parity of the five BASELINE.json configs: the unmodified reference binary (oracle/_ref/RabbitVar,
built by oracle/Makefile) against the drop-in CLI (build/rabbitvar_b200, the CUDA path) on the same seeded
synthetic BAM/FASTA/BED, same flags, TSV compared as sorted multisets of lines:
integer/string fields identical, %f-printed doubles within 2e-6 relative (the reference prints 6 decimals).

    python tools/parity_configs.py [--configs 1,1R,2,3,4,5] [--scale 1.0] [--gpus 1] [--th N] [--keep] [--json out.json]

Test + bench infrastructure (needs a GPU for the CLI side and oracle/_ref for the reference side); the product never
imports it.  `--scale` shrinks every contig (scale 0.1 = a tenth of the tiles) for quick runs; 1.0 = BASELINE size.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
BUILD = os.path.join(ROOT, "build")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "RabbitVar")
CLI = os.path.join(BUILD, "rabbitvar_b200")

# BASELINE.json configs: synthgen preset, full contig length, CLI flags of that config
CONFIGS = {
    "1": dict(cfg=1, length=1002600, chrom="chrS1", bam="S.bam", name="S", bed="tiles.bed", flags=["-f", "0.01"]),
    "1R": dict(cfg=1, length=1002600, chrom="chrS1", bam="S.bam", name="S", bed=None, flags=["-f", "0.01"]),
    "2": dict(cfg=2, length=5002600, chrom="chrS2", bam="T.bam|N.bam", name="T|N", bed="tiles.bed",
              flags=["-f", "0.01", "--fisher"]),
    "3": dict(cfg=3, length=2002600, chrom="chrS3", bam="S.bam", name="S", bed="panel.bed", flags=["-f", "0.005"]),
    "4": dict(cfg=4, length=50002600, chrom="chrS4", bam="S.bam", name="S", bed="tiles.bed", flags=["-f", "0.01"]),
    "5": dict(cfg=5, length=5002600, chrom="chrS5", bam="S.bam", name="S", bed="tiles.bed", flags=["-f", "0.01", "-3", "-u"]),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def dataset(key, scale, level=1):
    c = CONFIGS[key]
    length = c["length"] if scale >= 1.0 else max(22600, int((c["length"] - 2600) * scale) // 10000 * 10000 + 2600)
    d = os.path.join(ROOT, "_work", f"parity_c{c['cfg']}_{length}_l{level}")
    if not os.path.exists(os.path.join(d, "meta.txt")):
        os.makedirs(d, exist_ok=True)
        t0 = time.time()
        gen = [os.path.join(BUILD, "synthgen"), "--cfg", str(c["cfg"]), "--out", d, "--len", str(length), "--level", str(level)]
        if c["cfg"] == 3 and scale < 1.0:
            gen += ["--amplicons", str(max(3, int(500 * scale)))]
        subprocess.run(gen, check=True, stderr=subprocess.DEVNULL)
        log(f"[parity] generated {d} in {time.time() - t0:.1f}s")
    return d, length


def cli_args(key, d, length):
    c = CONFIGS[key]
    bam = "|".join(os.path.join(d, b) for b in c["bam"].split("|"))
    a = ["-G", os.path.join(d, "ref.fa"), "-b", bam, "-N", c["name"]]
    if c["bed"]:
        a += ["-i", os.path.join(d, c["bed"]), "-c", "1", "-S", "2", "-E", "3", "-g", "4"]
    else:
        a += ["-R", f"{c['chrom']}:1301-{length - 1300}"]
    return a + c["flags"]


def tsv_lines(path):
    with open(path) as f:
        return sorted(l for l in f.read().splitlines() if l)


def compare_tsv(want, got, rel_tol=2e-6):
    """Sorted-multiset comparison.  Returns (n_differing, examples)."""
    import dumpcmp
    if want == got:
        return 0, []

    def eq(a, b):
        if a == b:
            return True
        ta, tb = a.split("\t"), b.split("\t")
        return len(ta) == len(tb) and all(dumpcmp.fields_equal(x, y, rel_tol) for x, y in zip(ta, tb))

    def key(l):  # chr, start, end, ref, alt: identifies a line across the two outputs
        t = l.split("\t")
        return tuple(t[2:7])

    wk, gk = {}, {}
    for l in want:
        wk.setdefault(key(l), []).append(l)
    for l in got:
        gk.setdefault(key(l), []).append(l)
    bad, ex = 0, []
    for k in sorted(set(wk) | set(gk)):
        a, b = wk.get(k, []), gk.get(k, [])
        if len(a) != len(b):
            bad += max(len(a), len(b))
            if len(ex) < 6:
                ex.append(("missing" if len(a) > len(b) else "extra", k, (a[:1], b[:1])))
            continue
        for x, y in zip(a, b):
            if not eq(x, y):
                bad += 1
                if len(ex) < 6:
                    tx, ty = x.split("\t"), y.split("\t")
                    cols = [i for i in range(min(len(tx), len(ty))) if not dumpcmp.fields_equal(tx[i], ty[i], rel_tol)]
                    ex.append(("differs", k, [(i, tx[i], ty[i]) for i in cols[:6]]))
    return bad, ex


def run_config(key, scale=1.0, gpus=1, threads=None, keep=False, extra_cli=(), level=1):
    """Runs one config through both binaries.  Returns a dict for the bench line / the JSON report."""
    threads = threads or os.cpu_count() or 1
    d, length = dataset(key, scale, level)
    args = cli_args(key, d, length)
    ref_out, got_out = os.path.join(d, f"ref_{key}.tsv"), os.path.join(d, f"gpu_{key}_g{gpus}.tsv")
    res = {"config": key, "length": length, "scale": scale, "gpus": gpus}
    # single -R region is single-threaded by construction in the reference (simpleMode.cpp:218)
    if not (os.path.exists(ref_out) and keep):
        t0 = time.perf_counter()
        r = subprocess.run([REF_BIN] + args + ["--th", str(threads), "--out", ref_out], stdout=subprocess.DEVNULL,
                           stderr=subprocess.PIPE, text=True)
        res["ref_sec"] = time.perf_counter() - t0
        if r.returncode != 0:
            res["error"] = f"reference rc={r.returncode}: {r.stderr[-300:]}"
            return res
    t0 = time.perf_counter()
    r = subprocess.run([CLI] + args + ["--th", str(threads), "--gpus", str(gpus), "--out", got_out] + list(extra_cli),
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    res["cli_sec"] = time.perf_counter() - t0
    res["cli_stdout"] = r.stdout[-400:]
    if r.returncode != 0:
        res["error"] = f"cli rc={r.returncode}: {r.stderr[-600:]}"
        return res
    if r.stderr.strip():
        res["cli_stderr"] = r.stderr[-400:]
    want, got = tsv_lines(ref_out), tsv_lines(got_out)
    bad, ex = compare_tsv(want, got)
    res.update({"ref_lines": len(want), "cli_lines": len(got), "parity_lines_differing": bad, "examples": ex})
    if "|" in CONFIGS[key]["bam"]:
        try:
            a = [float(x) for x in open(ref_out + ".info").read().split()]
            b = [float(x) for x in open(got_out + ".info").read().split()]
            res["info_equal"] = len(a) == len(b) and all(abs(x - y) <= 2e-6 * max(1.0, abs(x)) for x, y in zip(a, b))
        except Exception as e:
            res["info_equal"] = False
            res["info_error"] = str(e)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,1R,2,3,4,5")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--th", type=int, default=0)
    ap.add_argument("--keep", action="store_true", help="reuse an existing reference TSV")
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    out = []
    for key in args.configs.split(","):
        r = run_config(key, args.scale, args.gpus, args.th or None, args.keep)
        out.append(r)
        log(f"[parity] config {key}: " + json.dumps({k: v for k, v in r.items() if k not in ("examples", "cli_stdout")}))
        for e in r.get("examples", [])[:6]:
            log("    ", e)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)
    bad = sum(1 for r in out if r.get("error") or r.get("parity_lines_differing", 1) != 0)
    print(json.dumps({"configs": len(out), "configs_differing": bad,
                      "parity_lines_differing": {r["config"]: r.get("parity_lines_differing") for r in out}}))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
