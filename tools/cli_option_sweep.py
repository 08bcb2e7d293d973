"""CLI-level option sweep on a GPU box: for every (data set, option set) run the unmodified reference binary
(oracle/_ref/RabbitVar, built by oracle/Makefile; it travels with the snapshot) and the drop-in CLI
(build/rabbitvar_b200, the CUDA path) with the SAME flags and compare the sorted TSV lines field by field
(integers and strings exact, %f-printed doubles within one unit of the last printed digit).

    python tools/cli_option_sweep.py [--out gpurun_out/cli_sweep.txt] [name ...]

Prints one line per combination: data set, option set, reference lines, our lines, differing lines.
Test infrastructure: nothing here is on the product path."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import dumpcmp  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "RabbitVar")
OURS = os.path.join(ROOT, "build", "rabbitvar_b200")

# option sets appended to the case's own reference arguments (Launcher.cpp:295-460 names)
SIMPLE_SETS = {
    "base": [], "k0": ["-k", "0"], "p": ["-p"], "p_fisher": ["-p", "--fisher"], "fisher": ["--fisher"],
    "f0.05": ["-f", "0.05"], "Q30": ["-Q", "30"], "m3": ["-m", "3"], "M140": ["-M", "140"], "q30": ["-q", "30"],
    "X1": ["-X", "1"], "P10": ["-P", "10"], "r4": ["-r", "4"], "B4": ["-B", "4"], "o3": ["-o", "3"], "O30": ["-O", "30"],
    "V0.1": ["-V", "0.1"], "I20": ["-I", "20"], "F0": ["-F", "0"], "x50": ["-x", "50"], "Y600": ["-Y", "600"],
    "u": ["-u"], "UN": ["--UN"], "three": ["-3"], "t": ["-t"], "T100": ["-T", "100"], "T100_k0": ["-T", "100", "-k", "0"],
    "z": ["-z"], "T60_three_u": ["-T", "60", "-3", "-u"], "X4_q15_m5": ["-X", "4", "-q", "15", "-m", "5"],
}
SOMATIC_SETS = {
    "base": [], "k0": ["-k", "0"], "f0.05": ["-f", "0.05"], "Q30": ["-Q", "30"], "q30": ["-q", "30"], "r4": ["-r", "4"],
    "V0.1": ["-V", "0.1"], "u": ["-u"], "UN": ["--UN"], "three": ["-3"], "t": ["-t"], "m3": ["-m", "3"],
    "T100": ["-T", "100"], "X1": ["-X", "1"], "F0": ["-F", "0"], "M140": ["-M", "140"],
}


def lines_of(path):
    return sorted(l for l in open(path).read().splitlines() if l)


def line_equal(a, b):
    ta, tb = a.split("\t"), b.split("\t")
    return len(ta) == len(tb) and all(dumpcmp.fields_equal(x, y, 2e-6) for x, y in zip(ta, tb))


def run_one(binary, args, out):
    r = subprocess.run([binary] + args + ["--out", out], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    return r.returncode, r.stderr[-300:]


def main():
    argv = sys.argv[1:]
    report = None
    if "--out" in argv:
        i = argv.index("--out")
        report = open(argv[i + 1], "w")
        del argv[i:i + 2]
    todo = []
    for name in ("c1_k1", "c5_k1", "c3_k0", "dedup_t_F500"):
        todo += [(name, cases.CASES[name], s, o) for s, o in SIMPLE_SETS.items()]
    for name in ("c2_somatic_k1", "c2_somatic_bed_k0"):
        todo += [(name, cases.SOMATIC_CASES[name], s, o) for s, o in SOMATIC_SETS.items()]
    todo = [t for t in todo if not argv or t[0] in argv or t[2] in argv]

    def one(t):
        name, case, sname, opts = t
        d = cases.dataset_dir(name)
        args = case["ref_args"](d) + opts
        f_ref, f_ours = f"/tmp/sweep_{name}_{sname}_ref.tsv", f"/tmp/sweep_{name}_{sname}_ours.tsv"
        rc1, e1 = run_one(REF, args, f_ref)
        rc2, e2 = run_one(OURS, args, f_ours)
        if rc1 or rc2:
            return 1, f"{name}\t{sname}\trc ref={rc1} ours={rc2}\t{(e2 if rc2 else e1).strip()[-200:]!r}"
        want, got = lines_of(f_ref), lines_of(f_ours)
        if len(want) != len(got):
            only_w = sorted(set(want) - set(got))[:2]
            only_g = sorted(set(got) - set(want))[:2]
            return 1, f"{name}\t{sname}\tref {len(want)}\tours {len(got)}\tLINE COUNT DIFFERS\t{only_w}\t{only_g}"
        bad = [(w, g) for w, g in zip(want, got) if w != g and not line_equal(w, g)]
        msg = f"{name}\t{sname}\tref {len(want)}\tours {len(got)}\tdiffering {len(bad)}"
        if bad:
            w, g = bad[0]
            tw, tg = w.split("\t"), g.split("\t")
            cols = [i for i, (x, y) in enumerate(zip(tw, tg)) if not dumpcmp.fields_equal(x, y, 2e-6)]
            msg += f"\tfirst: cols {cols} key {tw[2:7]} ref {[tw[i] for i in cols][:6]} ours {[tg[i] for i in cols][:6]}"
        return (1 if bad else 0), msg

    for name in sorted(set(t[0] for t in todo)):
        cases.generate(name)
    n_bad = 0
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=8) as ex:
        for bad, msg in ex.map(one, todo):
            n_bad += bad
            print(msg, flush=True)
            if report:
                report.write(msg + "\n")
                report.flush()
    print(f"combinations with differences: {n_bad}")
    if report:
        report.write(f"combinations with differences: {n_bad}\n")


if __name__ == "__main__":
    main()
