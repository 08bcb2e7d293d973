"""Regenerates the committed golden vectors by running the UNMODIFIED reference (oracle/_ref, built by
oracle/Makefile from /root/reference) on seeded synthetic data.  Only runs where /root/reference exists.

    python tests/golden/make_golden.py [case ...]

Each case = (synthgen arguments, reference CLI arguments).  Outputs, gzipped, under tests/golden/:
    <case>.dump.txt.gz   stage dumps of oracle/_ref/ref_dump   (C, R, V stages)
    <case>.tsv.gz        sorted TSV of oracle/_ref/RabbitVar    (the reference binary's own output)
The inputs are NOT committed: synthgen is deterministic (seeded xoshiro), tests regenerate them.
"""
import gzip
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import CASES, SOMATIC_CASES, dataset_dir, generate  # noqa: E402


def main():
    ref_dump = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "RabbitVar")
    only = set(sys.argv[1:])  # optional: case names to regenerate (default: all)
    for name, case in CASES.items():
        if only and name not in only:
            continue
        d = generate(name)
        out_dump = os.path.join(d, "ref.dump.txt")
        env = dict(os.environ, RV_DUMP=out_dump, RV_DUMP_STAGES=case.get("stages", "CRV"))
        subprocess.run([ref_dump] + case["ref_args"](d), check=True, env=env, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
        with open(out_dump, "rb") as f, gzip.open(os.path.join(ROOT, "tests", "golden", name + ".dump.txt.gz"), "wb", 6) as g:
            g.write(f.read())
        tsv = os.path.join(d, "ref.tsv")
        subprocess.run([ref_bin] + case["ref_args"](d) + ["--out", tsv], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
        lines = sorted(open(tsv).read().splitlines())
        with gzip.open(os.path.join(ROOT, "tests", "golden", name + ".tsv.gz"), "wt") as g:
            g.write("\n".join(lines) + "\n")
        print(name, os.path.getsize(os.path.join(ROOT, "tests", "golden", name + ".dump.txt.gz")), "bytes dump,",
              len(lines), "tsv lines")
    for name, case in SOMATIC_CASES.items():  # paired mode: the reference binary's TSV only
        if only and name not in only:
            continue
        d = generate(name)
        tsv = os.path.join(d, "ref.tsv")
        subprocess.run([ref_bin] + case["ref_args"](d) + ["--out", tsv], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
        lines = sorted(open(tsv).read().splitlines())
        with gzip.open(os.path.join(ROOT, "tests", "golden", name + ".tsv.gz"), "wt") as g:
            g.write("\n".join(lines) + "\n")
        print(name, len(lines), "tsv lines; .info =", open(tsv + ".info").read().split())


if __name__ == "__main__":
    main()
