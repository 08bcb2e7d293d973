#!/usr/bin/env python
"""Where the wall clock of one drop-in CLI run goes that the CLI's own clock does not see: process start, and the
process exit after `_exit` (the kernel's address-space teardown, then the release of the CUDA device files).

    python tools/cli_anatomy.py [--config 2] [--runs 3]

The parent times three points per run: spawn, end-of-file on the child's stdout (the child's descriptors are closed
after its address space is gone, standard streams first) and the return of wait().  Variants: the default
environment and CUDA_DEVICE_MAX_CONNECTIONS = 4 / 1.  Measurement tool (needs a GPU); nothing imports it.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_configs as pc


def one(cmd, env):
    t0 = time.perf_counter()
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=env)
    out = p.stdout.read()
    t1 = time.perf_counter()
    rc = p.wait()
    t2 = time.perf_counter()
    s = out.decode(errors="replace")
    m = re.search(r"total time: ([0-9.]+) s", s)
    u = re.search(r"cuda start-up (\d+)", s)
    return dict(rc=rc, to_eof=round(t1 - t0, 3), to_exit=round(t2 - t0, 3), cli_total=float(m.group(1)) if m else None,
                cuda_startup=int(u.group(1)) / 1000.0 if u else None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="2")
    ap.add_argument("--runs", type=int, default=3)
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    d, length = pc.dataset(a.config, 1.0)
    cmd = [pc.CLI] + pc.cli_args(a.config, d, length) + ["--th", str(os.cpu_count() or 1), "--out", os.path.join(d, "anatomy.tsv")]
    res = {}
    variants = [("default", {}), ("max_connections_4", {"CUDA_DEVICE_MAX_CONNECTIONS": "4"}),
                ("max_connections_1", {"CUDA_DEVICE_MAX_CONNECTIONS": "1"})]
    one(cmd, dict(os.environ))  # file cache, driver persistence
    for name, extra in variants:
        env = dict(os.environ)
        env.update(extra)
        res[name] = [one(cmd, env) for _ in range(a.runs)]
        print(name, json.dumps(res[name]), flush=True)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
