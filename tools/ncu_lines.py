#!/usr/bin/env python
"""Attribute an ncu SASS source page to CUDA source lines.

  ncu -i rep.ncu-rep --page source --csv --kernel-name regex:KERNEL > sass.csv
  cuobjdump -xelf all librvgpu.so ; nvdisasm -g -c x.cubin > dis.txt
  python tools/ncu_lines.py sass.csv dis.txt MANGLED_SUBSTRING [top]

Instructions are matched by their order inside the function (same build as the one profiled)."""
import csv
import re
import sys


def main():
    sass_csv, dis, func = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lines = open(dis).read().splitlines()
    insts = []  # (file:line) per instruction in order
    infn = False
    cur = "?"
    for l in lines:
        if l.startswith("\t.section\t.text."):
            infn = func in l
            continue
        if l.startswith("\t.section"):
            infn = False
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = m.group(1).split("/")[-1] + ":" + m.group(2)
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            insts.append(cur)
    rows = list(csv.reader(open(sass_csv)))
    h = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
    hdr = rows[h]
    ix = {n: i for i, n in enumerate(hdr)}
    body = []
    for r in rows[h + 1:]:
        if r and r[0] == "Kernel Name":
            break  # a second capture of the same kernel follows: keep the first
        if len(r) >= len(hdr) - 2:
            body.append(r)
    if len(body) != len(insts):
        print(f"warning: {len(body)} profiled instructions vs {len(insts)} disassembled", file=sys.stderr)
    agg = {}
    ti = ts = 0
    for k, r in enumerate(body):
        key = insts[k] if k < len(insts) else "?"
        i = int(r[ix["Instructions Executed"]] or 0)
        s = int(r[ix["# Samples"]] or 0)
        t = int(r[ix["Thread Instructions Executed"]] or 0)
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += i
        a[1] += s
        a[2] += t
        ti += i
        ts += s
    print(f"total warp-instructions {ti}, samples {ts}")
    print(" inst%  samp%  thr/inst  line")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{100.0 * a[0] / max(ti, 1):6.2f} {100.0 * a[1] / max(ts, 1):6.2f}  {a[2] / max(a[0], 1):6.1f}   {key}")


if __name__ == "__main__":
    main()
