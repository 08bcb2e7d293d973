#!/usr/bin/env python
"""Markdown table + traffic.json fields from an ncu report:  python tools/ncu_table.py rep.ncu-rep"""
import csv, json, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return float("nan")
def unit_scale(k, r):  # ncu prints mixed units per column header row 1
    return rows[1][ix[k]]
print("| kernel | time ms | DRAM read GB | DRAM write GB | DRAM GB/s | L2 % | warps active % | regs | threads/inst | issue active % | warp-inst |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
per = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    t = f(r, "gpu__time_duration.sum"); tu = unit_scale("gpu__time_duration.sum", r)
    t_ms = t * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(tu.replace("second", "s").replace("m", "m") if tu in ("ms", "us", "ns", "s") else tu, 1)
    def gb(k):
        v = f(r, k); u = unit_scale(k, r)
        return v * {"Gbyte": 1, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "Tbyte": 1e3}[u]
    rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
    per[name] = (rd + wr) * 1e9
    print(f"| {name} | {t_ms:.3f} | {rd:.3f} | {wr:.3f} | {(rd + wr) / t_ms * 1e3:.0f} | {f(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {int(f(r, 'launch__registers_per_thread'))} | "
          f"{f(r, 'smsp__thread_inst_executed_per_inst_executed.ratio'):.1f} | {f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{f(r, 'smsp__inst_executed.sum'):.3g} |")
print(json.dumps(per, indent=1))
