#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel): headline metrics, stall mix, top stalled SASS lines.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25


def page(name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


rows = page("raw")
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sectors_op_red.sum"]
print("== headline")
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"  {h:72s} {v} {u}")
print("== warp stall mix (stalled warps per issue-active cycle)")
st = [(float(v), h) for h, v in zip(hdr, vals) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
for v, h in sorted(st, reverse=True)[:9]:
    print(f"  {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''):24s} {v:.3f}")
rows = page("source")
h2 = rows[1]
data = rows[2:]
ix = {h: i for i, h in enumerate(h2)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


tot = sum(f(r, "# Samples") for r in data)
keys = [k for k in h2 if k.startswith("stall_") and "Not Issued" not in k]
print(f"== top SASS lines by stall samples (total {tot:.0f})")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:ntop]:
    s = f(r, "# Samples")
    br = " ".join(f"{k[6:]}={f(r, k):.0f}" for k in keys if f(r, k) > 0.1 * s)
    print(f"  {s:6.0f} {100 * s / tot:5.1f}%  {r[ix['Source']][:62]:62s} {br}")
