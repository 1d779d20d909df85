#!/usr/bin/env python
"""Top stalled SASS lines of one kernel in an .ncu-rep with several kernels: tools/ncu_stalls.py rep regex [ntop]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; data = [r for r in rows[2:] if len(r) == len(h)]
ix = {k: i for i, k in enumerate(h)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
keys = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(f(r, k) for r in data) for k in keys}
print("== stall mix over all samples:", " ".join(f"{k[6:]}={100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
print(f"== top SASS lines by stall samples (total {tot:.0f}); instructions executed total {sum(f(r,'Instructions Executed') for r in data):.0f}")
for i, r in sorted(enumerate(data), key=lambda ir: -f(ir[1], "# Samples"))[:ntop]:
    s = f(r, "# Samples")
    br = " ".join(f"{k[6:]}={f(r, k):.0f}" for k in keys if f(r, k) > 0.15 * s)
    print(f"  line {i:5d} {s:6.0f} {100 * s / tot:5.1f}%  {r[ix['Source']].strip()[:70]:70s} {br}")
