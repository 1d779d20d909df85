#!/usr/bin/env python
"""Sum an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (cold-cache, serialised times:
use the SHARES, not the absolutes)."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    unit = r[mu]
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(unit, 1e-6)
    name = re.sub(r"\(.*", "", r[kn])
    tot[name][0] += 1; tot[name][1] += v
s = sum(v for _, v in tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {n:8d} {v:10.3f} {100*v/s:6.1f}%")
print(f"{'total':60s} {sum(n for n,_ in tot.values()):8d} {s:10.3f}")
