#!/usr/bin/env python
"""Per-kernel SASS opcode counts of libdeeprest_b200.so (what proves the Blackwell-native paths, B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk, UTMALDG/UTMASTG = tensor-map TMA,
HMMA = legacy mma.sync, SYNCS = mbarrier, RED = red.global, FFMA2 = packed fp32.  Usage: tools/sass_counts.py [lib.so]"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deeprest_b200", "libdeeprest_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA.2CTA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "SYNCS", "RED", "ATOM", "MUFU", "FFMA2", "FFMA", "LDG", "STG", "LDGSTS", "UCGABAR", "MEMBAR"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "")
        cur = re.sub(r"^void ", "", cur)
        if "<" in cur.split("(")[0]:                      # template: keep <...>, drop the parameter list after it
            depth, end = 0, len(cur)
            for i, ch in enumerate(cur):
                depth += ch == "<"
                depth -= ch == ">"
                if ch == ">" and depth == 0:
                    end = i + 1
                    break
            cur = cur[:end].replace("(bool)", "")
        else:
            cur = cur.split("(")[0]
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for p in pats:
            if op == p or op.startswith(p + "."):
                if p == "UTCHMMA" and op.startswith("UTCHMMA.2CTA"):
                    continue
                if p == "FFMA" and op.startswith("FFMA2"):
                    continue
                if p == "LDG" and op.startswith("LDGSTS"):
                    continue
                counts[cur][p] += 1
print(f"{'kernel':44s} {'instr':>7s} " + " ".join(f"{p[:9]:>9s}" for p in pats))
for k, c in counts.items():
    print(f"{k[:44]:44s} {c['_total']:7d} " + " ".join(f"{c[p]:9d}" for p in pats))
