#!/usr/bin/env python
"""Precision probe: can the x-part of the recurrence drop one of its three split-fp16 terms
(x_hi*W_lo or x_lo*W_hi) and still meet the fp32 parity bar |a-b| <= 1e-6 + 1e-4|b| on every golden?
Prints, per golden and variant, the worst ratio err / bound (<= 1 passes)."""
import glob, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from conftest import golden_cases, load_golden
from deeprest_b200 import QuantileRNN

res = {}
for path in golden_cases():
    g = load_golden(path)
    if g["F"] > 64:
        continue
    ref = g["out"] if "out" in g else g["out_f32"]
    m = QuantileRNN(g["F"], g["M"], engine="tcgen05").eval()
    m.load_blob(g["blob_arr"])
    row = {}
    for drop in (0, 1, 2):
        m.debug_read(f"tc_xdrop{drop}", 1)
        out = m(g["x"])
        err = np.abs(out.astype(np.float64) - ref.astype(np.float64))
        row[{0: "3 terms", 1: "drop x_hi*W_lo", 2: "drop x_lo*W_hi"}[drop]] = {
            "worst_err_over_bound": float((err / (1e-6 + 1e-4 * np.abs(ref))).max()), "max_abs": float(err.max())}
    m.close()
    res[os.path.basename(path)] = row
print(json.dumps(res, indent=1))
