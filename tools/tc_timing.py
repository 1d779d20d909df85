#!/usr/bin/env python
"""In-kernel cycle breakdown of the tcgen05 recurrence (work item 0): who waits for whom."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deeprest_b200 import QuantileRNN, synth
M, B, T, F = 128, 1024, 288, 64
if len(sys.argv) > 1: M = int(sys.argv[1])
model = QuantileRNN(F, M, engine="tcgen05").eval()
model.load_blob(synth.weights(11, M, F))
x = torch.from_numpy(synth.windows(2021, B, T, F)).cuda()
model(x); torch.cuda.synchronize()
model.debug_read("tc_timing_on", 1)
model(x); torch.cuda.synchronize()
c = model.debug_read("tc_timing", 18) / T
print(f"per step (cycles), T={T}:")
print(f"  epilogue warp0: total {c[0]:.0f} | wait GATE_FULL q0..3 {c[1]:.0f} {c[2]:.0f} {c[3]:.0f} {c[4]:.0f} | ld+rearm {c[5]:.0f} | math+st {c[6]:.0f} | head dot+RED {c[7]:.0f}")
print(f"  MMA thread    : total {c[8]:.0f} | wait X_FULL {c[9]:.0f} | wait GATE_FREE q0..3 {c[10]:.0f} {c[11]:.0f} {c[12]:.0f} {c[13]:.0f} | wait H_READY kq0..3 {c[14]:.0f} {c[15]:.0f} {c[16]:.0f} {c[17]:.0f}")
