#!/usr/bin/env python
"""Side measurements quoted in DESIGN.md (not the bench.py contract): BASELINE configs[0] latency on both
engines, and the fp32 training step on a reduced configs[2] shape."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deeprest_b200 import QuantileRNN, synth


def lat(engine, M, B, T, F, reps=40):
    m = QuantileRNN(F, M, engine=engine).eval(); m.load_blob(synth.weights(11, M, F))
    x = torch.from_numpy(synth.windows(1, B, T, F)).cuda()
    for _ in range(5): m(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); m(x); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    xh = synth.windows(1, B, T, F)
    t0 = time.perf_counter()
    for _ in range(reps): m(xh)
    host = (time.perf_counter() - t0) / reps * 1e3
    m.close()
    return float(np.median(ts)), host


for eng in ("ffma", "tcgen05"):
    d, h = lat(eng, 2, 1, 64, 16)
    print(f"configs[0] (1 service, 1 window, T=64, F=16) {eng:8s}: device {d:.3f} ms, host-buffer call {h:.3f} ms")
d, h = lat("tcgen05", 2, 32, 60, 16)
print(f"estimate.py default shape (M=2,B=32,T=60,F=16) tcgen05: device {d:.3f} ms, host-buffer call {h:.3f} ms")

M, B, T, F = 64, 256, 288, 64
m = QuantileRNN(F, M); m.load_blob(synth.weights(11, M, F))
x = synth.windows(1, B, T, F); y = synth.labels(2, B, T, M)
m.train_step(x, y, seed=1)
t0 = time.perf_counter()
for i in range(3): loss = m.train_step(x, y, seed=2 + i)
dt = (time.perf_counter() - t0) / 3
print(f"train step (fp32 CUDA-core path) M={M} B={B} T={T} F={F}: {dt*1e3:.1f} ms/step = {B/dt:.0f} windows/s ({M//2*B/dt:.0f} service-windows/s), loss {loss:.4f}")
m.close()
