#!/usr/bin/env python
"""Side measurement (not the bench.py contract): one training step (dropout, pinball loss, full backward, Adam) at a
given shape, e.g. BASELINE configs[2] = 256 services (512 experts) x batch 4096 x T=288.  Prints JSON."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from deeprest_b200 import QuantileRNN, synth

ap = argparse.ArgumentParser()
ap.add_argument("--experts", type=int, default=64)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--seq-len", type=int, default=288)
ap.add_argument("--features", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--engine", default="auto")
a = ap.parse_args()
M, B, T, F = a.experts, a.batch, a.seq_len, a.features
m = QuantileRNN(F, M, engine=a.engine)
m.load_blob(synth.weights(11, M, F))
x = synth.windows(1, B, T, F)
# labels: a cheap deterministic pattern (the counter-based generator would need B*T*M draws on the host)
y = np.broadcast_to((np.arange(T, dtype=np.float32)[None, :, None] % 17) / 17.0, (B, T, M)).copy()
t0 = time.perf_counter()
for i in range(a.warmup):
    m.train_step(x, y, seed=1 + i)
t1 = time.perf_counter()
for i in range(a.steps):
    loss = m.train_step(x, y, seed=100 + i)
dt = (time.perf_counter() - t1) / max(a.steps, 1)
print(json.dumps({"what": "train_step", "experts": M, "services": M // 2, "batch": B, "seq_len": T, "features": F, "engine_cfg": a.engine,
                  "forward_engine": m.last_engine, "ms_per_step": round(dt * 1e3, 2), "windows_per_s": round(B / dt, 1),
                  "service_windows_per_s": round(M // 2 * B / dt, 1), "loss": float(loss), "warmup_s": round(t1 - t0, 2),
                  "timing": "host wall clock around dr_train_step with host buffers (H2D of x,y inside)"}), flush=True)
m.close()
