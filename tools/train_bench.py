#!/usr/bin/env python
"""Side measurement (bench.py carries the contract line): one training step (dropout, pinball loss, full backward, Adam) at a
given shape, device resident (CUDA events around dr_train_step_dev), e.g. BASELINE configs[2] = 256 services (512 experts) x
batch 4096 x T=288 in bf16.  Prints JSON."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deeprest_b200 import QuantileRNN, synth

ap = argparse.ArgumentParser()
ap.add_argument("--experts", type=int, default=64)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--seq-len", type=int, default=288)
ap.add_argument("--features", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--engine", default="auto")
ap.add_argument("--dtype", default="bf16")
a = ap.parse_args()
M, B, T, F = a.experts, a.batch, a.seq_len, a.features
dev = torch.device("cuda", 0)
m = QuantileRNN(F, M, engine=a.engine, dtype=a.dtype)
m.load_blob(synth.weights(11, M, F))
x = torch.from_numpy(synth.windows(1, B, T, F)).to(dev)
# labels: a cheap deterministic pattern (the counter-based generator would need B*T*M draws on the host)
y = ((torch.arange(T, device=dev, dtype=torch.float32)[None, :, None] % 17) / 17.0).expand(B, T, M).contiguous()
for i in range(a.warmup):
    m.train_step(x, y, seed=1 + i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = m.launch_count
e0.record()
for i in range(a.steps):
    loss = m.train_step(x, y, seed=100 + i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / max(a.steps, 1)
flops = 3.0 * (1536 * F + 199680) * M * B * T
print(json.dumps({"what": "train_step", "dtype": a.dtype, "experts": M, "services": M // 2, "batch": B, "seq_len": T, "features": F,
                  "engine_cfg": a.engine, "forward_engine": m.last_engine, "ms_per_step": round(ms, 2),
                  "service_windows_per_s": round(M // 2 * B / (ms * 1e-3), 1), "algorithmic_tflops_per_s_3x_fwd": round(flops / (ms * 1e-3) / 1e12, 1),
                  "launches_per_step": (m.launch_count - l0) // max(a.steps, 1), "loss": float(loss),
                  "free_gb_after": round(torch.cuda.mem_get_info()[0] / 2**30, 1),
                  "timing": "CUDA events around dr_train_step_dev, inputs resident"}), flush=True)
m.close()
