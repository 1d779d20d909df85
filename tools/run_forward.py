#!/usr/bin/env python
"""Short driver for profilers: a few forwards of one shape on one engine (no timing claims)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from deeprest_b200 import QuantileRNN, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=16)
ap.add_argument("--B", type=int, default=1024)
ap.add_argument("--T", type=int, default=96)
ap.add_argument("--F", type=int, default=64)
ap.add_argument("--engine", default="tcgen05")
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
model = QuantileRNN(a.F, a.M, engine=a.engine).eval()
model.load_blob(synth.weights(11, a.M, a.F))
x = torch.from_numpy(synth.windows(2021, a.B, a.T, a.F)).cuda()
for _ in range(a.iters):
    out = model(x)
torch.cuda.synchronize()
print("ok", model.last_engine, float(out.abs().mean()))
