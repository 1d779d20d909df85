#!/usr/bin/env python
"""Diagnostic: one bf16-engine training step vs the fp32 numpy oracle at small shapes; prints per-parameter-family errors
(max |got - ref| / max |ref|) so that tolerances in tests/test_gpu_train.py are set from measurements."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from deeprest_b200 import QuantileRNN, layout, synth
from oracle import qrnn_numpy as oracle

shapes = [(3, 5, 7, 5, 0), (2, 9, 4, 16, 0), (4, 6, 12, 33, 0), (2, 300, 5, 16, 0), (2, 300, 3, 8, 140), (2, 130, 40, 64, 0)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
res = []
for M, B, T, F, mb in shapes:
    if mb:
        os.environ["DR_TRAIN_MICROBATCH"] = str(mb)
    else:
        os.environ.pop("DR_TRAIN_MICROBATCH", None)
    blob = synth.weights(40 + M, M, F, 1.5)
    x = synth.windows(3, B, T, F, "diurnal")
    y = synth.labels(4, B, T, M)
    dm = (synth.uniform(8, M * B * T * 2 * layout.H) >= 0.5).astype(np.uint8).reshape(M, B, T, 2 * layout.H)
    ref_loss, ref_out, ref_g = oracle.loss_and_grads(blob, x, y, M, F, dropout_masks=dm.astype(np.float32))
    m = QuantileRNN(F, M, dtype="bf16")
    m.load_blob(blob)
    loss = m.train_step(x, y, lr=1e-3, dropout_mask=dm)
    g = m.grads()
    eng = m.last_engine
    m.close()
    fam = {}
    pe = layout.params_per_expert(F)
    for name, (off, shape) in layout.expert_offsets(F).items():
        n = int(np.prod(shape))
        worst = 0.0
        for e in range(M):
            a, b = g[e * pe + off:e * pe + off + n], ref_g[e * pe + off:e * pe + off + n]
            worst = max(worst, float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)))
        fam[name] = round(worst, 5)
    res.append({"shape": [M, B, T, F, mb], "engine": eng, "loss": float(loss), "ref_loss": float(ref_loss),
                "grad_err_over_max": float(np.abs(g - ref_g).max() / np.abs(ref_g).max()), "families": fam,
                "finite": bool(np.isfinite(g).all())})
    print(json.dumps(res[-1]), flush=True)
