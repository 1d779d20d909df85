"""ORACLE-side CPU baseline port (test/bench infrastructure, NOT product code).

A torch-CPU restatement of the reference forward that keeps the reference's *cost
structure* — one ``nn.GRU`` call per expert (qrnn.py:24,41) and the O(M²)
stack-then-mean of the other experts' outputs for every head (qrnn.py:46-52) — so that
timing it on the GPU box's host cores stands in for "the reference's own CPU
implementation" (which is Python and cannot travel with the repo).  Used only by
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs and by tests that pin it
against the golden vectors.
"""
from __future__ import annotations

import numpy as np
import torch

from deeprest_b200.layout import H, unpack_blob


class TorchCpuPort:
    def __init__(self, blob, M, F):
        self.M, self.F = M, F
        self.experts = []
        for ex in unpack_blob(np.asarray(blob, np.float32), M, F):
            t = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in ex.items()}
            gru = torch.nn.GRU(F, H, num_layers=1, bidirectional=True)
            with torch.no_grad():
                gru.weight_ih_l0.copy_(t["w_ih_f"]); gru.weight_hh_l0.copy_(t["w_hh_f"])
                gru.bias_ih_l0.copy_(t["b_ih_f"]); gru.bias_hh_l0.copy_(t["b_hh_f"])
                gru.weight_ih_l0_reverse.copy_(t["w_ih_r"]); gru.weight_hh_l0_reverse.copy_(t["w_hh_r"])
                gru.bias_ih_l0_reverse.copy_(t["b_ih_r"]); gru.bias_hh_l0_reverse.copy_(t["b_hh_r"])
            gru.eval()
            self.experts.append((t, gru))
        self.one = torch.ones(1)

    @torch.no_grad()
    def forward(self, x):
        x = torch.as_tensor(x, dtype=torch.float32)
        B = x.shape[0]
        outs = []
        for t, gru in self.experts:
            hid = torch.relu(torch.nn.functional.linear(self.one, t["mask_w1"], t["mask_b1"]))
            mask = torch.softmax(torch.nn.functional.linear(hid, t["mask_w2"], t["mask_b2"]), dim=-1)
            seq = (x * mask[None, None, :]).permute(1, 0, 2)
            h0 = torch.zeros(2, B, H)
            r, _ = gru(seq, h0)
            outs.append(r.permute(1, 0, 2))
        preds = []
        for i, (t, _) in enumerate(self.experts):
            others = torch.stack([outs[j] for j in range(self.M) if j != i])
            feat = torch.cat([others.mean(dim=0), outs[i]], dim=-1)
            preds.append(torch.nn.functional.linear(feat, t["head_w"], t["head_b"]))
        return torch.stack(preds).permute(1, 2, 0, 3).contiguous().numpy()
