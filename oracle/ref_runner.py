"""ORACLE-side runner of the reference's OWN CPU implementation (test / bench infrastructure, NOT product code).

`/root/reference` does not exist on the GPU box, and reference sources are never copied into this repository.
`ensure_ref_copy()` — called by `__graft_entry__.build()` in the authoring container — byte-compiles the one module the hot
path lives in (`resource-estimation/qrnn.py`, compiled where it lies) into `oracle/_ref/qrnn.pyc` (a build output:
git-ignored, but shipped to the GPU box with the snapshot like the built `.so`), so that `bench.py`'s `cpu_baseline` /
`--impl reference` legs time the UNMODIFIED reference module (`kind: "reference"`).  When the compiled module is absent the
torch port `oracle/qrnn_torch_cpu.py` stands in (`kind: "port"`).  Only `bench.py`'s CPU legs and tests
import this file.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import py_compile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/resource-estimation/qrnn.py"
REF_DST = os.path.join(HERE, "_ref", "qrnn.pyc")


def ensure_ref_copy() -> bool:
    """Byte-compile the reference module (from where it lies) into oracle/_ref/ — a build step; a no-op where
    /root/reference is absent (the GPU box uses the prebuilt file)."""
    if os.path.exists(REF_SRC):
        os.makedirs(os.path.dirname(REF_DST), exist_ok=True)
        if not os.path.exists(REF_DST) or os.path.getmtime(REF_DST) < os.path.getmtime(REF_SRC):
            py_compile.compile(REF_SRC, cfile=REF_DST, doraise=True)
    return os.path.exists(REF_DST)


def _import_ref():
    loader = importlib.machinery.SourcelessFileLoader("deeprest_reference_qrnn", REF_DST)
    spec = importlib.util.spec_from_loader("deeprest_reference_qrnn", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


class Runner:
    """The reference estimator on the host CPU: `kind` is "reference" (the byte-compiled reference qrnn.py) or "port"."""

    def __init__(self, blob, M, F, threads=None):
        import torch
        from deeprest_b200 import layout
        self.torch, self.M, self.F = torch, M, F
        if threads:
            torch.set_num_threads(threads)
        if os.path.exists(REF_DST):
            self.kind = "reference"
            self.model = _import_ref().QuantileRNN(input_size=F, num_metrics=M)
            sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in layout.state_dict_from_blob(blob, M, F).items()}
            self.model.load_state_dict(sd)
            self.model.eval()
        else:
            from oracle.qrnn_torch_cpu import TorchCpuPort
            self.kind = "port"
            self.model = TorchCpuPort(blob, M, F)

    def describe(self):
        t = self.torch
        what = ("the reference's own resource-estimation/qrnn.py (byte-compiled, unmodified, in oracle/_ref/)" if self.kind == "reference"
                else "reference algorithm restated on torch (oracle/qrnn_torch_cpu.py) incl. its O(M^2) stack/mean")
        return f"{what}, torch {t.__version__} CPU, {t.get_num_threads()} threads"

    def forward(self, x, chunk=16):
        """eval-mode forward of all experts, in chunks of <= `chunk` windows (exact: windows are independent; one call at
        BASELINE configs[1] would need > 62 GB)."""
        t = self.torch
        outs = []
        with t.no_grad():
            for b0 in range(0, len(x), chunk):
                xb = x[b0:b0 + chunk]
                o = self.model(t.from_numpy(np.ascontiguousarray(xb))) if self.kind == "reference" else self.model.forward(xb)
                outs.append(o.numpy() if hasattr(o, "numpy") else o)
        return np.concatenate(outs)

    def forward_sampled(self, x, expert_ids):
        """Forecasts of a SAMPLE of experts for the windows x, with the reference's arithmetic: every expert's masked
        bi-GRU (qrnn.py:33-44, needed for the mean), then for each sampled expert i the reference's own stack of the M-1
        other outputs, mean, concat and head (qrnn.py:46-54).  Used where all M heads would cost O(M^2) minutes."""
        t = self.torch
        if self.kind != "reference":
            raise RuntimeError("forward_sampled needs the reference module (oracle/_ref/qrnn.pyc)")
        m = self.model
        xs = t.from_numpy(np.ascontiguousarray(x))
        with t.no_grad():
            outs = []
            for l1, l2, rnn, _ in m.experts:
                mask = m.softmax(l2(m.relu(l1(m.mask_init))))
                seq = (xs * mask[None, None, :]).permute(1, 0, 2)
                h0 = t.zeros(2, xs.shape[0], m.hidden_layer_size)
                r, _ = rnn(seq, h0)
                outs.append(r.permute(1, 0, 2))
            preds = []
            for i in expert_ids:
                others = t.mean(t.stack([outs[j] for j in range(self.M) if j != i]), dim=0)
                preds.append(m.experts[i][-1](t.cat([others, outs[i]], dim=-1)))
            return t.stack(preds).permute(1, 2, 0, 3).contiguous().numpy()

    def time_step_sampled(self, x, k_experts):
        """(seconds, scale) of one bounded sample of a forward on x: the bi-GRUs of the first k experts and k heads, each head
        with a full-size stack of M-1 tensors (op shapes identical to the full forward).  GRU cost is exactly linear in the
        expert count (independent modules) and so is the cost of one head's stack/mean, so the full forward costs
        scale = M/k times the sample."""
        import time
        t = self.torch
        if self.kind != "reference":
            raise RuntimeError("needs the reference module")
        m = self.model
        xs = t.from_numpy(np.ascontiguousarray(x))
        t0 = time.perf_counter()
        with t.no_grad():
            outs = []
            for l1, l2, rnn, _ in list(m.experts)[:k_experts]:
                mask = m.softmax(l2(m.relu(l1(m.mask_init))))
                seq = (xs * mask[None, None, :]).permute(1, 0, 2)
                h0 = t.zeros(2, xs.shape[0], m.hidden_layer_size)
                r, _ = rnn(seq, h0)
                outs.append(r.permute(1, 0, 2))
            for i in range(k_experts):
                others = t.mean(t.stack([outs[j % k_experts] for j in range(self.M - 1)]), dim=0)
                m.experts[i][-1](t.cat([others, outs[i]], dim=-1))
        return time.perf_counter() - t0, self.M / float(k_experts)

    def train_step(self, x, y, lr=1e-3):
        """One iteration of the reference training loop (estimate.py:67-74) on the reference module; returns the loss."""
        t = self.torch
        if self.kind != "reference":
            raise RuntimeError("needs the reference module")
        if not hasattr(self, "opt"):
            self.opt = t.optim.Adam(self.model.parameters(), lr=lr)
        self.model.train()
        out = self.model(t.from_numpy(np.ascontiguousarray(x)))
        loss = self.model.quantile_loss(out, t.from_numpy(np.ascontiguousarray(y)))
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        self.model.eval()
        return float(loss.detach())
