"""Mint golden vectors by EXECUTING the reference (authoring container only).

Imports ``/root/reference/resource-estimation/qrnn.py`` (never copied into this
repo), runs it on torch CPU and writes small ``.npz`` fixtures to
``tests/golden/``.  ``/root/reference`` does not exist on the GPU box, so tests
only ever read the committed fixtures.  Re-run:  ``python oracle/make_golden.py``.

Inputs and (except G1) weights are pure functions of integer seeds
(``deeprest_b200.synth``), so fixtures hold outputs + checksums, not tensors.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/resource-estimation")

from qrnn import QuantileRNN  # noqa: E402  (the reference itself)
from deeprest_b200 import layout, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name, M, B, T, F, weight source, weight scale, x kind
CASES = [
    ("g1_cfg1_torchinit", 2, 1, 64, 16, "torch", 1.0, "uniform"),   # BASELINE configs[0]
    ("g2_small", 4, 8, 64, 16, "synth", 1.0, "uniform"),
    ("g2b_small_diurnal", 4, 8, 64, 16, "synth", 1.0, "diurnal"),
    ("g3_long", 2, 2, 1440, 64, "synth", 1.0, "diurnal"),           # configs[4] horizon
    ("g4_saturating", 4, 8, 64, 16, "synth", 3.0, "diurnal"),       # "trained-like": saturating gates
    ("g6_odd", 3, 5, 7, 5, "synth", 2.0, "uniform"),                 # ragged: odd M, tiny T, F not /4
    ("g7_f64_wide", 6, 130, 24, 64, "synth", 1.5, "uniform"),        # B crosses a 128-row tile
]
WSEED, XSEED, YSEED = 11, 2021, 2022


def build_model(M, F, src, scale, dtype=torch.float32):
    torch.manual_seed(0)
    model = QuantileRNN(input_size=F, num_metrics=M)
    if src == "synth":
        blob = synth.weights(WSEED, M, F, scale)
        sd = {k: torch.from_numpy(np.ascontiguousarray(v))
              for k, v in layout.state_dict_from_blob(blob, M, F).items()}
        model.load_state_dict(sd)
    else:
        blob = layout.blob_from_state_dict(model.state_dict(), M, F)
    return model.to(dtype), blob


def run_fp64(model, x):
    """Same module in double precision (error floor). qrnn.py:39 builds h0 with the
    default dtype, so flip the default while it runs."""
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        with torch.no_grad():
            return model.double()(torch.from_numpy(x).double()).numpy()
    finally:
        torch.set_default_dtype(prev)
        model.float()


class ReplayDropout(torch.nn.Module):
    """Stands in for ``model.dropout`` (qrnn.py:15,43): same arithmetic as
    nn.Dropout(p) — ``x * mask / (1-p)`` — but with a supplied mask per call."""

    def __init__(self, masks, p):
        super().__init__()
        self.masks, self.p, self.i = masks, p, 0

    def forward(self, t):
        m = self.masks[self.i]
        self.i += 1
        return t * (m / (1.0 - self.p))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name, M, B, T, F, src, scale, kind in CASES:
        model, blob = build_model(M, F, src, scale)
        model.eval()
        x = synth.windows(XSEED, B, T, F, kind)
        y = synth.labels(YSEED, B, T, M)
        with torch.no_grad():
            out = model(torch.from_numpy(x))
            loss = model.quantile_loss(out, torch.from_numpy(y)).item()
        out64 = run_fp64(model, x)
        payload = dict(
            M=M, B=B, T=T, F=F, wsrc=src, wscale=scale, wseed=WSEED, xseed=XSEED, yseed=YSEED,
            xkind=kind, out=out.numpy().astype(np.float32), out64=out64.astype(np.float64),
            loss=np.float32(loss), blob_sum=np.float64(blob.astype(np.float64).sum()),
            x_sum=np.float64(x.astype(np.float64).sum()), torch_version=torch.__version__,
        )
        if src == "torch":
            payload["blob"] = blob
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **payload)
        print(f"{name}: out {out.shape} |out|max {np.abs(out.numpy()).max():.4f} "
              f"fp32-fp64 {np.abs(out.numpy() - out64).max():.2e} loss {loss:.6f}")

    # G5 — one full training step with a replayed dropout mask (estimate.py:67-74).
    M, B, T, F = 2, 4, 16, 16
    model, blob = build_model(M, F, "synth", 1.0)
    model.train()
    x = synth.windows(XSEED, B, T, F, "diurnal")
    y = synth.labels(YSEED, B, T, M)
    mask_np = (synth.uniform(77, M * B * T * 2 * layout.H) >= 0.5).astype(np.float32)
    mask_np = mask_np.reshape(M, B, T, 2 * layout.H)
    model.dropout = ReplayDropout([torch.from_numpy(m) for m in mask_np], 0.5)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)      # estimate.py:61
    out = model(torch.from_numpy(x))
    loss = model.quantile_loss(out, torch.from_numpy(y))
    opt.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    opt.step()
    gblob = layout.blob_from_state_dict(grads, M, F)
    wblob = layout.blob_from_state_dict(model.state_dict(), M, F)
    np.savez_compressed(
        os.path.join(OUT, "g5_train_step.npz"), M=M, B=B, T=T, F=F, wseed=WSEED, wscale=1.0,
        xseed=XSEED, yseed=YSEED, xkind="diurnal", mask_seed=77, lr=np.float32(1e-3),
        out=out.detach().numpy(), loss=np.float32(loss.item()), grads=gblob,
        weights_after=wblob, blob_sum=np.float64(blob.astype(np.float64).sum()),
        torch_version=torch.__version__)
    print(f"g5_train_step: loss {loss.item():.6f} |grad|max {np.abs(gblob).max():.3e}")

    # G8 — host-side helpers either side of the path (utils.py:4-5, qrnn.py:69-75).
    sys.path.insert(0, "/root/reference/resource-estimation")
    from utils import sliding_window
    ts = synth.uniform(5, 40 * 3).reshape(40, 3).astype(np.float64) * 7.0 - 1.0
    win = sliding_window(ts, 6)
    nm, lo, hi = QuantileRNN.normalization_minmax(win.copy(), split=12)
    np.savez_compressed(os.path.join(OUT, "g8_window_norm.npz"), ts=ts, window=6, split=12,
                        win=win, norm=nm, lo=lo, hi=hi)
    print("g8_window_norm:", win.shape, lo, hi)


def trained_golden():
    """G10 — weights after actually TRAINING the reference for 60 Adam steps (estimate.py:65-74 loop, default
    dropout) on a learnable synthetic task: the feature masks become peaky and the gates saturate the way a trained
    DeepRest model's do.  The blob travels in the fixture (torch's RNG stream for dropout cannot be regenerated)."""
    M, B, T, F = 2, 16, 60, 16
    torch.manual_seed(1)
    model = QuantileRNN(input_size=F, num_metrics=M)
    x = synth.windows(31, B, T, F, "diurnal")
    y = np.stack([np.clip(0.6 * x[:, :, 0] + 0.3 * x[:, :, 3], 0, 1), np.clip(x[:, :, 5] ** 2, 0, 1)], axis=-1).astype(np.float32)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    model.train()
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    for _ in range(60):
        loss = model.quantile_loss(model(xt), yt)
        opt.zero_grad(); loss.backward(); opt.step()
    model.eval()
    blob = layout.blob_from_state_dict(model.state_dict(), M, F)
    xe = synth.windows(XSEED, 6, T, F, "diurnal")
    with torch.no_grad():
        out = model(torch.from_numpy(xe))
        l = model.quantile_loss(out, torch.from_numpy(synth.labels(YSEED, 6, T, M))).item()
    out64 = run_fp64(model, xe)
    np.savez_compressed(os.path.join(OUT, "g10_trained.npz"), M=M, B=6, T=T, F=F, wsrc="torch", wscale=1.0, wseed=0,
                        xseed=XSEED, yseed=YSEED, xkind="diurnal", out=out.numpy(), out64=out64, loss=np.float32(l),
                        blob=blob, blob_sum=np.float64(blob.astype(np.float64).sum()),
                        x_sum=np.float64(xe.astype(np.float64).sum()), torch_version=torch.__version__,
                        train_loss=np.float32(loss.item()))
    mk = torch.softmax(model.experts[0][1](torch.relu(model.experts[0][0](model.mask_init))), -1)
    print(f"g10_trained: train loss {loss.item():.4f}, max mask {mk.max().item():.3f}, |out|max {out.abs().max().item():.3f}, "
          f"fp32-fp64 {np.abs(out.numpy() - out64).max():.2e}")


def synthetic_buckets(seed, n_buckets=6):
    """Small random call trees over a social-network-like component set (own generator, counter based)."""
    comps = ["nginx-thrift", "compose-post-service", "text-service", "user-mention-service", "media-mongodb", "post-storage"]
    ops = ["Compose", "Upload", "Read", "Store", "Find"]
    u = iter(synth.uniform(seed, 20000))

    def tree(depth):
        node = {"component": comps[int(next(u) * len(comps))], "operation": ops[int(next(u) * len(ops))], "children": []}
        if depth < 3:
            for _ in range(int(next(u) * 3)):
                node["children"].append(tree(depth + 1))
        return node

    buckets = []
    for _ in range(n_buckets):
        traces = [tree(0) for _ in range(1 + int(next(u) * 6))]
        metrics = [{"component": c, "resource": r, "value": float(next(u) * 100)} for c in comps[:3] for r in ("cpu", "memory")]
        buckets.append({"traces": traces, "metrics": metrics})
    return buckets


def featurize_golden():
    """G9 — the featurizer (featurize.py:11-101) run on synthetic buckets AND on the reference's own shipped sample."""
    import json
    import pickle
    import featurize as ref_feat                       # the reference module (functions only; its script part is under __main__)
    cases = {"synthetic": synthetic_buckets(123)}
    with open("/root/reference/resource-estimation/raw_data.pkl", "rb") as f:
        cases["shipped_sample"] = pickle.load(f)
    out = {}
    for name, raw in cases.items():
        Mf = {}
        for bucket in raw:
            Mf = ref_feat.construct_feature_space(Mf, bucket["traces"])
        traffic = np.asarray([ref_feat.extract_feature(Mf, bucket["traces"]) for bucket in raw])
        inv = {}
        for bucket in raw:
            c = ref_feat.count_invocations(bucket["traces"])
            for k, v in c.items():
                inv.setdefault(k, [0] * len(raw))
        for i, bucket in enumerate(raw):
            for k, v in ref_feat.count_invocations(bucket["traces"]).items():
                inv[k][i] = v
        out[name] = {"raw": raw, "keys": list(Mf.keys()), "traffic": traffic.tolist(), "invocations": inv}
    with open(os.path.join(OUT, "g9_featurize.json"), "w") as f:
        json.dump(out, f)
    print("g9_featurize:", {k: (len(v["keys"]), len(v["raw"])) for k, v in out.items()})


def synthesizer_golden():
    """G11 — the trace synthesizer (synthesizer.py:10-52) run on synthetic buckets and on the shipped sample.
    The reference's synthesize() uses the removed alias np.int; it is restored for the run (np.int = int)."""
    import json
    import pickle
    import synthesizer as ref_syn                      # the reference module (script part is under __main__)
    if not hasattr(np, "int"):
        np.int = int
    cases = {"synthetic": synthetic_buckets(123)}
    with open("/root/reference/resource-estimation/raw_data.pkl", "rb") as f:
        cases["shipped_sample"] = pickle.load(f)
    out = {}
    for name, raw in cases.items():
        ref = ref_syn.TraceSynthesizer().fit(raw)
        apis = list(ref.api2dist)
        dist = {api: {"candidates": [eval(c) for c in ref.api2dist[api][0]], "weights": list(ref.api2dist[api][1])} for api in apis}
        requests, vectors = [], []
        for k in range(6):
            u = synth.uniform(900 + k, 2 * len(apis))
            req = {api: int(u[2 * i] * 9) for i, api in enumerate(apis) if u[2 * i + 1] < 0.8 or i == 0}
            np.random.seed(4242 + k)
            vectors.append([int(v) for v in ref.synthesize(req)])
            requests.append(req)
        out[name] = {"raw": raw, "keys": list(ref.M.keys()), "api2dist": dist, "seed0": 4242, "requests": requests, "vectors": vectors}
    with open(os.path.join(OUT, "g11_synthesizer.json"), "w") as f:
        json.dump(out, f)
    print("g11_synthesizer:", {k: (len(v["api2dist"]), len(v["requests"])) for k, v in out.items()})


def evaluation_golden():
    """G12 — the test stage of estimate.py:79-122 executed with the reference QuantileRNN, utils.sliding_window and
    QuantileRNN.normalization_minmax on a synthetic series: per-window selection, clamp, de-normalisation, error
    percentiles and the console lines.  (estimate.py itself cannot be imported — it is a script that needs matplotlib —
    so its loop is re-typed here around the reference's own functions; the DataLoader is replaced by indexing.)"""
    import json
    from utils import sliding_window as ref_sliding_window
    M, F, N, W, split_frac = 4, 16, 900, 60, 0.40
    names = ["svc%d_%s" % (i // 2, ("cpu", "memory")[i % 2]) for i in range(M)]
    traffic = np.floor(synth.uniform(77, N * F).reshape(N, F) * 40.0)
    res = synth.uniform(78, N * M).reshape(N, M) * np.asarray([200.0, 3000.0, 50.0, 900.0]) + np.asarray([10.0, 500.0, 1.0, 100.0])
    X = ref_sliding_window(traffic, W)
    y = ref_sliding_window(res, W)
    split = int(len(X) * split_frac)
    X, xmin, xmax = QuantileRNN.normalization_minmax(X, split=split)
    scales = []
    for idx in range(M):
        y_, mn, mx = QuantileRNN.normalization_minmax(y[:, :, [idx]], split=split)
        y[:, :, [idx]] = y_
        scales.append((float(mx - mn), float(mn)))
    X_test, y_test = torch.Tensor(X[split:]), torch.Tensor(y[split:])
    model, blob = build_model(M, F, "synth", 1.5)
    model.eval()
    yerr = [[] for _ in names]
    losses, lines = [], []
    with torch.no_grad():
        num_cycles = 0
        for iv in range(len(X_test)):
            if iv % W != 0 or num_cycles >= 9:
                continue
            num_cycles += 1
            inputs, labels = X_test[iv:iv + 1], y_test[iv:iv + 1]
            outputs = model(inputs)
            losses.append(model.quantile_loss(outputs, labels).item())
            labels = labels.numpy()[0]
            outputs_deeprest = np.maximum(outputs.numpy(), 1e-6)[0]
            for idx in range(M):
                labels_ = labels[:, idx] * scales[idx][0] + scales[idx][1]
                outputs_ = outputs_deeprest[:, idx, 1] * scales[idx][0] + scales[idx][1]
                yerr[idx] += list(np.abs(outputs_ - labels_))
    for idx, name in enumerate(names):
        lines.append('===== %s =====' % name)
        lines.append('   DEEPR => Median: %.4f | 95-th: %.4f | 99-th: %.4f | Max: %.4f' % (
            float(np.median(yerr[idx])), np.percentile(yerr[idx], q=95), np.percentile(yerr[idx], q=99), np.max(yerr[idx])))
    out = {"M": M, "F": F, "N": N, "W": W, "split": split, "names": names, "wseed": WSEED, "wscale": 1.5,
           "traffic_seed": 77, "res_seed": 78, "xmin": float(xmin), "xmax": float(xmax), "scales": scales,
           "loss": float(np.mean(losses)), "n_windows": num_cycles,
           "summary": {n: [float(np.median(e)), float(np.percentile(e, 95)), float(np.percentile(e, 99)), float(np.max(e))]
                       for n, e in zip(names, yerr)},
           "lines": lines}
    with open(os.path.join(OUT, "g12_evaluation.json"), "w") as f:
        json.dump(out, f)
    print("g12_evaluation:", num_cycles, "windows, loss", out["loss"])
    print("\n".join(lines[:4]))


if __name__ == "__main__":
    main()
    evaluation_golden()
    featurize_golden()
    trained_golden()
    synthesizer_golden()
