"""ORACLE (test infrastructure, NOT product code) — numpy restatement of the
DeepRest ``QuantileRNN`` hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import this module.  The product path (``deeprest_b200``) never does and
fails loudly when its CUDA library is missing.

Parity pinning: the reference ships NO tests or golden vectors for this path
(SURVEY §4), so by the reference's own fixtures this oracle is **parity
unpinned**.  It is pinned instead against outputs of the reference itself —
``/root/reference/resource-estimation/qrnn.py`` executed on torch 2.11.0 CPU in
the authoring container by ``oracle/make_golden.py``; the resulting vectors are
committed under ``tests/golden/`` and ``tests/test_oracle.py`` checks this
restatement against every one of them.

Each function cites the reference lines it restates.  All arithmetic runs in
``dtype`` (float32 to mirror the reference, float64 for the error floor).
"""
from __future__ import annotations

import numpy as np

from deeprest_b200.layout import H, Q, QUANTILES, unpack_blob


def _sigmoid(v):
    return 1.0 / (1.0 + np.exp(-v))


def feature_mask(ex, dtype=np.float32):
    """qrnn.py:34 — ``softmax(L2(relu(L1(mask_init))))`` with mask_init == 1.

    Input independent: one [F] vector per expert, sums to 1.
    """
    w1 = ex["mask_w1"].astype(dtype)[:, 0]
    hid = np.maximum(w1 * dtype(1.0) + ex["mask_b1"].astype(dtype), 0)       # Linear(1,H)+ReLU
    logits = ex["mask_w2"].astype(dtype) @ hid + ex["mask_b2"].astype(dtype)  # Linear(H,F)
    logits = logits - logits.max()
    e = np.exp(logits)
    return (e / e.sum()).astype(dtype)


def gru_direction(xm, w_ih, w_hh, b_ih, b_hh, reverse, dtype=np.float32, keep=None):
    """One direction of ``nn.GRU`` (qrnn.py:24,39-41; torch nn/modules/rnn.py GRU eqs).

    xm [T,B,F] → [T,B,H].  h0 = 0 (qrnn.py:39).  Gate order (r,z,n):
      r = σ(W_ir x + b_ir + W_hr h + b_hr);  z likewise;
      n = tanh(W_in x + b_in + r ⊙ (W_hn h + b_hn));  h' = (1−z) ⊙ n + z ⊙ h.
    ``keep`` (dict) optionally receives the per-step tensors the backward needs.
    """
    T, B, _ = xm.shape
    h = np.zeros((B, H), dtype)
    out = np.empty((T, B, H), dtype)
    gi_all = xm.reshape(T * B, -1) @ w_ih.T.astype(dtype) + b_ih.astype(dtype)
    gi_all = gi_all.reshape(T, B, 3 * H)
    w_hh_t = np.ascontiguousarray(w_hh.T.astype(dtype))
    b_hh = b_hh.astype(dtype)
    if keep is not None:
        keep.update(r=np.empty((T, B, H), dtype), z=np.empty((T, B, H), dtype),
                    n=np.empty((T, B, H), dtype), q=np.empty((T, B, H), dtype),
                    hprev=np.empty((T, B, H), dtype))
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        gi = gi_all[t]
        gh = h @ w_hh_t + b_hh
        r = _sigmoid(gi[:, :H] + gh[:, :H])
        z = _sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
        n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
        if keep is not None:
            keep["r"][t], keep["z"][t], keep["n"][t] = r, z, n
            keep["q"][t], keep["hprev"][t] = gh[:, 2 * H:], h
        h = ((1.0 - z) * n + z * h).astype(dtype)
        out[t] = h
    return out


def expert_rnn_out(ex, x, dtype=np.float32, keep=None):
    """qrnn.py:33-42 for one expert: mask, masked input, bi-GRU → r̃ [B,T,2H]."""
    mask = feature_mask(ex, dtype)
    xm = (x.astype(dtype) * mask[None, None, :]).transpose(1, 0, 2)   # qrnn.py:36-37
    kf = {} if keep is not None else None
    kr = {} if keep is not None else None
    fwd = gru_direction(xm, ex["w_ih_f"], ex["w_hh_f"], ex["b_ih_f"], ex["b_hh_f"], False, dtype, kf)
    rev = gru_direction(xm, ex["w_ih_r"], ex["w_hh_r"], ex["b_ih_r"], ex["b_hh_r"], True, dtype, kr)
    if keep is not None:
        keep.update(mask=mask, xm=xm, fwd=kf, rev=kr)
    return np.concatenate([fwd, rev], axis=-1).transpose(1, 0, 2)      # qrnn.py:42


def forward(blob, x, M, F, dtype=np.float32, dropout_masks=None, dropout_p=0.5, keep=None):
    """qrnn.py:28-56 — QuantileRNN.forward.  x [B,T,F] → out [B,T,M,Q].

    ``dropout_masks`` (optional, [M,B,T,2H] of 0/1) replays ``nn.Dropout``
    (qrnn.py:43): kept entries are scaled by 1/(1-p); ``None`` == eval mode.
    The cross-expert mean is formed exactly as the reference does — a mean over
    the M−1 *other* experts' outputs (qrnn.py:46-52) — not via the sum trick the
    CUDA path uses, so the two are independent derivations.
    """
    if M < 2:
        raise ValueError("the reference needs num_metrics >= 2 (torch.stack([]) at qrnn.py:52)")
    experts = unpack_blob(np.asarray(blob), M, F)
    x = np.asarray(x)
    rnn_outs = []
    for e, ex in enumerate(experts):
        ke = {} if keep is not None else None
        r = expert_rnn_out(ex, x, dtype, ke)
        if dropout_masks is not None:
            r = r * dropout_masks[e].astype(dtype) * dtype(1.0 / (1.0 - dropout_p))
        if keep is not None:
            keep.setdefault("experts", []).append(ke)
        rnn_outs.append(r.astype(dtype))
    preds = []
    for i, ex in enumerate(experts):
        others = np.stack([rnn_outs[j] for j in range(M) if j != i])
        m = others.mean(axis=0, dtype=dtype)                              # qrnn.py:52
        cat = np.concatenate([m, rnn_outs[i]], axis=-1)                   # qrnn.py:53
        preds.append(cat @ ex["head_w"].T.astype(dtype) + ex["head_b"].astype(dtype))  # qrnn.py:54
    if keep is not None:
        keep["rnn_outs"] = rnn_outs
    return np.stack(preds).transpose(1, 2, 0, 3).astype(dtype)            # qrnn.py:55


def quantile_loss(out, y, quantiles=QUANTILES, dtype=np.float32):
    """qrnn.py:58-67 — pinball loss: mean over M of mean_{B,T} Σ_q max((q−1)e, q·e)."""
    out = np.asarray(out, dtype)
    y = np.asarray(y, dtype)
    M = out.shape[2]
    per_metric = []
    for idx in range(M):
        tot = np.zeros(out.shape[:2], dtype)
        for i, q in enumerate(quantiles):
            err = y[:, :, idx] - out[:, :, idx, i]
            tot = tot + np.maximum(dtype(q - 1.0) * err, dtype(q) * err)
        per_metric.append(tot.mean(dtype=dtype))
    return dtype(np.mean(np.asarray(per_metric, dtype), dtype=dtype))


def quantile_loss_grad(out, y, quantiles=QUANTILES, dtype=np.float32):
    """dL/dout of :func:`quantile_loss` as torch autograd produces it.

    ``torch.max(a, b)`` splits the gradient 50/50 on ties:
    d/dŷ = (1−q) if e<0, −q if e>0, (0.5−q) if e==0, all × 1/(M·B·T).
    """
    out = np.asarray(out, dtype)
    y = np.asarray(y, dtype)
    B, T, M, _ = out.shape
    g = np.empty_like(out)
    for i, q in enumerate(quantiles):
        err = y - out[..., i]
        g[..., i] = np.where(err < 0, 1.0 - q, np.where(err > 0, -q, 0.5 - q))
    return (g / dtype(M * B * T)).astype(dtype)


def normalization_minmax(Mx, split):
    """qrnn.py:69-75 — min-max over the train split; identity if constant."""
    lo = np.min(Mx[:split])
    hi = np.max(Mx[:split])
    if (hi - lo) != 0.0:
        Mx = (Mx - lo) / (hi - lo)
    return Mx, lo, hi


def sliding_window(ts, window):
    """utils.py:4-5 — stride-1 windows; NOTE drops the final window (range(len−W))."""
    return np.asarray([ts[i:i + window] for i in range(len(ts) - window)])


# ---------------------------------------------------------------------------------------------
# Training step: what ``loss.backward(); optimizer.step()`` computes (estimate.py:70-74).
# The reference has no backward code — it is torch autograd over qrnn.py:28-67 — so this is a
# restatement of those adjoints (SURVEY §8a "Backward"), pinned by tests/golden/g5_train_step.npz.
# ---------------------------------------------------------------------------------------------

def _gru_direction_backward(keep, d_out, w_ih, w_hh, reverse, dtype):
    """Adjoint of :func:`gru_direction`.  d_out [T,B,H] is dL/d(output_t).
    Returns dW_ih, dW_hh, db_ih, db_hh, dxm [T,B,F]."""
    T, B, _ = d_out.shape
    F = w_ih.shape[1]
    w_ih = w_ih.astype(dtype); w_hh = w_hh.astype(dtype)
    dW_ih = np.zeros_like(w_ih); dW_hh = np.zeros_like(w_hh)
    db_ih = np.zeros(3 * H, dtype); db_hh = np.zeros(3 * H, dtype)
    dxm = np.zeros((T, B, F), dtype)
    dh = np.zeros((B, H), dtype)
    xm = keep["xm"]
    steps = range(T) if reverse else range(T - 1, -1, -1)     # opposite to the forward order
    for t in steps:
        r, z, n, q, hp = keep["r"][t], keep["z"][t], keep["n"][t], keep["q"][t], keep["hprev"][t]
        dh = dh + d_out[t]
        dn = dh * (1.0 - z)
        dz = dh * (hp - n)
        da_n = dn * (1.0 - n * n)
        dr = da_n * q
        dq = da_n * r
        da_z = dz * z * (1.0 - z)
        da_r = dr * r * (1.0 - r)
        dgi = np.concatenate([da_r, da_z, da_n], axis=1)
        dgh = np.concatenate([da_r, da_z, dq], axis=1)
        dW_hh += dgh.T @ hp
        db_hh += dgh.sum(0)
        dW_ih += dgi.T @ xm[t]
        db_ih += dgi.sum(0)
        dxm[t] = dgi @ w_ih
        dh = dh * z + dgh @ w_hh
    return dW_ih, dW_hh, db_ih, db_hh, dxm


def loss_and_grads(blob, x, y, M, F, dropout_masks=None, dropout_p=0.5, quantiles=QUANTILES, dtype=np.float32):
    """One forward+backward of the reference training step.  Returns (loss, out, grad_blob)."""
    from deeprest_b200.layout import expert_offsets, params_per_expert
    blob = np.asarray(blob)
    experts = unpack_blob(blob, M, F)
    keep = {}
    out = forward(blob, x, M, F, dtype, dropout_masks, dropout_p, keep)
    loss = quantile_loss(out, y, quantiles, dtype)
    dout = quantile_loss_grad(out, y, quantiles, dtype)                       # [B,T,M,Q]
    rt = keep["rnn_outs"]                                                     # r~_j (after dropout)  [B,T,2H]
    pe = params_per_expert(F)
    offs = expert_offsets(F)
    gblob = np.zeros(M * pe, dtype)

    def put(e, name, g):
        o, shape = offs[name]
        gblob[e * pe + o: e * pe + o + g.size] = g.astype(dtype).reshape(-1)

    d_rt = [np.zeros_like(rt[0]) for _ in range(M)]
    Bsz, T = out.shape[:2]
    for i, ex in enumerate(experts):                                          # head adjoint, qrnn.py:46-54
        W = ex["head_w"].astype(dtype)
        m_i = np.stack([rt[j] for j in range(M) if j != i]).mean(axis=0, dtype=dtype)
        cat = np.concatenate([m_i, rt[i]], axis=-1).reshape(Bsz * T, 4 * H)
        dy = dout[:, :, i, :].reshape(Bsz * T, Q)
        put(i, "head_w", dy.T @ cat)
        put(i, "head_b", dy.sum(0))
        dcat = (dy @ W).reshape(Bsz, T, 4 * H)
        d_rt[i] += dcat[..., 2 * H:]
        share = dcat[..., :2 * H] / dtype(M - 1)
        for j in range(M):
            if j != i:
                d_rt[j] += share
    for e, ex in enumerate(experts):
        ke = keep["experts"][e]
        d_r = d_rt[e]
        if dropout_masks is not None:
            d_r = d_r * dropout_masks[e].astype(dtype) * dtype(1.0 / (1.0 - dropout_p))
        d_r = d_r.transpose(1, 0, 2)                                          # [T,B,2H]
        dxm_tot = 0
        for dname, half, rev in (("f", slice(0, H), False), ("r", slice(H, 2 * H), True)):
            kd = dict(ke["fwd" if not rev else "rev"]); kd["xm"] = ke["xm"]
            dWi, dWh, dbi, dbh, dxm = _gru_direction_backward(
                kd, np.ascontiguousarray(d_r[..., half]), ex["w_ih_" + dname], ex["w_hh_" + dname], rev, dtype)
            put(e, "w_ih_" + dname, dWi); put(e, "w_hh_" + dname, dWh)
            put(e, "b_ih_" + dname, dbi); put(e, "b_hh_" + dname, dbh)
            dxm_tot = dxm_tot + dxm
        # x*mask (qrnn.py:36): dmask_f = sum_{b,t} x[b,t,f] * dxm[t,b,f]; then softmax / ReLU / Linear adjoints (qrnn.py:34)
        dmask = (x.astype(dtype).transpose(1, 0, 2) * dxm_tot).sum(axis=(0, 1))
        mask = ke["mask"]
        dlogit = mask * (dmask - (mask * dmask).sum())
        w1 = ex["mask_w1"].astype(dtype)[:, 0]
        pre = w1 + ex["mask_b1"].astype(dtype)
        hid = np.maximum(pre, 0)
        put(e, "mask_w2", np.outer(dlogit, hid)); put(e, "mask_b2", dlogit)
        dhid = (ex["mask_w2"].astype(dtype).T @ dlogit) * (pre > 0)
        put(e, "mask_w1", dhid); put(e, "mask_b1", dhid)                       # input is the constant 1
    return loss, out, gblob


def adam_step(w, g, m, v, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, dtype=np.float32):
    """torch.optim.Adam defaults (estimate.py:61), in torch's own operation order:
    p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps).  Returns (w, m, v)."""
    w = np.asarray(w, dtype); g = np.asarray(g, dtype)
    m = (dtype(beta1) * m + dtype(1 - beta1) * g).astype(dtype)
    v = (dtype(beta2) * v + dtype(1 - beta2) * g * g).astype(dtype)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = (np.sqrt(v) / dtype(np.sqrt(bc2)) + dtype(eps)).astype(dtype)
    w = (w - dtype(lr / bc1) * (m / denom)).astype(dtype)
    return w, m, v
