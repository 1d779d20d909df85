"""Training-step parity on the GPU: gradients of every parameter family, the loss and the
post-Adam weights against the golden minted from the reference's autograd (G5) and against the
numpy oracle at other shapes (incl. the micro-batched path)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from deeprest_b200 import QuantileRNN, layout, synth
from oracle import qrnn_numpy as oracle

pytestmark = pytest.mark.gpu


ENGINES = ["auto", "ffma"]       # auto: forward and backward recurrences on the tcgen05 kernels (split-fp16); ffma: exact fp32


def check_grads(got, ref, F, tag, engine="ffma", rows=1):
    scale = np.abs(ref).max()
    # the tensor-core forward saves activations that differ from exact fp32 by ~1e-7 relative (same arithmetic as the
    # inference engine, whose outputs meet the 1e-6 + 1e-4|ref| bar); the gradient bound is doubled for it.  Every
    # gradient is an fp32 reduction over the B*T rows in a different order than the oracle's: the bound grows with
    # sqrt(rows) beyond the ~100 rows of the small cases.
    overall = (5e-6 if engine == "ffma" else 1e-5) * max(1.0, (rows / 100.0) ** 0.5)
    assert np.abs(got - ref).max() <= overall * scale + 1e-9, f"{tag}: max grad err {np.abs(got - ref).max():.3e} (scale {scale:.3e})"
    pe = layout.params_per_expert(F)
    for name, (off, shape) in layout.expert_offsets(F).items():
        n = int(np.prod(shape))
        for e in range(got.size // pe):
            a, b = got[e * pe + off:e * pe + off + n], ref[e * pe + off:e * pe + off + n]
            tol = 2e-5 * max(1.0, (rows / 100.0) ** 0.5) * max(np.abs(b).max(), 1e-7) + 1e-9
            assert np.abs(a - b).max() <= tol, f"{tag}: {name}[expert {e}] err {np.abs(a - b).max():.3e} vs max {np.abs(b).max():.3e}"


@pytest.mark.parametrize("engine", ENGINES)
def test_train_step_matches_reference_golden(engine):
    g = dict(np.load(os.path.join(GOLDEN_DIR, "g5_train_step.npz")))
    M, B, T, F = (int(g[k]) for k in ("M", "B", "T", "F"))
    blob = synth.weights(int(g["wseed"]), M, F, float(g["wscale"]))
    x = synth.windows(int(g["xseed"]), B, T, F, str(g["xkind"]))
    y = synth.labels(int(g["yseed"]), B, T, M)
    dm = (synth.uniform(int(g["mask_seed"]), M * B * T * 2 * layout.H) >= 0.5).astype(np.uint8).reshape(M, B, T, 2 * layout.H)
    m = QuantileRNN(F, M, engine=engine)
    try:
        m.load_blob(blob)
        loss = m.train_step(x, y, lr=float(g["lr"]), dropout_mask=dm)
        assert m.last_engine == ("ffma" if engine == "ffma" else "tcgen05")     # the recurrences really ran there
        grads = m.grads()
        after = m.blob()
    finally:
        m.close()
    assert abs(loss - float(g["loss"])) < 2e-6
    check_grads(grads, g["grads"], F, "vs reference autograd", engine, rows=B * T)
    if engine == "ffma":
        assert np.abs(after - g["weights_after"]).max() < 2e-6, np.abs(after - g["weights_after"]).max()
    else:
        # Adam's first step moves a weight by lr*g/(|g|+eps): where |g| ~ eps a 1e-9 gradient difference legitimately changes
        # the update, so the tensor-core variant is checked on its own gradients (torch's Adam arithmetic, oracle.adam_step)
        ref_w, _, _ = oracle.adam_step(blob, grads, np.zeros_like(blob), np.zeros_like(blob), step=1, lr=float(g["lr"]))
        assert np.abs(after - ref_w).max() <= 3e-7 * max(1.0, np.abs(ref_w).max())
        big = np.abs(g["grads"]) > 1e-3 * np.abs(g["grads"]).max()        # and directly where the gradient is not tiny
        assert np.abs(after - g["weights_after"])[big].max() < 2e-6


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("M,B,T,F,mb", [(3, 5, 7, 5, 0), (2, 9, 4, 16, 4), (4, 6, 12, 33, 0), (2, 300, 5, 16, 0), (2, 300, 3, 8, 140), (2, 3, 4, 70, 0)])
def test_train_step_matches_oracle(M, B, T, F, mb, engine, monkeypatch):
    if mb:
        monkeypatch.setenv("DR_TRAIN_MICROBATCH", str(mb))        # force the multi-micro-batch path
    blob = synth.weights(40 + M, M, F, 1.5)
    x = synth.windows(3, B, T, F, "diurnal")
    y = synth.labels(4, B, T, M)
    dm = (synth.uniform(8, M * B * T * 2 * layout.H) >= 0.5).astype(np.uint8).reshape(M, B, T, 2 * layout.H)
    ref_loss, _, ref_g = oracle.loss_and_grads(blob, x, y, M, F, dropout_masks=dm.astype(np.float32))
    m = QuantileRNN(F, M, engine=engine)
    try:
        m.load_blob(blob)
        loss = m.train_step(x, y, lr=1e-3, dropout_mask=dm)
        grads = m.grads()
        after = m.blob()
        # the inference images follow the updated weights
        out_after = m.eval()(x)
    finally:
        m.close()
    assert abs(loss - float(ref_loss)) < 2e-6
    check_grads(grads, ref_g, F, "vs oracle", engine, rows=B * T)
    # Adam is checked on the GPU's own gradients: at step 1 the update is lr*g/(|g|+eps), so where |g| ~ eps a
    # 1e-9 gradient difference legitimately moves the weight by more than any fixed tolerance
    ref_w, _, _ = oracle.adam_step(blob, grads, np.zeros_like(blob), np.zeros_like(blob), step=1)
    assert np.abs(after - ref_w).max() <= 3e-7 * max(1.0, np.abs(ref_w).max())
    ref_out = oracle.forward(after, x, M, F)
    assert np.all(np.abs(out_after - ref_out) <= 1e-6 + 1e-4 * np.abs(ref_out))


def test_training_reduces_the_loss_with_device_rng():
    M, B, T, F = 2, 32, 20, 16
    blob = synth.weights(7, M, F)
    x = synth.windows(3, B, T, F, "diurnal")
    y = np.clip(x[:, :, :M] * 0.5 + 0.1, 0, 1).astype(np.float32)   # learnable target
    m = QuantileRNN(F, M, dropout=0.5)
    try:
        m.load_blob(blob)
        losses = [m.train_step(x, y, lr=1e-2, seed=100 + i) for i in range(30)]
    finally:
        m.close()
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < 0.7 * np.mean(losses[:5]), losses


# ---- bf16 training engine (dr_config.dtype = bf16; BASELINE configs[2], [4]) -------------------------------------------------
# Stated tolerance against the fp32 oracle (bf16 configs state their own): single-pass bf16 operands and bf16-stored
# activations give, measured on B200 over the shapes below (tools/train16_check.py, profiles/r02_train16_parity.log),
# loss |diff| <= 3e-5 and every gradient tensor within 0.2 % of its largest element.  Bounds: 5e-4 on the loss, 1e-2 of the
# per-tensor maximum on every gradient tensor (5e-3 of the global maximum overall), forecasts |a-b| <= 3e-3 + 1e-2|b|.
BF16_SHAPES = [(3, 5, 7, 5, 0), (2, 9, 4, 16, 0), (4, 6, 12, 33, 0), (2, 300, 5, 16, 0), (2, 300, 3, 8, 140), (2, 130, 40, 64, 0)]


def check_grads_bf16(got, ref, F, tag):
    assert np.isfinite(got).all(), tag
    assert np.abs(got - ref).max() <= 5e-3 * np.abs(ref).max(), f"{tag}: overall {np.abs(got - ref).max():.3e} vs max {np.abs(ref).max():.3e}"
    pe = layout.params_per_expert(F)
    for name, (off, shape) in layout.expert_offsets(F).items():
        n = int(np.prod(shape))
        for e in range(got.size // pe):
            a, b = got[e * pe + off:e * pe + off + n], ref[e * pe + off:e * pe + off + n]
            assert np.abs(a - b).max() <= 1e-2 * max(np.abs(b).max(), 1e-9) + 1e-10, \
                f"{tag}: {name}[expert {e}] err {np.abs(a - b).max():.3e} vs max {np.abs(b).max():.3e}"


@pytest.mark.parametrize("M,B,T,F,mb", BF16_SHAPES)
def test_bf16_train_step_matches_fp32_oracle(M, B, T, F, mb, monkeypatch):
    if mb:
        monkeypatch.setenv("DR_TRAIN_MICROBATCH", str(mb))        # two micro-batches of 256 + 44 windows
    blob = synth.weights(40 + M, M, F, 1.5)
    x = synth.windows(3, B, T, F, "diurnal")
    y = synth.labels(4, B, T, M)
    dm = (synth.uniform(8, M * B * T * 2 * layout.H) >= 0.5).astype(np.uint8).reshape(M, B, T, 2 * layout.H)
    ref_loss, ref_out, ref_g = oracle.loss_and_grads(blob, x, y, M, F, dropout_masks=dm.astype(np.float32))
    m = QuantileRNN(F, M, dtype="bf16")
    try:
        m.load_blob(blob)
        loss = m.train_step(x, y, lr=1e-3, dropout_mask=dm)
        assert m.last_engine == "tcgen05-bf16"
        grads = m.grads()
        after = m.blob()
    finally:
        m.close()
    assert abs(loss - float(ref_loss)) < 5e-4
    check_grads_bf16(grads, ref_g, F, "bf16 vs fp32 oracle")
    ref_w, _, _ = oracle.adam_step(blob, grads, np.zeros_like(blob), np.zeros_like(blob), step=1)      # Adam stays fp32
    assert np.abs(after - ref_w).max() <= 3e-7 * max(1.0, np.abs(ref_w).max())


def test_bf16_train_step_matches_reference_golden():
    g = dict(np.load(os.path.join(GOLDEN_DIR, "g5_train_step.npz")))
    M, B, T, F = (int(g[k]) for k in ("M", "B", "T", "F"))
    blob = synth.weights(int(g["wseed"]), M, F, float(g["wscale"]))
    x = synth.windows(int(g["xseed"]), B, T, F, str(g["xkind"]))
    y = synth.labels(int(g["yseed"]), B, T, M)
    dm = (synth.uniform(int(g["mask_seed"]), M * B * T * 2 * layout.H) >= 0.5).astype(np.uint8).reshape(M, B, T, 2 * layout.H)
    m = QuantileRNN(F, M, dtype="bf16")
    try:
        m.load_blob(blob)
        loss = m.train_step(x, y, lr=float(g["lr"]), dropout_mask=dm)
        grads = m.grads()
    finally:
        m.close()
    assert abs(loss - float(g["loss"])) < 5e-4
    check_grads_bf16(grads, g["grads"], F, "bf16 vs reference autograd")


def test_bf16_long_horizon_training_parity():
    """BASELINE configs[4]'s horizon (seq_len 1440) at a small expert count: the bf16 recurrences stay within the stated
    tolerance over 1440 steps in both directions."""
    M, B, T, F = 2, 3, 1440, 64
    blob = synth.weights(17, M, F, 1.0)
    x = synth.windows(5, B, T, F, "diurnal")
    y = synth.labels(6, B, T, M)
    dm = (synth.uniform(9, M * B * T * 2 * layout.H) >= 0.5).astype(np.uint8).reshape(M, B, T, 2 * layout.H)
    ref_loss, _, ref_g = oracle.loss_and_grads(blob, x, y, M, F, dropout_masks=dm.astype(np.float32))
    m = QuantileRNN(F, M, dtype="bf16")
    try:
        m.load_blob(blob)
        loss = m.train_step(x, y, lr=1e-3, dropout_mask=dm)
        grads = m.grads()
    finally:
        m.close()
    assert abs(loss - float(ref_loss)) < 5e-4
    check_grads_bf16(grads, ref_g, F, "T=1440 bf16 vs fp32 oracle")


def test_fp32_long_horizon_training_parity():
    """same horizon on the fp32-parity (split-fp16) engine, fp32 bounds"""
    M, B, T, F = 2, 3, 1440, 64
    blob = synth.weights(17, M, F, 1.0)
    x = synth.windows(5, B, T, F, "diurnal")
    y = synth.labels(6, B, T, M)
    dm = (synth.uniform(9, M * B * T * 2 * layout.H) >= 0.5).astype(np.uint8).reshape(M, B, T, 2 * layout.H)
    ref_loss, _, ref_g = oracle.loss_and_grads(blob, x, y, M, F, dropout_masks=dm.astype(np.float32))
    m = QuantileRNN(F, M)
    try:
        m.load_blob(blob)
        loss = m.train_step(x, y, lr=1e-3, dropout_mask=dm)
        grads = m.grads()
    finally:
        m.close()
    assert abs(loss - float(ref_loss)) < 2e-6
    check_grads(grads, ref_g, F, "T=1440 fp32 engine vs oracle", "auto", rows=B * T)


def test_bf16_training_reduces_the_loss_with_device_rng():
    M, B, T, F = 2, 32, 20, 16
    blob = synth.weights(7, M, F)
    x = synth.windows(3, B, T, F, "diurnal")
    y = np.clip(x[:, :, :M] * 0.5 + 0.1, 0, 1).astype(np.float32)
    m = QuantileRNN(F, M, dropout=0.5, dtype="bf16")
    try:
        m.load_blob(blob)
        losses = [m.train_step(x, y, lr=1e-2) for i in range(30)]       # constant seed: the step counter varies the mask
    finally:
        m.close()
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < 0.7 * np.mean(losses[:5]), losses


def test_train_mode_forecasts_of_the_bf16_step_match_oracle():
    import torch
    M, B, T, F = 4, 6, 12, 33
    blob = synth.weights(44, M, F, 1.5)
    x = synth.windows(3, B, T, F, "diurnal")
    y = synth.labels(4, B, T, M)
    dm = (synth.uniform(8, M * B * T * 2 * layout.H) >= 0.5).astype(np.uint8).reshape(M, B, T, 2 * layout.H)
    _, ref_out, _ = oracle.loss_and_grads(blob, x, y, M, F, dropout_masks=dm.astype(np.float32))
    m = QuantileRNN(F, M, dtype="bf16")
    try:
        m.load_blob(blob)
        m.train_step(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), dropout_mask=dm)     # device-resident entry point
        out = m.train_outputs.cpu().numpy()
    finally:
        m.close()
    assert np.all(np.abs(out - ref_out) <= 3e-3 + 1e-2 * np.abs(ref_out)), np.abs(out - ref_out).max()
