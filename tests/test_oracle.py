"""Pin the numpy oracle against golden vectors minted by executing the reference
(oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, assert_parity, golden_cases, load_golden
from deeprest_b200 import layout, synth
from oracle import qrnn_numpy as oracle


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: os.path.basename(p)[:-4])
def test_forward_matches_reference(path):
    g = load_golden(path)
    if g["T"] > 512:
        pytest.skip("long horizon covered by test_long_horizon")
    out = oracle.forward(g["blob_arr"], g["x"], g["M"], g["F"])
    assert_parity(out, g["out"], what="oracle fp32 vs reference fp32")
    # and the fp32 reference sits on the fp64 error floor the survey measured
    assert np.abs(g["out"] - g["out64"]).max() < 2e-6


def test_long_horizon():
    g = load_golden(os.path.join(GOLDEN_DIR, "g3_long.npz"))
    out = oracle.forward(g["blob_arr"], g["x"], g["M"], g["F"])
    assert_parity(out, g["out"], what="T=1440")


@pytest.mark.parametrize("path", golden_cases()[:4], ids=lambda p: os.path.basename(p)[:-4])
def test_fp64_oracle_matches_fp64_reference(path):
    g = load_golden(path)
    if g["T"] > 512:
        pytest.skip("slow")
    out = oracle.forward(g["blob_arr"].astype(np.float64), g["x"], g["M"], g["F"], dtype=np.float64)
    assert np.abs(out - g["out64"]).max() < 1e-12


@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: os.path.basename(p)[:-4])
def test_quantile_loss(path):
    g = load_golden(path)
    y = synth.labels(int(g["yseed"]), g["B"], g["T"], g["M"])
    loss = oracle.quantile_loss(g["out"], y)
    assert abs(float(loss) - float(g["loss"])) <= 1e-6 * max(1.0, abs(float(g["loss"])))


def test_mask_sums_to_one_and_is_input_independent():
    blob = synth.weights(3, 2, 16)
    ex = layout.unpack_blob(blob, 2, 16)[1]
    m = oracle.feature_mask(ex)
    assert m.shape == (16,) and abs(m.sum() - 1) < 1e-6 and (m > 0).all()


def test_single_metric_rejected_like_reference():
    # the reference crashes for num_metrics == 1 (torch.stack([]) at qrnn.py:52)
    with pytest.raises(ValueError):
        oracle.forward(synth.weights(1, 1, 16), synth.windows(1, 1, 4, 16), 1, 16)


def test_window_and_minmax_helpers():
    g = np.load(os.path.join(GOLDEN_DIR, "g8_window_norm.npz"))
    win = oracle.sliding_window(g["ts"], int(g["window"]))
    assert win.shape == g["win"].shape and np.array_equal(win, g["win"])   # drops last window
    nm, lo, hi = oracle.normalization_minmax(win.copy(), int(g["split"]))
    assert lo == g["lo"] and hi == g["hi"] and np.array_equal(nm, g["norm"])
    const = np.full((4, 3), 2.0)
    same, _, _ = oracle.normalization_minmax(const, 2)
    assert np.array_equal(same, const)                                     # zero range: identity


def test_pinball_grad_matches_finite_differences():
    rng = np.random.default_rng(0)
    out = rng.standard_normal((2, 3, 2, 3))
    y = rng.standard_normal((2, 3, 2))
    g = oracle.quantile_loss_grad(out, y, dtype=np.float64)
    eps = 1e-6
    for idx in [(0, 0, 0, 0), (1, 2, 1, 2), (0, 1, 1, 1)]:
        p = out.copy(); p[idx] += eps
        m = out.copy(); m[idx] -= eps
        fd = (oracle.quantile_loss(p, y, dtype=np.float64) - oracle.quantile_loss(m, y, dtype=np.float64)) / (2 * eps)
        assert abs(fd - g[idx]) < 1e-6


def test_torch_cpu_baseline_port_matches_reference_golden():
    from oracle.qrnn_torch_cpu import TorchCpuPort
    g = load_golden(os.path.join(GOLDEN_DIR, "g2b_small_diurnal.npz"))
    out = TorchCpuPort(g["blob_arr"], g["M"], g["F"]).forward(g["x"])
    assert np.abs(out - g["out"]).max() < 1e-7     # same torch kernels as the reference


def _g5():
    g = dict(np.load(os.path.join(GOLDEN_DIR, "g5_train_step.npz")))
    M, B, T, F = (int(g[k]) for k in ("M", "B", "T", "F"))
    blob = synth.weights(int(g["wseed"]), M, F, float(g["wscale"]))
    assert abs(blob.astype(np.float64).sum() - float(g["blob_sum"])) < 1e-9
    x = synth.windows(int(g["xseed"]), B, T, F, str(g["xkind"]))
    y = synth.labels(int(g["yseed"]), B, T, M)
    dm = (synth.uniform(int(g["mask_seed"]), M * B * T * 2 * layout.H) >= 0.5).astype(np.float32)
    return g, M, B, T, F, blob, x, y, dm.reshape(M, B, T, 2 * layout.H)


def test_train_step_grads_match_reference_autograd():
    g, M, B, T, F, blob, x, y, dm = _g5()
    loss, out, grads = oracle.loss_and_grads(blob, x, y, M, F, dropout_masks=dm)
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    assert np.abs(out - g["out"]).max() < 1e-6
    ref = g["grads"]
    scale = np.abs(ref).max()
    assert np.abs(grads - ref).max() < 2e-6 * scale + 1e-9, np.abs(grads - ref).max()
    # per-tensor: every parameter family gets a gradient of the right size
    for name, (off, shape) in layout.expert_offsets(F).items():
        n = int(np.prod(shape))
        a, b = grads[off:off + n], ref[off:off + n]
        assert np.abs(a - b).max() <= 1e-5 * max(np.abs(b).max(), 1e-6) + 1e-9, name


def test_adam_step_matches_reference_optimizer():
    g, M, B, T, F, blob, x, y, dm = _g5()
    w, m, v = oracle.adam_step(blob, g["grads"], np.zeros_like(blob), np.zeros_like(blob), step=1, lr=float(g["lr"]))
    assert np.abs(w - g["weights_after"]).max() < 1e-7


def test_fp64_grads_match_finite_differences():
    M, B, T, F = 2, 2, 3, 3
    blob = synth.weights(9, M, F, 1.5).astype(np.float64)
    x = synth.windows(3, B, T, F).astype(np.float64)
    y = synth.labels(4, B, T, M).astype(np.float64)
    dm = (synth.uniform(8, M * B * T * 2 * layout.H) >= 0.5).astype(np.float64).reshape(M, B, T, 2 * layout.H)
    _, _, g = oracle.loss_and_grads(blob, x, y, M, F, dropout_masks=dm, dtype=np.float64)
    rng = np.random.default_rng(0)
    offs = layout.expert_offsets(F)
    for name in ("mask_w1", "mask_w2", "w_ih_f", "w_hh_r", "b_hh_f", "head_w", "head_b"):
        off, shape = offs[name]
        idx = layout.params_per_expert(F) * int(rng.integers(M)) + off + int(rng.integers(int(np.prod(shape))))
        eps = 1e-6
        p = blob.copy(); p[idx] += eps
        q = blob.copy(); q[idx] -= eps
        lp = oracle.quantile_loss(oracle.forward(p, x, M, F, np.float64, dm), y, dtype=np.float64)
        lq = oracle.quantile_loss(oracle.forward(q, x, M, F, np.float64, dm), y, dtype=np.float64)
        fd = (lp - lq) / (2 * eps)
        assert abs(fd - g[idx]) < 1e-6 * max(1.0, abs(fd)), (name, fd, g[idx])


def test_product_normalization_minmax_matches_reference_golden():
    """The product's own host helpers (estimator.sliding_window / QuantileRNN.normalization_minmax; utils.py:4-5,
    qrnn.py:69-75) against the reference-minted G8 — not only the oracle's copies."""
    from deeprest_b200.estimator import QuantileRNN, sliding_window
    g = np.load(os.path.join(GOLDEN_DIR, "g8_window_norm.npz"))
    win = sliding_window(g["ts"], int(g["window"]))
    assert win.shape == g["win"].shape and np.array_equal(win, g["win"])
    got, lo, hi = QuantileRNN.normalization_minmax(win.copy(), int(g["split"]))
    assert np.array_equal(got, g["norm"]) and lo == g["lo"] and hi == g["hi"]
    const = np.full((7, 3), 2.5)
    same, lo, hi = QuantileRNN.normalization_minmax(const, 4)          # span == 0: the reference returns M itself
    assert same is const and lo == 2.5 and hi == 2.5
