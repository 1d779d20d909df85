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
