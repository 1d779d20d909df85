"""GPU parity tests proper: the CUDA path, called through the C ABI, against the numpy oracle
and the golden vectors minted from the reference.  Run on the B200 box: pytest -m gpu."""
import os

import numpy as np
import pytest

from conftest import ATOL, GOLDEN_DIR, RTOL, assert_parity, golden_cases, load_golden
from deeprest_b200 import QuantileRNN, layout, synth
from oracle import qrnn_numpy as oracle

pytestmark = pytest.mark.gpu

ENGINES = ["ffma", "tcgen05"]


def make_model(M, F, blob, engine):
    from deeprest_b200 import _lib
    if not _lib.load().dr_has_engine(_lib.ENGINES[engine]):
        pytest.skip(f"{engine} engine is not in this build")
    if engine == "tcgen05" and F > 64:
        pytest.skip("tcgen05 engine covers input_size <= 64 (one 64-wide K block); larger F runs on the FFMA engine")
    m = QuantileRNN(input_size=F, num_metrics=M, engine=engine).eval()
    m.load_blob(blob)
    return m


def run(M, F, blob, x, engine):
    m = make_model(M, F, blob, engine)
    try:
        out = m(x)
        if engine != "auto":
            assert m.last_engine == engine
        return out
    finally:
        m.close()


def tc_shape_ok(B, T, F):
    return True


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("path", golden_cases(), ids=lambda p: os.path.basename(p)[:-4])
def test_forward_matches_reference_golden(path, engine):
    g = load_golden(path)
    out = run(g["M"], g["F"], g["blob_arr"], g["x"], engine)
    assert_parity(out, g["out"], what=f"{engine} vs reference golden")
    print(f"{os.path.basename(path)} [{engine}]: MAE vs reference {np.abs(out - g['out']).mean():.3e} "
          f"max {np.abs(out - g['out']).max():.3e}")


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("M,B,T,F,scale", [
    (2, 1, 1, 1, 1.0),        # degenerate: one window, one step, one feature
    (2, 1, 64, 16, 1.0),      # BASELINE configs[0]
    (3, 17, 5, 7, 2.0),       # ragged everything
    (2, 33, 9, 16, 1.0),      # FFMA 64-row tile
    (5, 70, 12, 20, 1.5),     # FFMA 128-row tile with padding rows; F not multiple of 16
    (4, 129, 6, 64, 1.0),     # crosses a 128-row tile
    (2, 256, 4, 100, 1.0),    # F > 64
    (130, 3, 4, 6, 1.0),      # many experts: 390 head columns = two TMEM column groups, last chunk partial
])
def test_forward_matches_oracle(M, B, T, F, scale, engine):
    blob = synth.weights(100 + M + F, M, F, scale)
    x = synth.windows(7 + B, B, T, F, "diurnal")
    ref = oracle.forward(blob, x, M, F)
    out = run(M, F, blob, x, engine)
    assert_parity(out, ref, what=f"{engine} vs oracle M{M} B{B} T{T} F{F}")


def test_mask_and_cross_expert_sum_match_oracle():
    M, B, T, F = 4, 8, 16, 16
    blob = synth.weights(5, M, F, 2.0)
    x = synth.windows(9, B, T, F)
    m = make_model(M, F, blob, "ffma")
    try:
        m(x)
        mask = m.debug_read("mask", M * F).reshape(M, F)
        Bp = (B + 127) // 128 * 128          # S is stored k-group major: [T][2H/4][Bp][4]
        S = m.debug_read("S", T * 2 * layout.H * Bp).reshape(T, 2 * layout.H // 4, Bp, 4)
        S = S[:, :, :B, :].transpose(2, 0, 1, 3).reshape(B, T, 2 * layout.H)
    finally:
        m.close()
    experts = layout.unpack_blob(blob, M, F)
    ref_mask = np.stack([oracle.feature_mask(ex) for ex in experts])
    assert np.abs(mask - ref_mask).max() < 1e-6 * ref_mask.max() + 1e-8
    ref_S = sum(oracle.expert_rnn_out(ex, x) for ex in experts)
    assert_parity(S, ref_S, rtol=1e-5, atol=2e-6, what="S = sum of GRU outputs")


def test_quantile_loss_matches_oracle_and_golden():
    g = load_golden(os.path.join(GOLDEN_DIR, "g2_small.npz"))
    y = synth.labels(int(g["yseed"]), g["B"], g["T"], g["M"])
    m = make_model(g["M"], g["F"], g["blob_arr"], "ffma")
    try:
        loss = m.quantile_loss(g["out"], y)
    finally:
        m.close()
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    assert abs(float(loss) - float(oracle.quantile_loss(g["out"], y))) < 1e-6


def test_weights_round_trip_through_the_abi():
    M, F = 3, 9
    blob = synth.weights(1, M, F)
    m = make_model(M, F, blob, "ffma")
    try:
        assert np.array_equal(m.blob(), blob)
        sd = m.state_dict()
        assert list(sd)[1:] == layout.state_dict_keys(M)
    finally:
        m.close()


def test_error_behaviour_matches_reference():
    with pytest.raises(RuntimeError):              # reference: torch.stack([]) at qrnn.py:52
        QuantileRNN(input_size=4, num_metrics=1)
    m = QuantileRNN(input_size=4, num_metrics=2).eval()
    try:
        with pytest.raises(Exception):              # forward before weights
            m(np.zeros((1, 2, 4), np.float32))
        m.load_blob(synth.weights(1, 2, 4))
        with pytest.raises(ValueError):             # wrong feature count (torch would raise too)
            m(np.zeros((1, 2, 5), np.float32))
    finally:
        m.close()


@pytest.mark.parametrize("engine", ENGINES)
def test_size_independent_properties_at_scale(engine):
    """Properties that need no oracle, at a size the oracle would take minutes for:
    windows are independent (eval mode), experts are permutation-equivariant, and the
    device-pointer path agrees with the host-pointer path."""
    import torch
    M, B, T, F = 16, 384, 96, 64
    blob = synth.weights(3, M, F, 1.5)
    x = synth.windows(11, B, T, F, "diurnal")
    m = make_model(M, F, blob, engine)
    try:
        full = m(x)
        # (1) window independence: any sub-batch gives the same rows
        part = m(x[100:230])
        assert_parity(part, full[100:230], rtol=1e-5, atol=5e-7, what="window independence")
        # (2) device path == host path
        dev = m(torch.from_numpy(x).cuda()).cpu().numpy()
        assert_parity(dev, full, rtol=1e-5, atol=5e-7, what="device vs host entry point")
    finally:
        m.close()
    # (3) expert permutation equivariance
    perm = np.random.default_rng(0).permutation(M)
    pe = layout.params_per_expert(F)
    blob_p = blob.reshape(M, pe)[perm].reshape(-1)
    m2 = make_model(M, F, blob_p, engine)
    try:
        out_p = m2(x[:64])
    finally:
        m2.close()
    assert_parity(out_p, full[:64][:, :, perm, :], rtol=1e-5, atol=5e-7, what="expert permutation")
    # (4) and the first windows agree with the oracle (bounded oracle sample)
    ref = oracle.forward(blob, x[:4], M, F)
    assert_parity(full[:4], ref, what="oracle sample at scale")


def test_engines_agree_at_config2_shape_sample():
    """BASELINE configs[1] shape (T=288, F=64) on a reduced expert/window count: the tensor-core
    engine against the exact-fp32 FFMA engine."""
    M, B, T, F = 8, 256, 288, 64
    blob = synth.weights(21, M, F, 1.0)
    x = synth.windows(2021, B, T, F)
    a = run(M, F, blob, x, "ffma")
    b = run(M, F, blob, x, "tcgen05")
    assert_parity(b, a, what="tcgen05 vs ffma engine")
    print(f"engines: MAE {np.abs(a - b).mean():.3e} max {np.abs(a - b).max():.3e}")


@pytest.mark.parametrize("engine", ENGINES)
def test_forward_from_raw_series_matches_windowed_forward(engine):
    """N1: on-device windowing == sliding_window (utils.py:4-5, last window dropped) + forward."""
    M, F, W = 3, 12, 20
    blob = synth.weights(17, M, F, 1.5)
    series = synth.uniform(99, 150 * F).reshape(150, F)
    m = make_model(M, F, blob, engine)
    try:
        for stride in (1, 7, 20):
            win = oracle.sliding_window(series, W)[::stride]            # estimate.py:85-86 keeps every stride-th window
            ref = oracle.forward(blob, win, M, F)
            out = m.forward_series(series, W, stride)
            assert out.shape == ref.shape
            assert_parity(out, ref, what=f"series stride {stride}")
        assert m.forward_series(series[:W], W).shape[0] == 0             # N - W == 0: no window, like the reference
    finally:
        m.close()


def test_fused_clamp_and_denormalisation():
    """N2: estimate.py:96 (clamp at 1e-6 on normalised outputs) + :101-102 (x*range + min) in the head epilogue."""
    M, B, T, F = 4, 6, 9, 8
    blob = synth.weights(23, M, F, 2.0)
    x = synth.windows(5, B, T, F, "diurnal")
    scales = [(3.5, 0.25), (120.0, 7.0), (1.0, 0.0), (0.01, -2.0)]
    m = make_model(M, F, blob, "tcgen05")
    try:
        plain = m(x)
        m.set_denormalization(scales)
        fused = m(x)
        m.set_denormalization(None)
        again = m(x)
    finally:
        m.close()
    ref = np.maximum(plain, 1e-6) * np.array([a for a, _ in scales], np.float32)[None, None, :, None] \
        + np.array([b for _, b in scales], np.float32)[None, None, :, None]
    assert np.abs(fused - ref).max() <= 1e-5 * np.abs(ref).max()
    assert np.abs(again - plain).max() < 1e-6


@pytest.mark.parametrize("engine,M,B,T,F", [("tcgen05", 3, 700, 3, 20), ("ffma", 2, 600, 2, 100), ("tcgen05", 2, 1100, 2, 8)])
def test_chunked_host_entry_point_matches_oracle(engine, M, B, T, F):
    """dr_forward splits B >= 512 into 2-4 chunks on two compute streams + a copy stream; ragged last chunk."""
    blob = synth.weights(77, M, F, 1.5)
    x = synth.windows(12, B, T, F, "diurnal")
    ref = oracle.forward(blob, x, M, F)
    out = run(M, F, blob, x, engine)
    assert_parity(out, ref, what=f"chunked host path {engine} B={B}")
