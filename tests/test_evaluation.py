"""N2 — the evaluation stage (estimate.py:79-122) against a golden minted by executing the reference model,
window helper and normaliser (oracle/make_golden.py::evaluation_golden).  The CPU test drives the host logic with an
oracle-backed stand-in for the estimator; the GPU test runs the real one."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from deeprest_b200 import evaluation, synth
from oracle import qrnn_numpy as oracle

G = json.load(open(os.path.join(GOLDEN_DIR, "g12_evaluation.json")))


def _inputs():
    N, F, M = G["N"], G["F"], G["M"]
    traffic = np.floor(synth.uniform(G["traffic_seed"], N * F).reshape(N, F) * 40.0)
    res = synth.uniform(G["res_seed"], N * M).reshape(N, M) * np.asarray([200.0, 3000.0, 50.0, 900.0]) + np.asarray([10.0, 500.0, 1.0, 100.0])
    traffic_n = (traffic - G["xmin"]) / (G["xmax"] - G["xmin"])            # qrnn.py:69-75 with the golden's min / max
    res_n = np.stack([(res[:, i] - G["scales"][i][1]) / G["scales"][i][0] for i in range(M)], axis=1)
    return traffic_n, res_n


class _OracleModel:
    """forward_series / quantile_loss with the oracle's arithmetic (CPU stand-in for the estimator)."""

    def __init__(self):
        self.blob = synth.weights(G["wseed"], G["M"], G["F"], G["wscale"])

    def forward_series(self, series, window, stride=1):
        n = len(series) - window
        starts = range(0, max(n, 0), stride)
        x = np.stack([series[s:s + window] for s in starts]).astype(np.float32)
        return oracle.forward(self.blob, x, G["M"], G["F"])

    def quantile_loss(self, out, y):
        return oracle.quantile_loss(out, y)


def _check(result):
    assert len(result["offsets"]) == G["n_windows"] == evaluation.MAX_CYCLES
    assert result["offsets"][0] == G["split"] and result["offsets"][1] - result["offsets"][0] == G["W"]
    assert abs(result["loss"] - G["loss"]) <= 1e-5 * abs(G["loss"])
    for name in G["names"]:
        s = result["summary"][name]
        got = [s["median"], s["p95"], s["p99"], s["max"]]
        np.testing.assert_allclose(got, G["summary"][name], rtol=2e-4, atol=1e-4)   # de-normalised units (up to 3000)
    text = evaluation.report(G["names"], result["summary"])
    assert text.count("=====") == 2 * len(G["names"])
    for line_ref, line in zip(G["lines"], text.split("\n")):
        if line_ref.startswith("====="):
            assert line == line_ref
        else:                                                            # same format; digits agree to the parity bar
            assert line[:12] == line_ref[:12] == "   DEEPR => " and len(line.split("|")) == 4


def test_window_selection_follows_the_reference_loader():
    W = 60
    assert evaluation.eval_window_offsets(900, 336, W) == [336 + k * W for k in range(9)]     # capped at nine
    assert evaluation.eval_window_offsets(700, 256, W) == [256 + k * W for k in range(7)]     # 384 test windows -> 7
    assert evaluation.eval_window_offsets(100, 50, W) == []                                   # no test window at all
    assert evaluation.eval_window_offsets(171, 50, W) == [50, 110]                            # X has 111 windows: iv = 0, 60


def test_evaluation_stage_matches_reference_golden_with_oracle_model():
    traffic_n, res_n = _inputs()
    _check(evaluation.evaluate(_OracleModel(), traffic_n, res_n, G["names"], G["scales"], G["split"], G["W"]))


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["tcgen05", "ffma"])
def test_evaluation_stage_on_the_gpu_matches_reference_golden(engine):
    from deeprest_b200 import QuantileRNN
    traffic_n, res_n = _inputs()
    model = QuantileRNN(G["F"], G["M"], engine=engine).eval()
    model.load_blob(synth.weights(G["wseed"], G["M"], G["F"], G["wscale"]))
    _check(evaluation.evaluate(model, traffic_n, res_n, G["names"], G["scales"], G["split"], G["W"]))
    model.close()
