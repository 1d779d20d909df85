"""The CPU-baseline runner of bench.py (oracle/ref_runner.py): the copied reference module — or, without it, the torch port —
must agree with the numpy oracle, and the expert-sampled variants used at large expert counts must equal the full forward."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_parity
from deeprest_b200 import synth
from oracle import qrnn_numpy as oracle
from oracle import ref_runner

M, B, T, F = 6, 5, 9, 12


@pytest.fixture(scope="module")
def runner():
    ref_runner.ensure_ref_copy()              # no-op where /root/reference is absent (then the port stands in)
    return ref_runner.Runner(synth.weights(3, M, F, 1.5), M, F)


def test_runner_forward_matches_oracle(runner):
    x = synth.windows(4, B, T, F, "diurnal")
    ref = oracle.forward(synth.weights(3, M, F, 1.5), x, M, F)
    assert_parity(runner.forward(x, chunk=2), ref, what=f"ref_runner ({runner.kind})")
    assert runner.kind in ("reference", "port")


def test_sampled_heads_equal_the_full_forward(runner):
    if runner.kind != "reference":
        pytest.skip("needs the reference module (oracle/_ref/qrnn.pyc)")
    x = synth.windows(4, B, T, F, "diurnal")
    full = runner.forward(x)
    ids = [0, 3, 5]
    assert np.array_equal(runner.forward_sampled(x, ids), full[:, :, ids])
    secs, scale = runner.time_step_sampled(x[:1], 3)
    assert secs > 0 and scale == M / 3


def test_reference_train_step_matches_oracle_loss(runner):
    if runner.kind != "reference":
        pytest.skip("needs the reference module")
    r2 = ref_runner.Runner(synth.weights(3, M, F, 1.5), M, F)
    r2.model.dropout.p = 0.0                   # deterministic: no dropout draw
    x = synth.windows(4, B, T, F, "diurnal")
    y = synth.labels(5, B, T, M)
    loss = r2.train_step(x, y)
    out = oracle.forward(synth.weights(3, M, F, 1.5), x, M, F)
    assert abs(loss - float(oracle.quantile_loss(out, y))) < 1e-6


def test_both_bench_arms_print_the_same_config():
    """the driver compares the `config` objects of the two arms (same_config)"""
    sys.path.insert(0, ROOT)
    import bench
    a = bench.workload_config(64, 128, 1024, 288, 64, 1, "tcgen05")
    b = bench.workload_config(64, 128, 1024, 288, 64, 1)
    assert a == b and "configs[1]" in a["workload"]
    assert "configs[3]" in bench.workload_config(1024, 2048, 1024, 288, 64, 8)["workload"]


def test_reference_arm_runs_on_cpu_and_never_truncates_time():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--services", "2", "--windows", "3", "--seq-len", "7", "--features", "8"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["sample"]["time_steps"] == 7 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
