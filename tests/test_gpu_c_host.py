"""The C ABI from a non-Python host: tests/c_host/abi_driver.c (plain C, gcc) links libdeeprest_b200.so, runs the
forward on files, and its forecasts must match the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_parity
from deeprest_b200 import synth
from oracle import qrnn_numpy as oracle

pytestmark = pytest.mark.gpu


def test_c_host_forward_matches_oracle(tmp_path):
    libdir = os.path.join(ROOT, "deeprest_b200")
    exe = str(tmp_path / "abi_driver")
    subprocess.run(["gcc", "-O2", "-std=c99", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_host", "abi_driver.c"), "-o", exe,
                    "-L", libdir, "-ldeeprest_b200", f"-Wl,-rpath,{libdir}"], check=True)
    M, B, T, F = 4, 37, 21, 12
    blob = synth.weights(5, M, F, 1.5)
    x = synth.windows(6, B, T, F, "diurnal")
    blob.tofile(tmp_path / "blob.bin")
    x.tofile(tmp_path / "x.bin")
    r = subprocess.run([exe, str(tmp_path / "blob.bin"), str(tmp_path / "x.bin"), str(tmp_path / "out.bin"),
                        str(F), str(M), str(B), str(T)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(tmp_path / "out.bin", np.float32).reshape(B, T, M, 3)
    ref = oracle.forward(blob, x, M, F)
    assert_parity(out, ref, what="C host")
    loss = float(r.stdout.split("loss=")[1])
    assert abs(loss - float(oracle.quantile_loss(ref, np.zeros((B, T, M), np.float32)))) < 1e-5
    assert "engine=tcgen05" in r.stdout


def test_c_host_sharded_forward_two_gpus(tmp_path):
    """SURVEY §8e from a plain-C host: two handles in one process, dr_comm_init / dr_comm_attach / dr_forward_sharded_dev."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    libdir = os.path.join(ROOT, "deeprest_b200")
    exe = str(tmp_path / "abi_driver_sharded")
    subprocess.run(["gcc", "-O2", "-std=c99", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_host", "abi_driver_sharded.c"), "-o", exe,
                    "-L", libdir, "-ldeeprest_b200", "-L/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{libdir}",
                    "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    M, B, T, F = 8, 600, 24, 32                       # 3 chunks of 256 windows, the last one ragged
    blob = synth.weights(5, M, F, 1.5)
    x = synth.windows(6, B, T, F, "diurnal")
    blob.tofile(tmp_path / "blob.bin")
    x.tofile(tmp_path / "x.bin")
    r = subprocess.run([exe, str(tmp_path / "blob.bin"), str(tmp_path / "x.bin"), str(tmp_path / "out.bin"),
                        str(F), str(M), str(B), str(T), "2"], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stderr + r.stdout
    out = np.fromfile(tmp_path / "out.bin", np.float32).reshape(B, T, M, 3)
    ref = oracle.forward(blob, x[:8], M, F)
    assert_parity(out[:8], ref, what="C host, 2 GPUs")
    from deeprest_b200 import QuantileRNN
    single = QuantileRNN(F, M).eval()
    single.load_blob(blob)
    ref1 = single(x)
    single.close()
    assert np.abs(out - ref1).max() < 2e-6
    assert "engine=tcgen05 world=2" in r.stdout
