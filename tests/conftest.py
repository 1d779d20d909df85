import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# fp32 parity bar of north_star ("1e-4 rel"), made well-posed per SURVEY §0:
# |a-b| <= ATOL + RTOL*|b|  (outputs cross zero, so a pure relative bound is ill-posed)
RTOL, ATOL = 1e-4, 1e-6


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_cases(prefix="g"):
    paths = sorted(glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))
    return [p for p in paths if "train_step" not in p and "window_norm" not in p]


def load_golden(path):
    from deeprest_b200 import synth
    g = dict(np.load(path))
    M, B, T, F = (int(g[k]) for k in ("M", "B", "T", "F"))
    if "blob" in g:
        blob = g["blob"].astype(np.float32)
    else:
        blob = synth.weights(int(g["wseed"]), M, F, float(g["wscale"]))
    x = synth.windows(int(g["xseed"]), B, T, F, str(g["xkind"]))
    # generator drift check: the fixture was minted from exactly these tensors
    assert abs(blob.astype(np.float64).sum() - float(g["blob_sum"])) < 1e-9 * max(1.0, abs(float(g["blob_sum"]))) + 1e-9
    if "x_sum" in g:
        assert abs(x.astype(np.float64).sum() - float(g["x_sum"])) < 1e-6
    g.update(M=M, B=B, T=T, F=F, blob_arr=blob, x=x)
    return g


def assert_parity(got, ref, rtol=RTOL, atol=ATOL, what=""):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = np.abs(got - ref)
    bound = atol + rtol * np.abs(ref)
    worst = np.argmax(err - bound)
    assert np.all(err <= bound), (
        f"{what}: max abs err {err.max():.3e}, worst idx {np.unravel_index(worst, ref.shape)} "
        f"got {got.flat[worst]:.8g} ref {ref.flat[worst]:.8g}; MAE {err.mean():.3e}")


@pytest.fixture(scope="session")
def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
