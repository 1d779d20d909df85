"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/deeprest_b200.h declares, the ctypes struct matches the header, and the host-side
layout/sharding logic is right.  No compute calls (there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from deeprest_b200 import _lib, layout, synth


def header_functions():
    text = open(os.path.join(ROOT, "include", "deeprest_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dr_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signature table out of sync with the header"
    assert lib.dr_version() >= 100


def test_config_struct_matches_header():
    # int32 F,M,H,Q; float[8]; float; int32 engine,device,rank,world,dtype — field order as in the header
    assert C.sizeof(_lib.DrConfig) == 4 * 4 + 8 * 4 + 4 + 5 * 4
    text = open(os.path.join(ROOT, "include", "deeprest_b200.h")).read()
    body = re.search(r"typedef struct dr_config \{(.*?)\} dr_config;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:                                           # "int32_t rank, world" / "float quantiles[8]"
            names += [re.sub(r"\[\d+\]", "", n).strip() for n in decl.split(None, 1)[1].split(",")]
    assert names == [f[0] for f in _lib.DrConfig._fields_], names


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from deeprest_b200 import QuantileRNN
    with pytest.raises(_lib.DeepRestError) as ei:
        QuantileRNN(input_size=4, num_metrics=2)
    assert ei.value.code == _lib.DR_ECUDA


def test_argument_validation_happens_before_cuda():
    lib = _lib.load()
    h = C.c_void_p()
    cfg = _lib.DrConfig(F=4, M=2, H=64, Q=3, engine=0, device=0, rank=0, world=1)
    assert lib.dr_create(C.byref(cfg), C.byref(h)) == _lib.DR_EUNSUPPORTED
    cfg = _lib.DrConfig(F=4, M=1, H=128, Q=3, engine=0, device=0, rank=0, world=1)
    assert lib.dr_create(C.byref(cfg), C.byref(h)) == _lib.DR_EINVAL
    assert b"num_metrics" in lib.dr_last_error(None)
    cfg = _lib.DrConfig(F=4, M=6, H=128, Q=3, engine=0, device=0, rank=0, world=4)
    assert lib.dr_create(C.byref(cfg), C.byref(h)) == _lib.DR_EINVAL


def test_blob_layout_matches_reference_formula():
    # P_e = 256 + (128F+F) + 2(384F + 384*128 + 768) + 1539
    for F in (1, 5, 16, 64):
        assert layout.params_per_expert(F) == 256 + (128 * F + F) + 2 * (384 * F + 384 * 128 + 768) + 1539
    assert layout.params_per_expert(16) == 115987
    offs = layout.expert_offsets(16)
    assert list(offs)[0] == "mask_w1" and list(offs)[-1] == "head_b"
    blob = synth.weights(1, 3, 16)
    sd = layout.state_dict_from_blob(blob, 3, 16)
    assert np.array_equal(layout.blob_from_state_dict(sd, 3, 16), blob)


def test_cuda_offsets_mirror_python_layout():
    """csrc/dr_common.cuh::dr_blob_offsets is a hand mirror of layout.py; keep the order pinned."""
    text = open(os.path.join(ROOT, "deeprest_b200", "csrc", "dr_common.cuh")).read()
    body = text[text.index("dr_blob_offsets(int F)"):text.index("struct dr_model")]
    order = re.findall(r"o\.([a-z_0-9]+)(?:\[d\])?\s*=\s*off", body)
    assert order == ["mask_w1", "mask_b1", "mask_w2", "mask_b2", "w_ih", "w_hh", "b_ih", "b_hh",
                     "head_w", "head_b", "per_expert"]


def test_generator_is_counter_based():
    a = synth.uniform(2021, 1000)
    b = synth.uniform(2021, 10, offset=500)
    assert np.array_equal(a[500:510], b) and a.min() >= 0 and a.max() < 1
    assert abs(a.mean() - 0.5) < 0.05
    w = synth.weights(3, 2, 16)
    ex = layout.unpack_blob(w, 2, 16)[0]
    assert np.abs(ex["w_hh_f"]).max() <= 1 / np.sqrt(128) and np.abs(ex["head_w"]).max() <= 1 / np.sqrt(512)


def test_expert_sharding_plan():
    """layout.expert_range is the library's rule (dr_create: equal contiguous shards, M % world == 0)."""
    for M, world in [(2048, 8), (128, 1), (1024, 8), (6, 3), (6, 2)]:
        spans = [layout.expert_range(r, world, M) for r in range(world)]
        assert spans == [(r * (M // world), (r + 1) * (M // world)) for r in range(world)]
    with pytest.raises(ValueError):
        layout.expert_range(0, 4, 6)


def test_series_window_count_matches_reference_sliding_window():
    """dr_series_windows mirrors utils.py:4-5 (range(len - W): the last full window is dropped) + a stride."""
    from oracle import qrnn_numpy as oracle
    lib = _lib.load()
    for N, W, stride in [(150, 20, 1), (150, 20, 7), (150, 20, 20), (20, 20, 1), (21, 20, 1), (3, 60, 1), (100, 60, 60)]:
        ref = len(oracle.sliding_window(np.zeros((N, 2)), W)[::stride]) if N - W > 0 else 0
        assert lib.dr_series_windows(N, W, stride) == ref, (N, W, stride)


def test_every_cuda_source_is_part_of_the_build():
    """a .cu file that build.py does not list would silently stay out of libdeeprest_b200.so"""
    import glob
    from deeprest_b200 import build
    here = os.path.dirname(os.path.abspath(build.__file__))
    on_disk = sorted(os.path.basename(p) for p in glob.glob(os.path.join(here, "csrc", "*.cu")))
    assert sorted(build.SOURCES) == on_disk
