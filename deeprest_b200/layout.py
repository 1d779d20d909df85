"""Weight-blob layout for the DeepRest QuantileRNN estimator.

The flat fp32 blob is the reference module's ``state_dict()`` iteration order
with the leading ``mask_init`` scalar dropped (reference:
resource-estimation/qrnn.py:20-26).  Per expert ``e`` the tensors are, in order:

====================  ==========  =====================================
name                  shape       reference parameter
====================  ==========  =====================================
``mask_w1``           [H, 1]      experts.e.0.weight      (qrnn.py:22)
``mask_b1``           [H]         experts.e.0.bias
``mask_w2``           [F, H]      experts.e.1.weight      (qrnn.py:23)
``mask_b2``           [F]         experts.e.1.bias
``w_ih_f``            [3H, F]     experts.e.2.weight_ih_l0 (qrnn.py:24)
``w_hh_f``            [3H, H]     experts.e.2.weight_hh_l0
``b_ih_f``            [3H]        experts.e.2.bias_ih_l0
``b_hh_f``            [3H]        experts.e.2.bias_hh_l0
``w_ih_r``            [3H, F]     experts.e.2.weight_ih_l0_reverse
``w_hh_r``            [3H, H]     experts.e.2.weight_hh_l0_reverse
``b_ih_r``            [3H]        experts.e.2.bias_ih_l0_reverse
``b_hh_r``            [3H]        experts.e.2.bias_hh_l0_reverse
``head_w``            [Q, 4H]     experts.e.3.weight      (qrnn.py:25)
``head_b``            [Q]         experts.e.3.bias
====================  ==========  =====================================

Gate order inside the 3H axis is (r, z, n) as in ``torch.nn.GRU``.
This file is pure host logic (numpy only); the same offsets are compiled into
``csrc/dr_layout.h`` and a test keeps the two in sync.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

H = 128  # hidden size, fixed by the reference default (qrnn.py:7)
Q = 3    # quantiles (.05, .50, .95), qrnn.py:8
QUANTILES = (0.05, 0.50, 0.95)

_STATE_DICT_SUFFIX = OrderedDict([
    ("mask_w1", "0.weight"), ("mask_b1", "0.bias"),
    ("mask_w2", "1.weight"), ("mask_b2", "1.bias"),
    ("w_ih_f", "2.weight_ih_l0"), ("w_hh_f", "2.weight_hh_l0"),
    ("b_ih_f", "2.bias_ih_l0"), ("b_hh_f", "2.bias_hh_l0"),
    ("w_ih_r", "2.weight_ih_l0_reverse"), ("w_hh_r", "2.weight_hh_l0_reverse"),
    ("b_ih_r", "2.bias_ih_l0_reverse"), ("b_hh_r", "2.bias_hh_l0_reverse"),
    ("head_w", "3.weight"), ("head_b", "3.bias"),
])


def expert_shapes(F: int, h: int = H, q: int = Q) -> "OrderedDict[str, tuple]":
    """Ordered name → shape for one expert."""
    return OrderedDict([
        ("mask_w1", (h, 1)), ("mask_b1", (h,)),
        ("mask_w2", (F, h)), ("mask_b2", (F,)),
        ("w_ih_f", (3 * h, F)), ("w_hh_f", (3 * h, h)),
        ("b_ih_f", (3 * h,)), ("b_hh_f", (3 * h,)),
        ("w_ih_r", (3 * h, F)), ("w_hh_r", (3 * h, h)),
        ("b_ih_r", (3 * h,)), ("b_hh_r", (3 * h,)),
        ("head_w", (q, 4 * h)), ("head_b", (q,)),
    ])


def expert_offsets(F: int, h: int = H, q: int = Q) -> "OrderedDict[str, tuple]":
    """Ordered name → (offset_in_floats, shape) inside one expert's slice."""
    out = OrderedDict()
    off = 0
    for name, shape in expert_shapes(F, h, q).items():
        out[name] = (off, shape)
        off += int(np.prod(shape))
    return out


def params_per_expert(F: int, h: int = H, q: int = Q) -> int:
    """P_e = 2h + (hF + F) + 2(3hF + 3h·h + 6h) + (4h·q + q); 115,987 at F=16."""
    return sum(int(np.prod(s)) for s in expert_shapes(F, h, q).values())


def blob_size(M: int, F: int) -> int:
    return M * params_per_expert(F)


def unpack_blob(blob: np.ndarray, M: int, F: int):
    """Split a flat blob into a list (one per expert) of name → ndarray views."""
    blob = np.asarray(blob)
    pe = params_per_expert(F)
    if blob.ndim != 1 or blob.size != M * pe:
        raise ValueError(f"blob has {blob.size} floats, expected {M}*{pe}")
    offs = expert_offsets(F)
    experts = []
    for e in range(M):
        base = e * pe
        experts.append({n: blob[base + o: base + o + int(np.prod(s))].reshape(s)
                        for n, (o, s) in offs.items()})
    return experts


def state_dict_keys(M: int):
    """The reference ``state_dict`` keys in blob order (without ``mask_init``)."""
    return [f"experts.{e}.{sfx}" for e in range(M) for sfx in _STATE_DICT_SUFFIX.values()]


def blob_from_state_dict(sd, M: int, F: int) -> np.ndarray:
    """Flatten a torch/numpy ``state_dict`` (reference key names) into the blob."""
    parts = []
    for key in state_dict_keys(M):
        t = sd[key]
        parts.append(np.asarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t,
                                dtype=np.float32).reshape(-1))
    blob = np.concatenate(parts)
    assert blob.size == blob_size(M, F)
    return blob


def state_dict_from_blob(blob: np.ndarray, M: int, F: int):
    """Inverse of :func:`blob_from_state_dict` (numpy arrays; adds ``mask_init``)."""
    sd = OrderedDict()
    sd["mask_init"] = np.ones((1,), np.float32)
    names = list(_STATE_DICT_SUFFIX.items())
    for e, ex in enumerate(unpack_blob(np.asarray(blob, np.float32), M, F)):
        for name, sfx in names:
            sd[f"experts.{e}.{sfx}"] = ex[name].copy()
    return sd


def expert_range(rank: int, world: int, M: int):
    """Contiguous expert shard owned by ``rank`` — the same rule the library applies (csrc/dr_api.cu::dr_create,
    readable back through ``dr_local_experts``): equal shards of M/world experts, M divisible by world
    (shard by service ID; with M = 2*services and an even M/world both metrics of a service share a rank)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    if M % world:
        raise ValueError("num_metrics must be divisible by world (equal expert shards, as dr_create requires)")
    per = M // world
    return rank * per, (rank + 1) * per
