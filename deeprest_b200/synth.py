"""Counter-based synthetic trace generator, reproducible in numpy, C++ and CUDA.

``u(seed, i) = (mix64(seed * GOLDEN + i) >> 40) / 2**24`` — a 24-bit uniform in
[0, 1), exactly representable in fp32, a pure function of (seed, linear index).
SURVEY §8(d) asks for exactly this property so the GPU box can regenerate the
golden inputs without shipping them.  ``mix64`` is the splitmix64 finaliser.

The "diurnal" variant shapes the per-feature rate with the two-peak day curve
of the reference's load generator (locust/locustfile-normal.py:53-74 describes
the shape: two Gaussian-ish peaks over a base load); only the shape idea is
reused, the arithmetic here is our own.
"""
from __future__ import annotations

import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def mix64(z: np.ndarray) -> np.ndarray:
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return z


def uniform(seed: int, n: int, offset: int = 0) -> np.ndarray:
    """n fp32 uniforms in [0,1) for linear indices offset..offset+n-1."""
    idx = np.arange(offset, offset + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        state = np.uint64(seed) * GOLDEN + idx
    return ((mix64(state) >> np.uint64(40)).astype(np.float32)
            * np.float32(1.0 / (1 << 24)))


def windows(seed: int, B: int, T: int, F: int, kind: str = "uniform") -> np.ndarray:
    """Synthetic normalised API-rate windows x[B,T,F] (fp32, >= 0)."""
    u = uniform(seed, B * T * F).reshape(B, T, F)
    if kind == "uniform":
        return u
    if kind == "diurnal":
        t = (np.arange(T, dtype=np.float32) + 0.5) / np.float32(T)
        shape = (0.15 + 0.55 * np.exp(-((t - 0.33) / 0.09) ** 2)
                 + 0.85 * np.exp(-((t - 0.72) / 0.11) ** 2)).astype(np.float32)
        base = uniform(seed + 7919, B * F).reshape(B, 1, F) * np.float32(2.5)
        x = base * shape[None, :, None] + np.float32(0.2) * (u - np.float32(0.5))
        return np.clip(x, 0.0, 3.0).astype(np.float32)
    raise ValueError(kind)


def labels(seed: int, B: int, T: int, M: int) -> np.ndarray:
    """Synthetic normalised utilisation labels y[B,T,M] in [0,1)."""
    return uniform(seed, B * T * M).reshape(B, T, M)


def weights(seed: int, M: int, F: int, scale: float = 1.0, experts=None) -> np.ndarray:
    """A full weight blob with the reference's default-init *distribution*.

    torch defaults: ``nn.Linear`` U(±1/sqrt(fan_in)), ``nn.GRU``
    U(±1/sqrt(H)).  ``scale`` > 1 widens every range to mimic trained weights
    with saturating gates.  Values come from :func:`uniform`, so the blob is a
    pure function of (seed, M, F, scale).  ``experts=(lo, hi)`` fills only that shard of the
    full-size blob (the rest stays zero) — what a sharded rank needs, without generating all M.
    """
    from .layout import H, expert_offsets, params_per_expert

    pe = params_per_expert(F)
    lo, hi = (0, M) if experts is None else experts
    blob = np.zeros((M, pe), np.float32)
    if hi <= lo:
        return blob.reshape(-1)
    u = uniform(seed, (hi - lo) * pe, offset=lo * pe).reshape(hi - lo, pe)
    bound = {
        "mask_w1": 1.0, "mask_b1": 1.0,                    # fan_in = 1
        "mask_w2": 1.0 / np.sqrt(H), "mask_b2": 1.0 / np.sqrt(H),
        "head_w": 1.0 / np.sqrt(4 * H), "head_b": 1.0 / np.sqrt(4 * H),
    }
    for name, (off, shape) in expert_offsets(F).items():
        n = int(np.prod(shape))
        b = np.float32(bound.get(name, 1.0 / np.sqrt(H)) * scale)
        blob[lo:hi, off:off + n] = (u[:, off:off + n] * np.float32(2.0) - np.float32(1.0)) * b
    return blob.reshape(-1)
