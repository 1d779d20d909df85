"""Pins the tcgen05 building blocks the tensor-core GRU engine relies on (descriptor bits,
SW128 K-major image, bulk copy, A-in-TMEM packing, cta_group::2 operand split, multicast
commit, tcgen05.ld lane mapping) with one GEMM tile per variant, checked against numpy."""
import ctypes as C

import numpy as np
import pytest

from deeprest_b200 import _lib

pytestmark = pytest.mark.gpu


def bf16_round(a):
    """fp32 -> bf16 (round to nearest even) as uint16 bit patterns and the rounded fp32 values."""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    b = ((u + r) >> 16).astype(np.uint16)
    return b, (b.astype(np.uint32) << 16).view(np.float32)


def sw128_image(bits):
    """[rows, K] uint16 -> byte image [K/64][rows x 128 B], K-major 128B-swizzled (dr_tc.cuh::sw128_offset)."""
    rows, K = bits.shape
    assert rows % 8 == 0 and K % 64 == 0
    img = np.zeros((K // 64, rows * 64), np.uint16)
    r = np.arange(rows)[:, None]
    k = np.arange(64)[None, :]
    off = (r // 8) * 1024 + (r % 8) * 128 + (((k // 8) ^ (r % 8)) * 16) + (k % 8) * 2
    for kb in range(K // 64):
        img[kb, off // 2] = bits[:, kb * 64:(kb + 1) * 64]
    return img


def f16_round(a):
    h = np.ascontiguousarray(a, np.float32).astype(np.float16)
    return h.view(np.uint16), h.astype(np.float32)


def run_probe(variant, A, B, flags=0):
    lib = _lib.load()
    cg = 2 if variant & 2 else 1
    N, K = B.shape
    rnd = f16_round if flags & 2 else bf16_round
    a_bits, a_val = rnd(A)
    b_bits, b_val = rnd(B)
    if variant & 1:
        a_buf = np.ascontiguousarray(a_bits)
    else:
        a_buf = np.ascontiguousarray(np.stack([sw128_image(a_bits[c * 128:(c + 1) * 128]) for c in range(cg)]))
    nloc = N // cg
    b_buf = np.ascontiguousarray(np.stack([sw128_image(b_bits[c * nloc:(c + 1) * nloc]) for c in range(cg)]))
    out = np.zeros((cg * 128, N), np.float32)
    rc = lib.dr_tc_probe(variant, a_buf.ctypes.data, a_buf.nbytes, b_buf.ctypes.data, b_buf.nbytes,
                         N, K, flags, out.ctypes.data_as(C.POINTER(C.c_float)))
    assert rc == 0, lib.dr_last_error(None)
    ref = a_val.astype(np.float64) @ b_val.astype(np.float64).T
    return out, ref


def test_tcgen05_tile_fp16_operands():
    """the GRU engine's operand format: fp16 x fp16 -> fp32 (idesc format code 0), 2-CTA, A from TMEM"""
    rng = np.random.default_rng(5)
    A = rng.standard_normal((256, 128)).astype(np.float32)
    B = rng.standard_normal((96, 128)).astype(np.float32)
    for variant in (2, 3):
        out, ref = run_probe(variant, A, B, flags=2)
        assert np.abs(out - ref).max() < 1e-3, (variant, np.abs(out - ref).max())


@pytest.mark.parametrize("variant,name", [(0, "SS cta_group::1"), (1, "TS cta_group::1"),
                                           (2, "SS cta_group::2"), (3, "TS cta_group::2")])
@pytest.mark.parametrize("N,K", [(96, 64), (96, 128), (256, 128), (32, 192)])
def test_tcgen05_tile(variant, name, N, K):
    rng = np.random.default_rng(variant * 100 + N + K)
    cg = 2 if variant & 2 else 1
    A = rng.standard_normal((cg * 128, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    out, ref = run_probe(variant, A, B)
    err = np.abs(out - ref).max()
    print(f"{name} N={N} K={K}: max err {err:.3e}")
    if err > 1e-3:
        # diagnostics: which hypothesis would have matched?
        hints = []
        if variant & 1:
            o2, _ = run_probe(variant, A, B, flags=1)
            hints.append(f"swapped bf16 halves -> {np.abs(o2 - ref).max():.3e}")
        if cg == 2:
            for perm_name, cols in [("B halves swapped", np.r_[N // 2:N, 0:N // 2])]:
                hints.append(f"{perm_name} -> {np.abs(out - ref[:, cols]).max():.3e}")
            hints.append(f"rows swapped -> {np.abs(out - ref[np.r_[128:256, 0:128]]).max():.3e}")
        np.save(f"gpurun_out/probe_v{variant}_N{N}_K{K}_out.npy", out)
        np.save(f"gpurun_out/probe_v{variant}_N{N}_K{K}_ref.npy", ref)
        pytest.fail(f"{name} N={N} K={K}: max err {err:.3e}; " + "; ".join(hints))


def mn_image(bits, nblk_pad=None):
    """[K, MN] uint16 (MN contiguous per reduced index k) -> byte image [MN/64 blocks][K/8 groups][8 rows x 128 B]:
    the MN-major SW128 canonical layout (16-byte chunks of a 128-byte row XOR-swizzled by k & 7)."""
    K, MN = bits.shape
    nb = (MN + 63) // 64
    pad = np.zeros((K, nb * 64), np.uint16)
    pad[:, :MN] = bits
    img = np.zeros((nb, K // 8, 8, 64), np.uint16)
    k = np.arange(K)
    for blk in range(nb):
        for ch in range(8):
            src = pad[:, blk * 64 + ch * 8: blk * 64 + ch * 8 + 8]              # [K, 8]
            dst_ch = ch ^ (k & 7)
            for kk in range(K):
                img[blk, kk // 8, kk % 8, dst_ch[kk] * 8: dst_ch[kk] * 8 + 8] = src[kk]
    return img


def run_probe_mn(A, B, params=None):
    """A [K,128], B [K,N] fp32 -> D[128,N] = A^T B on the tensor core with both operands MN-major."""
    lib = _lib.load()
    K, N = B.shape
    a_bits, a_val = bf16_round(A)
    b_bits, b_val = bf16_round(B)
    a_img = np.ascontiguousarray(mn_image(a_bits))
    b_img = np.ascontiguousarray(mn_image(b_bits))
    blk = (K // 8) * 1024                        # bytes between 64-element MN blocks
    if params is None:                           # hypothesis H1: LBO = MN-block stride, SBO = 8-k group stride, 2 groups per K=16
        params = [blk, 1024, 2048, blk, 1024, 2048, 1, 1]
    prm = (C.c_uint32 * 8)(*params)
    out = np.zeros((128, N), np.float32)
    rc = lib.dr_tc_probe_mn(a_img.ctypes.data, a_img.nbytes, b_img.ctypes.data, b_img.nbytes, N, K, prm,
                            out.ctypes.data_as(C.POINTER(C.c_float)))
    assert rc == 0, lib.dr_last_error(None)
    ref = a_val.astype(np.float64).T @ b_val.astype(np.float64)
    return out, ref


@pytest.mark.parametrize("N,K", [(128, 64), (208, 64), (80, 128), (16, 32)])
def test_tcgen05_tile_mn_major(N, K):
    rng = np.random.default_rng(N + K)
    A = rng.standard_normal((K, 128)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    out, ref = run_probe_mn(A, B)
    err = np.abs(out - ref).max()
    print(f"MN-major N={N} K={K}: max err {err:.3e}")
    if err > 1e-3:
        blk = (K // 8) * 1024
        hints = []
        for name, prm in [("LBO/SBO swapped", [1024, blk, 2048, 1024, blk, 2048, 1, 1]),
                          ("kstep = 1 group", [blk, 1024, 1024, blk, 1024, 1024, 1, 1]),
                          ("major bits off", [blk, 1024, 2048, blk, 1024, 2048, 0, 0])]:
            try:
                o2, _ = run_probe_mn(A, B, prm)
                hints.append(f"{name} -> {np.abs(o2 - ref).max():.3e}")
            except AssertionError as exc:
                hints.append(f"{name} -> {exc}")
        np.save(f"gpurun_out/probe_mn_N{N}_K{K}_out.npy", out)
        np.save(f"gpurun_out/probe_mn_N{N}_K{K}_ref.npy", ref)
        pytest.fail(f"MN-major N={N} K={K}: max err {err:.3e}; " + "; ".join(hints))
