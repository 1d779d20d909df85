"""Expert-sharded forward orchestration, backend agnostic.

The exchange steps of the path — one ``all_reduce(sum)`` of the cross-expert sum S and one
``all_gather`` of the per-rank forecasts — run through ``torch.distributed`` on whatever device
the buffers live on (NCCL over NVLink on the GPUs; gloo in the CPU tests).  The three compute
phases are callables so the same orchestration is exercised by ``tests/test_sharding_gloo.py``
without a GPU (there the callables are oracle code; in the product they are the C-ABI phases
``dr_forward_local_dev`` / ``dr_forward_heads_dev`` / ``dr_interleave_dev``).
"""
from __future__ import annotations


def sharded_forward(x, *, world, m_local, q, s_elems, local_fn, heads_fn, interleave_fn, group=None):
    """x [B,T,F] (replicated on every rank) -> forecasts [B,T,world*m_local,q] on every rank."""
    import torch
    import torch.distributed as dist

    B, T = int(x.shape[0]), int(x.shape[1])
    S = torch.empty((int(s_elems),), device=x.device, dtype=torch.float32)
    out_local = torch.empty((B, T, m_local, q), device=x.device, dtype=torch.float32)
    local_fn(x, S, out_local)                                   # local bi-GRUs: partial S, own-expert head term
    dist.all_reduce(S, op=dist.ReduceOp.SUM, group=group)       # head i needs every other expert's output
    heads_fn(S, out_local)                                      # + (A_i/(M-1))·S + b_i
    flat = torch.empty((world * B, T, m_local, q), device=x.device, dtype=torch.float32)
    dist.all_gather_into_tensor(flat, out_local, group=group)   # rank-major concatenation along dim 0
    gathered = flat.view(world, B, T, m_local, q)
    out = torch.empty((B, T, world * m_local, q), device=x.device, dtype=torch.float32)
    interleave_fn(gathered, out)                                # [w][B,T,M/w,Q] -> reference layout [B,T,M,Q]
    return out
