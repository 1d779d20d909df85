"""Expert-sharded forward orchestration, backend agnostic.

The exchange steps of the path — one ``all_reduce(sum)`` of the cross-expert sum S and one
gather of the per-rank forecasts — run through ``torch.distributed`` on whatever device
the buffers live on (NCCL over NVLink on the GPUs; gloo in the CPU tests); on CUDA the gather goes
over peer-mapped memory instead when it is available (K2 peer stores or DMA-engine 2-D copies).  The three compute
phases are callables so the same orchestration is exercised by ``tests/test_sharding_gloo.py``
without a GPU (there the callables are oracle code; in the product they are the C-ABI phases
``dr_forward_local_dev`` / ``dr_forward_heads_dev`` / ``dr_interleave_dev``).

On CUDA the batch is cut into chunks of whole 256-window pair tiles and each chunk's pipeline
(recurrence, all-reduce, heads, all-gather, interleave) runs on one of two alternating side
streams: the exchange of chunk c overlaps the recurrence kernel of chunk c+1, and the CTAs of
chunk c+1's recurrence fill the SMs that chunk c's last wave leaves idle.  Windows are
independent, so this is exact.  (The library keeps 4 operand-image workspace slots for this.)
"""
from __future__ import annotations

_side_streams = {}


def _chunks(B, is_cuda):
    if not is_cuda or B < 512:
        return [(0, B)]
    n = max(2, min(4, B // 256))
    step = ((B + n - 1) // n + 255) // 256 * 256
    return [(b0, min(B, b0 + step)) for b0 in range(0, B, step)]


class PeerBuffers:
    """Full forecast tensors [B,T,M,Q] in peer-mapped symmetric memory (torch.distributed._symmetric_memory):
    every rank can store into every rank's buffer, which lets the head kernel write the stacked forecast tensor on all
    GPUs directly over NVLink (dr_forward_heads_p2p_dev) instead of all-gather + interleave.  Two buffers per shape
    alternate, so a result stays valid while the next forward runs."""

    def __init__(self, group):
        self.group, self.cache, self.failed = group, {}, None

    def acquire(self, shape, device):
        import ctypes as C
        import torch
        import torch.distributed._symmetric_memory as symm
        key = (tuple(shape), str(device))
        ent = self.cache.get(key)
        if ent is None:
            bufs = []
            for _ in range(2):
                t = symm.empty(tuple(shape), dtype=torch.float32, device=device)
                hdl = symm.rendezvous(t, self.group)
                ptrs = (C.c_void_p * hdl.world_size)(*[int(p) for p in hdl.buffer_ptrs])
                bufs.append((t, hdl, ptrs))
            ent = self.cache[key] = {"bufs": bufs, "next": 0}
        t, hdl, ptrs = ent["bufs"][ent["next"]]
        ent["next"] ^= 1
        return t, hdl, ptrs


def sharded_forward(x, *, world, m_local, q, s_elems, local_fn, heads_fn, interleave_fn, group=None,
                    peer=None, heads_p2p_fn=None, scatter_fn=None):
    """x [B,T,F] (replicated on every rank) -> forecasts [B,T,world*m_local,q] on every rank.

    local_fn(x_c, S_c, out_local_c), heads_fn(S_c, out_local_c), interleave_fn(gathered_c, out_c) act on
    a chunk of windows; s_elems(bn) gives the size of S for bn windows."""
    import torch
    import torch.distributed as dist

    B, T = int(x.shape[0]), int(x.shape[1])
    p2p = None
    if peer is not None and (heads_p2p_fn is not None or scatter_fn is not None) and x.is_cuda and peer.failed is None:
        try:
            p2p = peer.acquire((B, T, world * m_local, q), x.device)
        except Exception as exc:                     # no symmetric memory on this system: NCCL all-gather path
            peer.failed = repr(exc)
    if p2p is not None:
        out, hdl, ptrs = p2p
        hdl.barrier(channel=0)                       # every rank is done with the previous contents of this buffer
    else:
        out = torch.empty((B, T, world * m_local, q), device=x.device, dtype=torch.float32)
    chunks = _chunks(B, x.is_cuda)
    overlap = x.is_cuda and len(chunks) > 1
    if overlap:
        main = torch.cuda.current_stream(x.device)
        sides = _side_streams.get(x.device)
        if sides is None:
            sides = _side_streams[x.device] = [torch.cuda.Stream(device=x.device) for _ in range(2)]
        start = main.record_event()
    for c, (b0, b1) in enumerate(chunks):
        bn = b1 - b0
        S = torch.empty((int(s_elems(bn)),), device=x.device, dtype=torch.float32)
        out_local = torch.empty((bn, T, m_local, q), device=x.device, dtype=torch.float32)
        flat = None if p2p is not None else torch.empty((world * bn, T, m_local, q), device=x.device, dtype=torch.float32)

        def pipeline():
            local_fn(x[b0:b1], S, out_local)                       # local bi-GRUs: partial S, own-expert head term
            dist.all_reduce(S, op=dist.ReduceOp.SUM, group=group)  # head i needs every other expert's output
            if p2p is not None and heads_p2p_fn is not None:       # heads + all-gather + interleave in ONE kernel:
                heads_p2p_fn(S, bn, ptrs, b0)                      # stores go to every rank's out[b0:b1] over NVLink
                return
            heads_fn(S, out_local)                                 # + (A_i/(M-1))·S + b_i
            if p2p is not None:                                    # strided 2-D peer copies by the DMA engines place the
                scatter_fn(out_local, bn, ptrs, b0)                # columns into every rank's out[b0:b1]: no SM time
                return
            dist.all_gather_into_tensor(flat, out_local, group=group)   # rank-major concatenation along dim 0
            interleave_fn(flat.view(world, bn, T, m_local, q), out[b0:b1])   # -> reference layout [B,T,M,Q]

        if overlap:
            side = sides[c % 2]
            with torch.cuda.stream(side):
                side.wait_event(start)
                pipeline()
            for t in (S, out_local, flat, out, x):
                if t is not None:
                    t.record_stream(side)
        else:
            pipeline()
    if overlap:
        for side in sides:
            main.wait_stream(side)
    if p2p is not None:
        hdl.barrier(channel=1)                       # every rank's stores into this rank's buffer have landed
    return out
