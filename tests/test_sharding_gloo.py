"""world_size-2 gloo test (CPU) of the expert-sharded forward: the orchestration in
deeprest_b200/sharding.py (all_reduce of S, all_gather of forecasts, interleave) with oracle code
standing in for the three CUDA phases must reproduce the unsharded oracle forward."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deeprest_b200 import layout, synth
from deeprest_b200.sharding import sharded_forward
from oracle import qrnn_numpy as oracle

M, B, T, F, Q, H = 6, 5, 9, 7, layout.Q, layout.H


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        blob = synth.weights(4, M, F, 1.5)
        x = synth.windows(6, B, T, F, "diurnal")
        m_local = M // world
        lo, hi = rank * m_local, (rank + 1) * m_local
        experts = layout.unpack_blob(blob, M, F)[lo:hi]
        def local_fn(xx, S, out_local):
            r_local = [oracle.expert_rnn_out(ex, xx.numpy()) for ex in experts]          # [bn,T,2H] each
            S.copy_(torch.from_numpy(sum(r_local).reshape(-1)))
            own = np.stack([r @ (ex["head_w"][:, 2 * H:] - ex["head_w"][:, :2 * H] / (M - 1)).T
                            for r, ex in zip(r_local, experts)], axis=2)     # (C - A/(M-1))·r_i
            out_local.copy_(torch.from_numpy(own.astype(np.float32)))

        def heads_fn(S, out_local):
            Sn = S.numpy().reshape(out_local.shape[0], T, 2 * H)
            add = np.stack([Sn @ (ex["head_w"][:, :2 * H] / (M - 1)).T + ex["head_b"] for ex in experts], axis=2)
            out_local.add_(torch.from_numpy(add.astype(np.float32)))

        def interleave_fn(gathered, out):
            out.copy_(gathered.permute(1, 2, 0, 3, 4).reshape(out.shape[0], T, world * m_local, Q))

        out = sharded_forward(torch.from_numpy(x), world=world, m_local=m_local, q=Q, s_elems=lambda bn: bn * T * 2 * H,
                              local_fn=local_fn, heads_fn=heads_fn, interleave_fn=interleave_fn)
        ref = oracle.forward(blob, x, M, F)
        err = float(np.abs(out.numpy() - ref).max())
        with open(f"{result_path}.{rank}", "w") as f:
            f.write(repr(err))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_forward_matches_unsharded_oracle(tmp_path):
    world = 2
    port = _free_port()
    path = str(tmp_path / "err")
    mp.spawn(_worker, args=(world, port, path), nprocs=world, join=True)
    for r in range(world):
        err = float(open(f"{path}.{r}").read())
        assert err < 2e-6, f"rank {r}: sharded vs unsharded max err {err}"


def test_expert_shards_are_equal_and_cover():
    for world in (1, 2, 4, 8):
        M_ = 2048
        spans = [(r * (M_ // world), (r + 1) * (M_ // world)) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == M_


def test_chunk_plan_covers_the_batch_in_pair_tiles():
    from deeprest_b200.sharding import _chunks
    assert _chunks(300, True) == [(0, 300)] and _chunks(4096, False) == [(0, 4096)]
    for B in (512, 1024, 1000, 4096, 777):
        ch = _chunks(B, True)
        assert ch[0][0] == 0 and ch[-1][1] == B and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))
        assert all((b1 - b0) % 256 == 0 for b0, b1 in ch[:-1]) and 2 <= len(ch) <= 4
