"""Host-side mirror of the reference estimator interface, backed by libdeeprest_b200.so.

``QuantileRNN`` keeps the reference's constructor, method names, argument meaning and
error behaviour (resource-estimation/qrnn.py:6-75) so that ``estimate.py``-style code and
the reference's own call sites (estimate.py:60,70,71,91,92) work unchanged:

    model = QuantileRNN(input_size=F, num_metrics=M)      # qrnn.py:7
    model.load_state_dict(sd); model.eval()
    out = model(x)                                        # [B,T,F] -> [B,T,M,Q]   qrnn.py:28
    loss = model.quantile_loss(out, y)                    # qrnn.py:58

All arithmetic runs in the CUDA library through its C ABI (ctypes).  numpy arrays use the
host entry points (copies inside the call); CUDA torch tensors use the ``*_dev`` entry
points on torch's current stream.  There is no CPU path.

Expert-sharded multi-GPU execution: pass ``process_group`` (any
``torch.distributed`` group — NCCL on GPUs); the forward then runs
local bi-GRUs -> all_reduce(S) -> heads -> all_gather(forecasts) -> interleave.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, layout


def sliding_window(ts, window_size):
    """utils.py:4-5 — stride-1 windows; like the reference it drops the final window."""
    ts = np.asarray(ts)
    n = len(ts) - window_size
    if n <= 0:
        return np.asarray([])
    idx = np.arange(window_size)[None, :] + np.arange(n)[:, None]
    return ts[idx]


def _is_torch(t):
    return type(t).__module__.startswith("torch")


class QuantileRNN:
    def __init__(self, input_size, num_metrics, hidden_layer_size=128, num_layers=1, bidirectional=True,
                 quantiles=(.05, .50, .95), dropout=0.50, *, engine="auto", device=None,
                 process_group=None, rank=None, world=None, dtype="fp32"):
        if hidden_layer_size != layout.H or num_layers != 1 or not bidirectional:
            raise NotImplementedError(
                "libdeeprest_b200 implements the reference defaults only: hidden_layer_size=128, "
                "num_layers=1, bidirectional=True (qrnn.py:7)")
        if len(quantiles) != layout.Q:
            raise NotImplementedError("exactly 3 quantiles are supported (reference default, qrnn.py:8)")
        self.input_size, self.num_metrics = int(input_size), int(num_metrics)
        self.hidden_layer_size, self.num_layers, self.bidirectional = hidden_layer_size, num_layers, bidirectional
        self.quantiles, self.dropout_p = tuple(float(q) for q in quantiles), float(dropout)
        self.training = True                     # nn.Module default
        self._pg = process_group
        self._engine, self._peer = engine, None
        # sharded runs (world > 1), how the partial sums S and the forecast columns cross the ranks:
        #   "dma"   : the library's own exchange behind ONE C-ABI call (csrc/dr_comm.cu): one recurrence launch, copy engines
        #             move S and the forecasts, no NCCL and no SM-resident collective — round-2 default
        #   "copy"  : round-1 path: NCCL all-reduce of S per chunk, forecasts by strided 2-D peer copies after the head kernel
        #   "kernel": as "copy", but the head kernel stores straight into the peers' tensors (good at 2 GPUs, slow at 8:
        #             49.0 vs 37.1 ms/step, profiles/r01_run_r / r01_run_l)
        #   "nccl"  : NCCL all-reduce + all-gather + interleave kernel
        #   "auto"  : "dma" where the tcgen05 engine runs (input_size <= 64), else "copy"
        self.gather_mode = "auto"
        if process_group is not None or (world or 1) > 1:
            import torch.distributed as dist
            rank = dist.get_rank(process_group) if rank is None else rank
            world = dist.get_world_size(process_group) if world is None else world
        self.rank, self.world = int(rank or 0), int(world or 1)
        if device is None:
            device = 0
            try:
                import torch
                if torch.cuda.is_available():
                    device = torch.cuda.current_device()
            except ImportError:
                pass
        self.device = int(device)
        self._lib = _lib.load()
        cfg = _lib.DrConfig(F=self.input_size, M=self.num_metrics, H=layout.H, Q=layout.Q,
                            dropout_p=self.dropout_p, engine=_lib.ENGINES[engine], device=self.device,
                            rank=self.rank, world=self.world, dtype=_lib.DTYPES[dtype])
        for i, q in enumerate(self.quantiles):
            cfg.quantiles[i] = q
        self._h = C.c_void_p()
        rc = self._lib.dr_create(C.byref(cfg), C.byref(self._h))
        if rc != _lib.DR_OK:
            msg = self._lib.dr_last_error(None).decode()
            if self.num_metrics < 2:
                # the reference fails here too: torch.stack([]) raises RuntimeError (qrnn.py:52)
                raise RuntimeError(msg)
            raise _lib.DeepRestError(rc, msg)
        self.m_local = self.num_metrics // self.world

    # ---- lifecycle -------------------------------------------------------------------
    def close(self):
        """Frees the handle.  For an expert-sharded handle that has run the library's exchange this is COLLECTIVE: every rank
        drains its own sharded forwards, the ranks synchronise (no peer is still copying into this rank), then the handle
        goes away.  The exchange arena itself is process-wide and stays (include/deeprest_b200.h, dr_comm_init)."""
        if getattr(self, "_h", None) and self._h.value and getattr(self, "_comm_shape", None):
            self._lib.dr_comm_detach(self._h)
            self._comm_shape = None
            try:
                import torch.distributed as dist
                if dist.is_initialized():
                    dist.barrier(group=self._pg)
            except Exception:
                pass
        if getattr(self, "_h", None) and self._h.value:
            self._lib.dr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def to(self, device):                         # estimate.py:60 — handle is already on its GPU
        return self

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        self.training = bool(mode)
        return self

    # ---- weights ---------------------------------------------------------------------
    def load_blob(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.float32).reshape(-1)
        _lib.check(self._h, self._lib.dr_load_weights(
            self._h, blob.ctypes.data_as(C.POINTER(C.c_float)), blob.size))
        return self

    def load_state_dict(self, sd):
        return self.load_blob(layout.blob_from_state_dict(sd, self.num_metrics, self.input_size))

    def blob(self):
        out = np.zeros(layout.blob_size(self.num_metrics, self.input_size), np.float32)
        _lib.check(self._h, self._lib.dr_get_weights(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), out.size))
        return out

    def state_dict(self):
        return layout.state_dict_from_blob(self.blob(), self.num_metrics, self.input_size)

    # ---- forward ---------------------------------------------------------------------
    def __call__(self, input_seq, out=None, borrow=False):
        return self.forward(input_seq, out=out, borrow=borrow)

    def forward(self, input_seq, out=None, borrow=False):
        """qrnn.py:28-56 in eval mode: [B,T,F] -> [B,T,M,Q].

        Expert-sharded handles (world > 1) take CUDA tensors and return the stacked forecasts on every rank.  The library
        keeps TWO result tensors per handle and alternates between them, so by default the result is cloned; pass
        ``borrow=True`` to get a view of the library's tensor instead — it is overwritten by the second-next forward on
        this handle (a list comprehension ``[model(x, borrow=True) for x in batches]`` would alias)."""
        if self.training and self.dropout_p > 0:
            raise NotImplementedError(
                "training-mode forward (dropout + autograd, qrnn.py:43 / estimate.py:70-74) is served by "
                "train_step(); call .eval() for inference")
        if _is_torch(input_seq) and input_seq.is_cuda:
            return self._forward_torch(input_seq, borrow, out)
        if self.world != 1:
            return self._forward_sharded_host(input_seq, out)
        x = np.ascontiguousarray(input_seq.detach().cpu().numpy() if _is_torch(input_seq) else input_seq,
                                 dtype=np.float32)
        if x.ndim != 3 or x.shape[2] != self.input_size:
            raise ValueError(f"input_seq must be [B,T,{self.input_size}], got {x.shape}")
        B, T, _ = x.shape
        if out is None:
            out = np.empty((B, T, self.num_metrics, layout.Q), np.float32)
        elif out.shape != (B, T, self.num_metrics, layout.Q) or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float32 [B,T,M,Q] array (pinned memory makes the D2H copy fast)")
        fp = C.POINTER(C.c_float)
        _lib.check(self._h, self._lib.dr_forward(self._h, x.ctypes.data_as(fp), B, T, out.ctypes.data_as(fp)))
        if _is_torch(input_seq):
            import torch
            return torch.from_numpy(out)
        return out

    def _bind_stream(self):
        import torch
        _lib.check(self._h, self._lib.dr_set_stream(
            self._h, C.c_void_p(torch.cuda.current_stream().cuda_stream), 1))

    # ---- expert-sharded forward through the library's own exchange (csrc/dr_comm.cu) ----
    def init_comm(self, max_windows, seq_len):
        """Collective: size this rank's arena for calls up to [max_windows, seq_len] and map the peers' arenas (CUDA IPC
        handles exchanged through the process group — the only thing torch.distributed does for the forward)."""
        import torch.distributed as dist
        if getattr(self, "_comm_shape", None):           # re-sizing: no peer may still be copying into this rank
            self._lib.dr_comm_detach(self._h)
            dist.barrier(group=self._pg)
        handle = (C.c_ubyte * 64)()
        _lib.check(self._h, self._lib.dr_comm_init(self._h, int(max_windows), int(seq_len), handle, None))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=self._pg)
        buf = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(handles))
        _lib.check(self._h, self._lib.dr_comm_attach(self._h, buf, None))
        self._comm_shape = (int(max_windows), int(seq_len))

    def _ensure_comm(self, B, T):
        shp = getattr(self, "_comm_shape", None)
        if shp is None or B > shp[0] or T != shp[1]:
            self.init_comm(B, T)

    def _forward_sharded_host(self, input_seq, out):
        """numpy / CPU tensors on a sharded handle: H2D of x, the exchange, and the D2H of THIS rank's forecast columns into
        ``out`` — the full [B,T,M,Q] host array (ranks of one host may share it, e.g. a shared-memory mapping) — all inside
        ``dr_forward_sharded``."""
        x = np.ascontiguousarray(input_seq.detach().cpu().numpy() if _is_torch(input_seq) else input_seq, dtype=np.float32)
        if x.ndim != 3 or x.shape[2] != self.input_size:
            raise ValueError(f"input_seq must be [B,T,{self.input_size}], got {x.shape}")
        B, T, _ = x.shape
        if out is None:
            out = np.zeros((B, T, self.num_metrics, layout.Q), np.float32)
        elif out.shape != (B, T, self.num_metrics, layout.Q) or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float32 [B,T,M,Q] array")
        self._ensure_comm(B, T)
        self._lib.dr_set_stream(self._h, None, 0)
        fp = C.POINTER(C.c_float)
        _lib.check(self._h, self._lib.dr_forward_sharded(self._h, x.ctypes.data_as(fp), B, T, out.ctypes.data_as(fp), None))
        return out

    def forward_async(self, x):
        """Expert-sharded handles: issue a forward without making torch's current stream wait for it; returns an object
        whose ``wait()`` orders the current stream after the result and returns the (borrowed) forecast tensor.  With the
        next batch issued before the previous result is consumed, the exchange tail of one forward runs under the
        recurrence of the next (two forwards in flight; the library alternates two result tensors)."""
        import torch
        if self.world == 1:
            raise ValueError("forward_async is for expert-sharded handles")
        x = x.contiguous()
        B, T, _ = x.shape
        self._bind_stream()
        self._ensure_comm(B, T)
        ptr, ticket = C.c_void_p(), C.c_int32()
        _lib.check(self._h, self._lib.dr_forward_sharded_issue_dev(self._h, x.data_ptr(), B, T, C.byref(ptr), C.byref(ticket)))
        owner, shape, dev = self, (B, T, self.num_metrics, layout.Q), x.device

        class _Pending:
            def wait(self_inner):
                owner._bind_stream()
                _lib.check(owner._h, owner._lib.dr_forward_sharded_wait(owner._h, ticket.value))

                class _Buf:
                    __cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr.value, False), "version": 2}
                return torch.as_tensor(_Buf(), device=dev)
        return _Pending()

    def _forward_torch(self, x, borrow=False, out=None):
        import torch
        if x.dtype != torch.float32 or x.dim() != 3 or x.shape[2] != self.input_size:
            raise ValueError(f"input_seq must be float32 [B,T,{self.input_size}]")
        x = x.contiguous()
        B, T, _ = x.shape
        self._bind_stream()
        if self.world == 1:
            if out is None:
                out = torch.empty((B, T, self.num_metrics, layout.Q), device=x.device, dtype=torch.float32)
            elif (not _is_torch(out) or not out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous()
                  or tuple(out.shape) != (B, T, self.num_metrics, layout.Q)):
                raise ValueError("out must be a contiguous float32 CUDA tensor [B,T,M,Q]")
            _lib.check(self._h, self._lib.dr_forward_dev(self._h, x.data_ptr(), B, T, out.data_ptr()))
            return out
        from .sharding import sharded_forward
        lib, h = self._lib, self._h

        # each phase binds the handle to the stream it is called on (the exchange runs on a side stream)
        def local_fn(xx, S, out_local):
            self._bind_stream()
            _lib.check(h, lib.dr_forward_local_dev(h, xx.data_ptr(), xx.shape[0], T, S.data_ptr(), out_local.data_ptr()))

        def heads_fn(S, out_local):
            self._bind_stream()
            _lib.check(h, lib.dr_forward_heads_dev(h, S.data_ptr(), out_local.shape[0], T, out_local.data_ptr()))

        def interleave_fn(gathered, out):
            self._bind_stream()
            _lib.check(h, lib.dr_interleave_dev(h, gathered.data_ptr(), out.shape[0], T, out.data_ptr()))

        def heads_p2p_fn(S, bn, ptrs, row0):
            self._bind_stream()
            _lib.check(h, lib.dr_forward_heads_p2p_dev(h, S.data_ptr(), bn, T, ptrs, self.world, row0))

        def scatter_fn(out_local, bn, ptrs, row0):
            self._bind_stream()
            _lib.check(h, lib.dr_scatter_forecasts_dev(h, out_local.data_ptr(), bn, T, ptrs, self.world, row0))

        peer = None
        mode = self.gather_mode
        if mode == "auto":
            mode = "dma" if (self.input_size <= 64 and self._engine != "ffma") else "copy"
        if mode == "dma":
            # the library's own exchange: ONE recurrence launch, partial sums of S moved by the copy engines and added by
            # the head kernel, forecasts scattered into every rank's tensor by 2-D peer copies (no NCCL, no SM-resident
            # collective)
            self._ensure_comm(B, T)
            ptr = C.c_void_p()
            _lib.check(h, lib.dr_forward_sharded_dev(h, x.data_ptr(), B, T, C.byref(ptr)))

            class _Buf:
                def __init__(self, p, shape):
                    self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (p, False), "version": 2}
            res = torch.as_tensor(_Buf(ptr.value, (B, T, self.num_metrics, layout.Q)), device=x.device)
            return res if borrow else res.clone()
        if mode == "kernel" and (self.input_size > 64 or self._engine == "ffma"):
            mode = "copy"                    # peer stores from K2 need the tcgen05 head kernel
        if mode == "copy":
            from .sharding import PeerBuffers
            if self._peer is None:
                self._peer = PeerBuffers(self._pg)
            res = sharded_forward(x, world=self.world, m_local=self.m_local, q=layout.Q,
                                  s_elems=lambda bn: lib.dr_s_elems(bn, T), local_fn=local_fn, heads_fn=heads_fn,
                                  interleave_fn=interleave_fn, group=self._pg, peer=self._peer, scatter_fn=scatter_fn)
            return res if borrow else res.clone()          # the symmetric-memory tensors alternate
        if mode == "kernel":
            from .sharding import PeerBuffers
            if self._peer is None:
                self._peer = PeerBuffers(self._pg)
            peer = self._peer
        res = sharded_forward(x, world=self.world, m_local=self.m_local, q=layout.Q,
                              s_elems=lambda bn: lib.dr_s_elems(bn, T), local_fn=local_fn, heads_fn=heads_fn,
                              interleave_fn=interleave_fn, group=self._pg, peer=peer, heads_p2p_fn=heads_p2p_fn)
        return res if (borrow or peer is None) else res.clone()

    # ---- steps either side of the path ----------------------------------------
    def forward_series(self, series, window_size, stride=1):
        """Forecast straight from the raw traffic series [N,F]: windows series[k*stride : k*stride+window_size]
        are formed on the device (utils.py:4-5 semantics incl. the dropped last window; estimate.py:85-86 uses
        stride = step_size).  Returns [n_windows, window_size, M, Q]."""
        s = np.ascontiguousarray(series.detach().cpu().numpy() if _is_torch(series) else series, np.float32)
        if s.ndim != 2 or s.shape[1] != self.input_size:
            raise ValueError(f"series must be [N,{self.input_size}]")
        n = self._lib.dr_series_windows(s.shape[0], window_size, stride)
        if n < 1:
            return np.empty((0, window_size, self.num_metrics, layout.Q), np.float32)
        out = np.empty((n, window_size, self.num_metrics, layout.Q), np.float32)
        fp = C.POINTER(C.c_float)
        _lib.check(self._h, self._lib.dr_forward_series(self._h, s.ctypes.data_as(fp), s.shape[0], window_size, stride,
                                                        out.ctypes.data_as(fp)))
        return out

    def set_denormalization(self, scales=None, clamp=1e-6):
        """Fuse estimate.py:96,101-102 into the head kernel: out = max(out, clamp) * range_m + min_m.
        ``scales`` = [(max-min, min), ...] per metric, exactly the list estimate.py:44-47 builds; None disables."""
        if scales is None:
            _lib.check(self._h, self._lib.dr_set_output_transform(self._h, None, None, 0.0))
            return self
        sc = np.ascontiguousarray([float(a) for a, _ in scales], np.float32)
        of = np.ascontiguousarray([float(b) for _, b in scales], np.float32)
        if sc.size != self.num_metrics:
            raise ValueError("one (range, min) pair per metric")
        fp = C.POINTER(C.c_float)
        _lib.check(self._h, self._lib.dr_set_output_transform(self._h, sc.ctypes.data_as(fp), of.ctypes.data_as(fp), float(clamp)))
        return self

    # ---- training ----------------------------------------------------------------------
    def train_step(self, inputs, labels, lr=1e-3, dropout_mask=None, seed=0):
        """One iteration of the reference training loop (estimate.py:67-74): train-mode forward
        (dropout on the GRU outputs), ``quantile_loss``, backward, ``Adam(lr)`` step.  Returns the
        loss (float).  ``dropout_mask`` ([M,B,T,2H] of 0/1) replays a mask for parity tests;
        otherwise a counter-based RNG keyed by ``seed`` draws it on the device."""
        if _is_torch(inputs) and inputs.is_cuda:
            return self._train_step_torch(inputs, labels, lr, dropout_mask, seed)
        x = np.ascontiguousarray(inputs.detach().cpu().numpy() if _is_torch(inputs) else inputs, np.float32)
        y = np.ascontiguousarray(labels.detach().cpu().numpy() if _is_torch(labels) else labels, np.float32)
        if x.ndim != 3 or x.shape[2] != self.input_size or y.shape != (x.shape[0], x.shape[1], self.num_metrics):
            raise ValueError("inputs must be [B,T,F] and labels [B,T,M]")
        B, T, _ = x.shape
        mptr = None
        if dropout_mask is not None:
            dm = np.ascontiguousarray(dropout_mask, np.uint8)
            if dm.shape != (self.num_metrics, B, T, 2 * layout.H):
                raise ValueError("dropout_mask must be [M,B,T,2H]")
            mptr = dm.ctypes.data_as(C.c_void_p)
        loss = C.c_float()
        fp = C.POINTER(C.c_float)
        _lib.check(self._h, self._lib.dr_train_step(self._h, x.ctypes.data_as(fp), y.ctypes.data_as(fp), B, T, mptr,
                                                    C.c_uint64(int(seed)), C.c_float(lr), C.byref(loss)))
        return float(loss.value)

    def _train_step_torch(self, inputs, labels, lr, dropout_mask, seed):
        """Device-resident variant (CUDA torch tensors, ``dr_train_step_dev`` on torch's current stream): nothing is copied;
        returns the loss as a 0-dim CUDA tensor.  The train-mode forecasts of the step are kept in ``self.train_outputs``."""
        import torch
        if self.world != 1:
            raise ValueError("sharded handles use train_step_sharded")
        x = inputs.contiguous()
        y = labels.contiguous()
        if x.dtype != torch.float32 or y.dtype != torch.float32 or x.dim() != 3 or x.shape[2] != self.input_size \
                or tuple(y.shape) != (x.shape[0], x.shape[1], self.num_metrics):
            raise ValueError("inputs must be float32 [B,T,F] and labels float32 [B,T,M]")
        B, T, _ = x.shape
        mask = None
        if dropout_mask is not None:
            mask = torch.as_tensor(np.ascontiguousarray(dropout_mask, np.uint8)).to(x.device)
        out = getattr(self, "train_outputs", None)
        if out is None or tuple(out.shape) != (B, T, self.num_metrics, layout.Q) or out.device != x.device:
            out = self.train_outputs = torch.empty((B, T, self.num_metrics, layout.Q), device=x.device, dtype=torch.float32)
        loss = torch.empty((), device=x.device, dtype=torch.float32)
        self._bind_stream()
        _lib.check(self._h, self._lib.dr_train_step_dev(self._h, x.data_ptr(), y.data_ptr(), B, T,
                                                        mask.data_ptr() if mask is not None else None,
                                                        C.c_uint64(int(seed)), C.c_float(lr), loss.data_ptr(), out.data_ptr()))
        return loss

    def train_step_sharded(self, inputs, labels, lr=1e-3, dropout_mask=None, seed=0, local_labels=False, micro_batch=None):
        """Expert-sharded training step (world > 1).  ``inputs`` [B,T,F] and ``labels`` [B,T,M] are the full tensors
        (replicated); each rank trains its own experts.  The library's step is a state machine that asks for three kinds
        of cross-rank sums (S, the loss scalar, the head adjoint); they run here through torch.distributed on the device
        buffers.  Returns the (global) loss."""
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", self.device)
        x = torch.as_tensor(inputs, dtype=torch.float32).to(dev).contiguous()
        lo, hi = self.rank * self.m_local, (self.rank + 1) * self.m_local
        if local_labels:                                  # labels already hold only this rank's columns [B,T,M_loc]
            y = torch.as_tensor(labels, dtype=torch.float32).to(dev).contiguous()
        else:
            y = torch.as_tensor(labels, dtype=torch.float32)[:, :, lo:hi].to(dev).contiguous()
        B, T, _ = x.shape
        mask = None
        if dropout_mask is not None:
            mask = torch.as_tensor(np.ascontiguousarray(dropout_mask, np.uint8)).to(dev)
        loss = torch.zeros((), device=dev, dtype=torch.float32)
        out = torch.empty((B, T, self.m_local, layout.Q), device=dev, dtype=torch.float32)
        self._bind_stream()
        # every rank must cut the batch into the same micro-batches (the cross-rank sums are per micro-batch): a size taken
        # from each rank's own free memory could differ, so sharded steps use an explicit one (default 128 windows = one tile)
        if micro_batch is None and "DR_TRAIN_MICROBATCH" not in __import__("os").environ:
            micro_batch = 128
        if micro_batch is not None:
            _lib.check(self._h, self._lib.dr_train_set_microbatch(self._h, int(micro_batch)))
        _lib.check(self._h, self._lib.dr_train_begin_dev(self._h, x.data_ptr(), y.data_ptr(), B, T,
                                                         mask.data_ptr() if mask is not None else None,
                                                         C.c_uint64(int(seed)), C.c_float(lr), loss.data_ptr(), out.data_ptr()))
        kind, ptr, count, dtype = C.c_int32(), C.c_void_p(), C.c_int64(), C.c_int32()

        class _Buf:                                   # a device buffer owned by the library, seen through the CUDA array interface
            def __init__(self, p, n, f64):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8" if f64 else "<f4", "data": (p, False), "version": 2}

        while True:
            _lib.check(self._h, self._lib.dr_train_advance(self._h, C.byref(kind), C.byref(ptr), C.byref(count), C.byref(dtype)))
            if kind.value == 0:
                break
            t = torch.as_tensor(_Buf(ptr.value, count.value, dtype.value == 1), device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self._pg)
        self._last_loss_tensor = loss
        return float(loss.item())

    def grads(self):
        """Gradients of the last ``train_step`` as a blob in ``state_dict`` order."""
        out = np.zeros(layout.blob_size(self.num_metrics, self.input_size), np.float32)
        _lib.check(self._h, self._lib.dr_get_grads(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), out.size))
        return out

    # ---- loss ------------------------------------------------------------------------
    def quantile_loss(self, outputs, labels):
        if _is_torch(outputs) and outputs.is_cuda:
            import torch
            self._bind_stream()
            B, T = outputs.shape[:2]
            loss = torch.empty((), device=outputs.device, dtype=torch.float32)
            _lib.check(self._h, self._lib.dr_quantile_loss_dev(
                self._h, outputs.contiguous().data_ptr(), labels.contiguous().data_ptr(), B, T, loss.data_ptr()))
            return loss
        o = np.ascontiguousarray(outputs.detach().cpu().numpy() if _is_torch(outputs) else outputs, np.float32)
        y = np.ascontiguousarray(labels.detach().cpu().numpy() if _is_torch(labels) else labels, np.float32)
        if o.ndim != 4 or y.shape != o.shape[:3]:
            raise ValueError("outputs must be [B,T,M,Q] and labels [B,T,M]")
        loss = C.c_float()
        fp = C.POINTER(C.c_float)
        _lib.check(self._h, self._lib.dr_quantile_loss(self._h, o.ctypes.data_as(fp), y.ctypes.data_as(fp),
                                                       o.shape[0], o.shape[1], C.byref(loss)))
        return np.float32(loss.value)

    # ---- helpers either side of the path -----------------------------------------------
    @staticmethod
    def normalization_minmax(M, split):
        """qrnn.py:69-75 — host-side min-max over the train split (identity if constant)."""
        head = np.asarray(M)[:split]
        lo, hi = head.min(), head.max()
        span = hi - lo
        return ((M - lo) / span if span != 0.0 else M), lo, hi

    # ---- diagnostics -------------------------------------------------------------------
    def debug_read(self, what, n):
        buf = np.empty(n, np.float32)
        _lib.check(self._h, self._lib.dr_debug_read(self._h, what.encode(), buf.ctypes.data_as(C.POINTER(C.c_float)), n))
        return buf

    def profile(self, enable=True):
        _lib.check(self._h, self._lib.dr_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        """(forwards recorded, sum of recurrence-kernel ms, sum of head-kernel ms) since profile(True)."""
        n, g, h = C.c_int32(), C.c_float(), C.c_float()
        _lib.check(self._h, self._lib.dr_profile_read(self._h, C.byref(n), C.byref(g), C.byref(h)))
        return n.value, g.value, h.value

    @property
    def launch_count(self):
        return int(self._lib.dr_launch_count(self._h))

    @property
    def last_engine(self):
        return self._lib.dr_last_engine(self._h).decode()
