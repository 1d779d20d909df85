#!/usr/bin/env python
"""bench.py — forecast windows/sec of the DeepRest estimator hot path on N B200s.

A "step" is one pass of the hot path (QuantileRNN.forward, eval mode, fp32) over one batch
of synthetic trace windows.  At N=1 the workload is BASELINE.json configs[1]:
64 services (M=128 experts) x 1024 windows x seq_len 288, F=64.  For N>1 the services are
sharded by service ID, 64 services per GPU (weak scaling); the one exchange the path needs (the
cross-expert sum S) and the gather of the forecasts run inside the library (dr_forward_sharded:
copy-engine transfers of partial sums and forecast columns).  The same JSON line also
carries: `train` (N=1) = BASELINE configs[2], one bf16 training step at 256 services x batch 4096;
`configs3` (N=8) = BASELINE configs[3], 1024 services over 8 GPUs; `mae_vs_reference` at every N.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line on rank 0 (contract in the task statement): `value` = whole-job
service-windows/s with inputs resident in HBM; `e2e` = the same through the C-ABI call with
pinned HOST buffers (H2D of x and D2H of the forecasts inside the timed region);
`roofline` for the recurrence kernel; `cpu_baseline` = the reference's CPU algorithm
(oracle/qrnn_torch_cpu.py, a port that keeps the reference's cost structure) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WSEED, XSEED = 11, 2021
METRIC = "forecast_windows_per_sec"
UNIT = "service-windows/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "ffma", "tcgen05"])
    ap.add_argument("--services", type=int, default=0, help="total services (default 64 per GPU)")
    ap.add_argument("--windows", type=int, default=1024)
    ap.add_argument("--seq-len", type=int, default=288)
    ap.add_argument("--features", type=int, default=64)
    ap.add_argument("--gather", default="auto", choices=["auto", "dma", "kernel", "copy", "nccl"],
                    help="multi-GPU exchange: dma = the library's own (dr_forward_sharded: S partials and forecasts by copy engines); "
                         "kernel / copy / nccl = the round-1 phase-call paths with an NCCL all-reduce of S")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=200.0, help="budget of the bounded CPU sample (BASELINE.md §3: 64 windows)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the BASELINE configs[2] training block (N=1)")
    ap.add_argument("--no-configs3", action="store_true", help="skip the BASELINE configs[3] block (N=8)")
    ap.add_argument("--configs3", action="store_true", help="run the configs[3] block (1024 services) at this world size too (testing)")
    ap.add_argument("--no-configs4", action="store_true", help="skip the BASELINE configs[4] sharded-training block (N=8)")
    ap.add_argument("--configs4", action="store_true", help="run the configs[4] block at this world size too (testing)")
    ap.add_argument("--train4-services", type=int, default=512)
    ap.add_argument("--train4-batch", type=int, default=1024)
    ap.add_argument("--train4-seq-len", type=int, default=1440)
    ap.add_argument("--train4-steps", type=int, default=2)
    ap.add_argument("--train-services", type=int, default=256)
    ap.add_argument("--train-batch", type=int, default=4096)
    ap.add_argument("--train-steps", type=int, default=3)
    ap.add_argument("--train-parity-batch", type=int, default=16)
    return ap.parse_args()


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")])

    def window(self, t0, t1):
        """keep the samples taken inside [t0, t1] (fall back to the nearest ones)"""
        inside = [r[1:] for r in self.rows if t0 <= r[0] <= t1]
        if not inside and self.rows:
            mid = 0.5 * (t0 + t1)
            inside = [r[1:] for r in sorted(self.rows, key=lambda r: abs(r[0] - mid))[:3]]
        self.rows = inside

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- helpers
T0 = time.perf_counter()


def log(msg):
    print(f"[bench +{time.perf_counter() - T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def host_cores():
    """Cores this process may actually use: affinity mask capped by the cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except Exception:
        pass
    return n


def algorithmic_flops(M_loc, B, T, F):
    """forward FLOPs per expert-window-step = 1536*F + 199,680."""
    return float(1536 * F + 199680) * M_loc * B * T


def algorithmic_hbm_bytes(M_loc, M, B, T, F):
    """SURVEY §8(d) fused ideal: x once + forecasts + weights (+ S round trip, two-pass)."""
    from deeprest_b200 import layout
    return 4.0 * B * T * F + 4.0 * B * T * M_loc * 3 + 4.0 * M_loc * layout.params_per_expert(F) + 2 * 4.0 * B * T * 256


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"tensor_tflops": float(d.get("bf16_tflops_sustained") or d["bf16_tflops"]),
                    "hbm_gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
        except Exception:
            pass
    return {"tensor_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md, sustained)"}


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def workload_config(S, M, B, T, F, world, engine="tcgen05"):
    """The `config` object of the JSON line — the reference arm prints the identical one (same workload, bounded sample)."""
    named = 1 if (world == 1 and S == 64) else 3 if (world == 8 and S == 1024) else None
    wl = (f"BASELINE configs[{named}]" if named is not None else "BASELINE configs[1]-shaped, weak-scaled") + \
         f": {S} services x {B} windows x seq_len {T}, F={F}, fp32 inference, {S // world} services per GPU"
    return {"workload": wl, "services": S, "experts": M, "windows": B, "seq_len": T, "features": F,
            "parallelism": f"expert-shard x{world}" if world > 1 else "single GPU", "engine": engine,
            "l2": (f"inputs+outputs per GPU (x {B * T * F * 4 / 1e6:.0f} MB, S {B * T * 256 * 4 / 1e6:.0f} MB, forecasts "
                   f"{B * T * M * 3 * 4 / 1e6:.0f} MB) exceed the 126 MB L2; no flush between steps needed")}


def cpu_sample(blob, M, F, x, budget_s, want_windows=64, chunk=16):
    """the reference on the host cores, eval + no_grad, all M experts, >= 64 windows in chunks of <= 16
    (bounded by `budget_s`: whole chunks only, at least one).  Also times a 1-window call: the operating point the
    reference's own evaluation loop uses (estimate.py:85-91 runs B = 1)."""
    from oracle.ref_runner import Runner
    cores = host_cores()
    r = Runner(blob, M, F, threads=cores)
    log(f"cpu baseline: {r.describe()}")
    t0 = time.perf_counter(); r.forward(x[:1]); t_one = time.perf_counter() - t0
    outs, secs, per_chunk = [], 0.0, []
    n_chunks = max(1, min(want_windows, x.shape[0]) // chunk)
    for ci in range(n_chunks):
        t0 = time.perf_counter()
        outs.append(r.forward(x[ci * chunk:(ci + 1) * chunk], chunk=chunk))
        dt = time.perf_counter() - t0
        secs += dt; per_chunk.append(dt)
        log(f"cpu baseline: chunk {ci + 1}/{n_chunks} ({chunk} windows) {dt:.1f}s")
        if ci + 1 < n_chunks and secs + dt > budget_s:
            log("cpu baseline: budget reached, stopping early")
            break
    out = np.concatenate(outs)
    return {"n": out.shape[0], "seconds": secs, "out": out, "cores": cores, "kind": r.kind, "what": r.describe(),
            "one_window_s": t_one, "chunk": chunk, "best_chunk_s": min(per_chunk), "runner": r}


# --------------------------------------------------------------------------- reference arm
def run_reference(args, rank):
    """The reference's own CPU implementation of the path on this box's host cores (task statement, tier ④): same `config`,
    `metric`, `unit` as our arm; each step is a bounded sample of that workload — n windows x k experts, ALL time steps."""
    if rank != 0:
        return
    from deeprest_b200 import synth
    from oracle.ref_runner import Runner
    N = args.gpus
    S = args.services or 64 * N
    M, B, T, F = 2 * S, args.windows, args.seq_len, args.features
    cores = host_cores()
    blob = synth.weights(WSEED, M, F)
    x = synth.windows(XSEED, min(B, 16), T, F)
    r = Runner(blob, M, F, threads=cores)
    log(f"reference arm: {r.describe()}")
    total_steps = args.steps + args.warmup
    per_step_budget = 170.0 / max(total_steps, 1)
    # calibrate on a small expert sample of one window (full T): GRU cost is linear in the expert count and one head's
    # stack/mean is linear in M, so a sample of k experts (with full-size M-1 stacks) costs k/M of the full forward
    k0 = min(M, 8)
    if r.kind == "reference":
        t_k0, _ = r.time_step_sampled(x[:1], k0)
        t_full_window = t_k0 * M / k0
    else:
        t0 = time.perf_counter(); r.forward(x[:1]); t_full_window = time.perf_counter() - t0
    log(f"reference arm: one full window x {M} experts ~ {t_full_window:.2f}s; budget per step {per_step_budget:.1f}s")
    if t_full_window <= per_step_budget or r.kind != "reference":
        k, n = M, int(max(1, min(x.shape[0], per_step_budget / t_full_window)))
    else:                                                  # one window of all experts does not fit: sample experts, never time steps
        n, k = 1, int(max(2, min(M, M * per_step_budget / t_full_window)))
    xs = np.ascontiguousarray(x[:n])

    def step():
        if k == M:
            t0 = time.perf_counter(); r.forward(xs, chunk=16); return time.perf_counter() - t0, 1.0
        return r.time_step_sampled(xs, k)

    for _ in range(args.warmup):
        step()
    tot, scale = 0.0, 1.0
    for _ in range(args.steps):
        dt, scale = step()
        tot += dt
    dt = tot / args.steps
    value = S * n / (dt * scale)
    sample = (f"{n} of {B} windows x {k} of {M} experts x all {T} time steps per step ({r.describe()}); "
              + ("every expert and head computed" if k == M else
                 f"expert sample: the bi-GRUs of {k} experts and {k} heads with full-size stacks of {M - 1} outputs are timed and "
                 f"the step time is scaled by M/k = {scale:.1f} (GRU cost and per-head stack cost are linear in the expert count)"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(S, M, B, T, F, N),
        "sample": {"windows_per_step": n, "experts_per_step": k, "time_steps": T, "scale_to_full_forward": scale},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": r.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- training block (N = 1)
def train_block(args, dev, peaks):
    """BASELINE configs[2]: one training step (dropout, pinball loss, backward, Adam) at 256 services x batch 4096 x T=288 in
    bf16 on one B200, device resident; parity of the bf16 engine against the fp32-parity engine on a sample; the reference's
    CPU training step beside it."""
    import torch
    from deeprest_b200 import QuantileRNN, layout, synth
    S, B, T, F = args.train_services, args.train_batch, args.seq_len, args.features
    M = 2 * S
    blk = {"workload": f"BASELINE configs[2]: training step, {S} services ({M} experts) x batch {B} x seq_len {T}, F={F}, bf16, 1xB200",
           "dtype": "bf16", "data": "synthetic"}
    blob = synth.weights(WSEED, M, F)
    m = QuantileRNN(F, M, dtype="bf16", device=dev.index)
    m.load_blob(blob)
    x = torch.from_numpy(synth.windows(XSEED, B, T, F)).to(dev)
    y = ((torch.arange(T, device=dev, dtype=torch.float32)[None, :, None] % 17) / 17.0).expand(B, T, M).contiguous()
    log(f"train: model ready ({M} experts), first step allocates the activation images")
    m.train_step(x, y, seed=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = m.launch_count
    e0.record()
    for i in range(args.train_steps):
        loss = m.train_step(x, y, seed=2 + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.train_steps
    flops = 3.0 * algorithmic_flops(M, B, T, F)               # forward + ~2x for the backward
    act_bytes = 5120.0 * 2 * M * B * T                         # bf16 images read+written per expert-window-step-direction
    blk.update({
        "ms_per_step": ms, "value": S * B / (ms * 1e-3), "unit": UNIT, "steps": args.train_steps, "warmup": 1,
        "gpu_launches_per_step": (m.launch_count - l0) // args.train_steps, "loss": float(loss), "engine": m.last_engine,
        "roofline": {"bound": "tensor", "algorithmic_flops_per_step": flops, "achieved": flops / (ms * 1e-3) / 1e12,
                     "peak": peaks["tensor_tflops"], "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / peaks["tensor_tflops"],
                     "hbm": {"activation_bytes_per_step": act_bytes, "achieved_gbs": act_bytes / (ms * 1e-3) / 1e9,
                             "peak_gbs": peaks["hbm_gbs"], "frac": act_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
                     "note": "3 x forward FLOPs (1536F + 199,680 per expert-window-step); single-pass bf16 tensor work, so frac is "
                             "directly the tensor-pipe share; the step is bound by the per-step latency of the two recurrences, "
                             "see DESIGN.md §7"},
    })
    log(f"train: {ms:.1f} ms/step = {blk['value']:.0f} service-windows/s, {blk['roofline']['achieved']:.0f} TFLOP/s algorithmic")
    # end to end through the host entry point (x, y pinned host -> H2D inside; loss D2H)
    try:
        xh = x.cpu().numpy(); yh = y.cpu().numpy()
        t0 = time.perf_counter(); m.train_step(xh, yh, seed=99); dt = time.perf_counter() - t0
        blk["e2e"] = {"ms_per_step": dt * 1e3, "value": S * B / dt, "unit": UNIT, "h2d_bytes_per_step": int(xh.nbytes + yh.nbytes),
                      "d2h_bytes_per_step": 4, "note": "dr_train_step with host buffers (pageable numpy), one step"}
        del xh, yh
    except Exception as exc:                                   # e.g. host memory
        blk["e2e"] = {"unavailable": repr(exc)}
    m.close()
    del x, y
    torch.cuda.empty_cache()
    # ---- parity sample: bf16 engine vs the fp32-parity engine (itself pinned to the reference autograd golden g5) ----
    Bs = args.train_parity_batch
    xs = torch.from_numpy(synth.windows(XSEED + 1, Bs, T, F)).to(dev)
    ys = ((torch.arange(T, device=dev, dtype=torch.float32)[None, :, None] % 13) / 13.0).expand(Bs, T, M).contiguous()
    mask = (torch.rand((M, Bs, T, 2 * layout.H), device=dev, generator=torch.Generator(device=dev).manual_seed(5)) >= 0.5).to(torch.uint8)
    res = {}
    for dt_name in ("bf16", "fp32"):
        mm = QuantileRNN(F, M, dtype=dt_name, device=dev.index)
        mm.load_blob(blob)
        lv = float(mm.train_step(xs, ys, dropout_mask=mask.cpu().numpy(), seed=0).item())
        res[dt_name] = (lv, mm.grads())
        mm.close()
        torch.cuda.empty_cache()
    g16, g32 = res["bf16"][1], res["fp32"][1]
    fam_err = {}
    pe = layout.params_per_expert(F)
    for name, (off, shape) in layout.expert_offsets(F).items():
        n = int(np.prod(shape))
        idx = (np.arange(M)[:, None] * pe + off + np.arange(n)[None, :])
        a, b = g16[idx], g32[idx]
        fam_err[name] = float((np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1e-20)).max())
    blk["parity_vs_fp32_engine"] = {
        "sample": f"all {M} experts x {Bs} windows x {T} steps, replayed dropout mask",
        "loss_bf16": res["bf16"][0], "loss_fp32": res["fp32"][0], "loss_abs_diff": abs(res["bf16"][0] - res["fp32"][0]),
        "grad_norm_rel_diff": float(abs(np.linalg.norm(g16) - np.linalg.norm(g32)) / np.linalg.norm(g32)),
        "grad_rel_l2_err": float(np.linalg.norm(g16 - g32) / np.linalg.norm(g32)),
        "worst_tensor_err_over_tensor_max": max(fam_err.values()), "stated_tolerance": "loss 5e-4, every gradient tensor within 1e-2 of its max",
        "within_tolerance": bool(abs(res["bf16"][0] - res["fp32"][0]) < 5e-4 and max(fam_err.values()) <= 1e-2)}
    del xs, ys, mask
    torch.cuda.empty_cache()
    # ---- the reference's CPU training step (estimate.py:67-74) at a shape its memory allows ----
    if not args.no_cpu_baseline:
        try:
            from oracle.ref_runner import Runner
            Mc, Bc = 32, 8                                     # autograd keeps M stacks of M-1 outputs: 512 experts would need ~77 GB per window
            rr = Runner(synth.weights(WSEED, Mc, F), Mc, F, threads=host_cores())
            xc, yc = synth.windows(XSEED, Bc, T, F), synth.labels(7, Bc, T, Mc)
            rr.train_step(xc, yc)
            t0 = time.perf_counter(); rr.train_step(xc, yc); dtc = time.perf_counter() - t0
            blk["cpu_baseline"] = {"value": (Mc // 2) * Bc / dtc, "unit": UNIT, "cores": host_cores(), "kind": rr.kind,
                                   "sample": f"one training step of {Mc} experts ({Mc // 2} services) x {Bc} windows x {T} steps "
                                             f"({dtc:.2f} s; {rr.describe()}); the reference's autograd keeps M stacks of M-1 GRU outputs, "
                                             f"so {M} experts do not fit host memory even at batch 1 — its cost per expert-window grows with M"}
        except Exception as exc:
            blk["cpu_baseline"] = {"unavailable": repr(exc)}
    return blk


# --------------------------------------------------------------------------- sharded training block (N = 8)
def train_sharded_block(args, rank, world, dev, peaks):
    """BASELINE configs[4]: long-horizon training, seq_len 1440, 512 services (1024 experts) sharded over the GPUs by
    service ID, bf16.  Experts, Adam moments and gradients are expert-local (no gradient all-reduce exists); the three
    cross-rank sums of a step (S, G-bar per micro-batch, the loss scalar) run through NCCL between the library's phases."""
    import torch
    import torch.distributed as dist
    from deeprest_b200 import QuantileRNN, synth
    S, B, T, F = args.train4_services, args.train4_batch, args.train4_seq_len, args.features
    M = 2 * S
    M_loc = M // world
    lo, hi = rank * M_loc, (rank + 1) * M_loc
    m = QuantileRNN(F, M, dtype="bf16", device=dev.index, process_group=dist.group.WORLD, rank=rank, world=world)
    m.load_blob(synth.weights(WSEED, M, F, experts=(lo, hi)))
    x = torch.from_numpy(synth.windows(XSEED, B, T, F)).to(dev)
    y = ((torch.arange(T, device=dev, dtype=torch.float32)[None, :, None] % 17) / 17.0).expand(B, T, M_loc).contiguous()
    loss = m.train_step_sharded(x, y, seed=1, local_labels=True)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.train4_steps):
        loss = m.train_step_sharded(x, y, seed=2 + i, local_labels=True)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.train4_steps
    flops = 3.0 * algorithmic_flops(M, B, T, F)
    named = (S == 512 and T == 1440 and world == 8)
    blk = {"workload": ("BASELINE configs[4]" if named else "BASELINE configs[4]-shaped") +
                       f": long-horizon training step, {S} services ({M} experts) sharded x{world} ({M_loc} experts per GPU), batch {B}, seq_len {T}, F={F}, bf16",
           "dtype": "bf16", "data": "synthetic", "ms_per_step": ms, "value": S * B / (ms * 1e-3), "unit": UNIT, "steps": args.train4_steps,
           "warmup": 1, "loss": float(loss), "engine": m.last_engine, "parallelism": f"expert-shard x{world}",
           "roofline": {"bound": "tensor", "algorithmic_flops_per_step": flops, "achieved": flops / (ms * 1e-3) / 1e12,
                        "peak": peaks["tensor_tflops"] * world, "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / (peaks["tensor_tflops"] * world)},
           "collectives": "all-reduce of S and of the head adjoint G-bar per micro-batch, and of the loss scalar (NCCL, torch.distributed); no gradient all-reduce",
           "parity": "tests/test_gpu_train.py::test_bf16_long_horizon_training_parity (T=1440) and tests/test_gpu_multi.py::test_two_gpu_sharded_bf16_train_step_matches_oracle"}
    m.close()
    del x, y
    torch.cuda.empty_cache()
    dist.barrier()
    return blk


# --------------------------------------------------------------------------- our arm
def shared_host_tensor(shape, rank, world, tag):
    """One pinned host tensor shared by the ranks of this box (POSIX shared memory): every rank's D2H lands in its own
    columns of the same [B,T,M,Q] array.  Falls back to a private pinned tensor per rank when /dev/shm is too small."""
    import torch
    import torch.distributed as dist
    n = int(np.prod(shape)) * 4
    path = f"/dev/shm/deeprest_b200_{tag}_{os.environ.get('MASTER_PORT', '0')}"
    ok = torch.zeros(1, device="cuda")
    if rank == 0:
        try:
            st = os.statvfs("/dev/shm")
            if st.f_bavail * st.f_frsize < n + (64 << 20):
                raise OSError("not enough space in /dev/shm")
            with open(path, "wb") as f:
                f.truncate(n)
            ok += 1
        except OSError as exc:
            log(f"shared host tensor unavailable ({exc}); every rank uses a private pinned tensor")
    if world > 1:
        dist.broadcast(ok, 0)
    if float(ok.item()) < 1:
        t = torch.empty(tuple(shape), dtype=torch.float32, pin_memory=True)
        return t.numpy(), True, False, (lambda: None)
    arr = np.memmap(path, dtype=np.float32, mode="r+", shape=tuple(shape))
    rt = torch.cuda.cudart()
    rc = rt.cudaHostRegister(arr.ctypes.data, n, 0)
    pinned = (int(rc) == 0) if not isinstance(rc, tuple) else (int(rc[0]) == 0)
    if world > 1:
        dist.barrier()
    if rank == 0:
        os.unlink(path)                                        # the mappings keep it alive

    def release():
        # the registration must be dropped BEFORE the mapping goes away: a munmap'ed but still registered range stays in the
        # CUDA address space and a later large cudaMalloc that lands on it fails with cudaErrorAlreadyMapped
        if pinned:
            rt.cudaHostUnregister(arr.ctypes.data)
    return arr, pinned, True, release


def measure_inference(args, rank, world, dev, S, B, T, F, peaks, tag):
    """device-resident throughput, end-to-end throughput, K1 roofline and a parity sample for one inference workload"""
    import torch
    from deeprest_b200 import QuantileRNN, layout, synth
    M = 2 * S
    M_loc = M // world
    lo, hi = rank * M_loc, (rank + 1) * M_loc
    pg = None
    if world > 1:
        import torch.distributed as dist
        pg = dist.group.WORLD
    blob = synth.weights(WSEED, M, F, experts=(lo, hi))
    x_host = torch.empty((B, T, F), dtype=torch.float32, pin_memory=True)
    x_host.numpy()[...] = synth.windows(XSEED, B, T, F)
    model = QuantileRNN(input_size=F, num_metrics=M, engine=args.engine, device=dev.index, process_group=pg, rank=rank, world=world).eval()
    model.load_blob(blob)
    model.gather_mode = args.gather
    x_dev = x_host.to(dev)
    log(f"[{tag}] model ready: M={M} (local {M_loc}) B={B} T={T} F={F}")

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        import torch.distributed as dist
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # one GPU: the forecasts go into two preallocated tensors used alternately (no allocator activity inside the timed region)
    outs = [torch.empty((B, T, M, layout.Q), device=dev, dtype=torch.float32) for _ in range(2)] if world == 1 else None
    flip = [0]

    def fwd():
        if world > 1:
            return model(x_dev, borrow=True)
        flip[0] ^= 1
        return model(x_dev, out=outs[flip[0]])

    def run_steps(k):
        """k forwards; sharded handles keep two batches in flight (issue n+1, then consume n), as a serving loop would"""
        if world == 1:
            for _ in range(k):
                o = fwd()
            return o
        prev = model.forward_async(x_dev)
        for _ in range(k - 1):
            cur = model.forward_async(x_dev)
            prev.wait()
            prev = cur
        return prev.wait()
    for i in range(max(args.warmup, 3)):
        out = fwd()
        if i == 0:
            torch.cuda.synchronize()
            log(f"[{tag}] first forward done (engine {model.last_engine})")
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index) as clocks:
        for _ in range(2):                    # nvidia-smi needs a moment to start: keep the GPU busy meanwhile
            out = fwd()
        barrier()
        model.profile(True)
        launches0 = model.launch_count
        t_wall0 = time.time()
        import gc
        gc.disable()
        ev0.record()
        out = run_steps(args.steps)
        ev1.record()
        gc.enable()
        barrier()
        t_wall1 = time.time()
        time.sleep(0.05)
    clocks.window(t_wall0, t_wall1)
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = model.launch_count - launches0
    n_prof, gru_ms_sum, head_ms_sum = model.profile_read()
    model.profile(False)
    ms_step = ms_total / args.steps
    value = S * B / (ms_step * 1e-3)
    log(f"[{tag}] device-resident: {ms_step:.2f} ms/step, recurrence kernel {gru_ms_sum / args.steps:.2f} ms/step in "
        f"{n_prof // max(args.steps, 1)} launch(es)")

    # ---- end to end through the host entry point: pinned x in, forecasts out (each rank its own columns of ONE host tensor) ----
    h2d = x_host.numel() * 4
    x_np = x_host.numpy()
    if world == 1:
        out_host = torch.empty((B, T, M, layout.Q), dtype=torch.float32, pin_memory=True)
        out_np, pinned, shared, release_host = out_host.numpy(), True, False, (lambda: None)
    else:
        out_np, pinned, shared, release_host = shared_host_tensor((B, T, M, layout.Q), rank, world, tag)
    d2h = B * T * M_loc * layout.Q * 4

    def e2e_step():
        model(x_np, out=out_np)                # dr_forward / dr_forward_sharded: H2D + kernels + D2H of this rank's columns, synchronous
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    barrier()
    log(f"[{tag}] e2e: {e2e_s * 1e3:.2f} ms/step")

    gru_ms = gru_ms_sum / max(args.steps, 1)
    flops = algorithmic_flops(M_loc, B, T, F)
    achieved = flops / (gru_ms * 1e-3) / 1e12 if gru_ms > 0 else 0.0
    traffic = ncu_traffic()
    roofline = {
        "bound": "tensor", "kernel": f"bi-GRU recurrence ({model.last_engine} engine)",
        "achieved": achieved, "peak": peaks["tensor_tflops"], "unit": "TFLOP/s",
        "frac": achieved / peaks["tensor_tflops"], "traffic": traffic.get(model.last_engine),
        "peak_source": peaks["source"], "kernel_ms": gru_ms, "head_kernel_ms": head_ms_sum / max(args.steps, 1),
        "kernel_share_of_step": gru_ms / ms_step if ms_step else None,
        "launches_per_step": n_prof // max(args.steps, 1),
        "algorithmic_flops_per_launch": flops,
        "algorithmic_hbm_bytes_per_launch": algorithmic_hbm_bytes(M_loc, M, B, T, F),
        "hbm_gbs_if_ideal_bytes": algorithmic_hbm_bytes(M_loc, M, B, T, F) / (gru_ms * 1e-3) / 1e9 if gru_ms > 0 else None,
        "note": ("fp32 parity needs split-fp16 operands: the tcgen05 engine issues 3 tensor passes per "
                 "algorithmic FLOP, so frac <= 1/3 by construction; the FFMA engine runs on CUDA cores")
    }
    cfg = workload_config(S, M, B, T, F, world, model.last_engine)
    res = {"value": value, "ms_per_step": ms_step, "config": cfg, "batch_windows_per_sec": B / (ms_step * 1e-3), "clocks": clocks.summary(), "gpu_launches": int(launches),
           "roofline": roofline, "batches_in_flight": 1 if world == 1 else 2,
           "e2e": {"value": S * B / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                   "ms_per_step": e2e_s * 1e3, "per_rank": {"h2d_bytes": h2d, "d2h_bytes": d2h},
                   "host_output_pinned": bool(pinned), "host_output_shared_by_ranks": bool(shared),
                   "note": ("C-ABI dr_forward with pinned host buffers" if world == 1 else
                            "C-ABI dr_forward_sharded per rank: x (replicated) H2D, and every rank's own forecast columns D2H into one "
                            "shared pinned host tensor [B,T,M,Q], chunk by chunk under the compute")}}
    # ---- parity of what was timed (every world size): rank 0's last stacked forecasts vs the reference on the host ----
    final = out.detach().cpu().numpy() if rank == 0 else None
    e2e_final = np.array(out_np[:4]) if rank == 0 else None
    model.close()
    torch.cuda.synchronize()
    release_host()
    del out_np
    del x_dev
    torch.cuda.empty_cache()
    barrier()
    return res, final, e2e_final, x_host


def parity_block(args, S, B, T, F, final, e2e_final, x_np, world, cs=None):
    """mae_vs_reference for the JSON line.  One GPU at <= 256 experts: every expert and head on the CPU-baseline windows.
    More experts: the reference's arithmetic for a sample of 16 experts on 2 windows (all experts' GRUs, sampled heads)."""
    from deeprest_b200 import synth
    M = 2 * S
    if cs is not None:
        n, ref, what = cs["n"], cs["out"], f"all {M} experts, first {cs['n']} windows ({cs['kind']})"
        ours = final[:n]
    else:
        from oracle.ref_runner import Runner
        r = Runner(synth.weights(WSEED, M, F), M, F, threads=host_cores())
        n = 2 if M <= 512 else 1                               # every expert's bi-GRU runs on the host for the mean
        if r.kind == "reference":
            ids = sorted(set(int(v) for v in np.linspace(0, M - 1, 16)))
            ref = r.forward_sampled(x_np[:n], ids)
            ours = final[:n][:, :, ids]
            what = f"{len(ids)} experts spread over all {world} ranks' shards x first {n} windows, all {M} experts' GRUs in the mean (reference module)"
        else:
            ref = r.forward(x_np[:n])
            ours = final[:n]
            what = f"all {M} experts, first {n} windows (port)"
    err = np.abs(ours - ref)
    blk = {"mae": float(err.mean()), "max_abs": float(err.max()),
           "allclose_rtol1e-4_atol1e-6": bool(np.all(err <= 1e-6 + 1e-4 * np.abs(ref))), "windows_compared": int(n), "sample": what}
    if e2e_final is not None:
        k = min(n, e2e_final.shape[0])
        blk["e2e_host_tensor_max_abs_vs_device_result"] = float(np.abs(e2e_final[:k] - final[:k]).max())
    return blk


def run_ours(args, rank, world, local_rank):
    import torch
    N = args.gpus
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    S = args.services or 64 * N
    B, T, F = args.windows, args.seq_len, args.features
    peaks = measured_peaks()

    blk4 = None
    # (runs first: it needs the most device memory, before the inference models and their exchange arenas exist)
    # ---- BASELINE configs[4]: long-horizon bf16 training, experts sharded over the GPUs ----
    if world > 1 and (world == 8 or args.configs4) and not args.no_configs4:
        try:
            blk4 = train_sharded_block(args, rank, world, dev, peaks)
        except Exception as exc:
            import traceback
            log("configs4 block failed:\n" + traceback.format_exc())
            blk4 = {"unavailable": repr(exc)}
    blk3 = None
    # (before the main measurement: the exchange arena of a device is allocated once per process and reused — the model with
    #  the most experts must come first, see csrc/dr_comm.cu::GlobalArena)
    # ---- BASELINE configs[3] (1024 services over 8 GPUs) as a second, labelled block of the same line ----
    if ((world == 8 and S != 1024) or args.configs3) and not args.no_configs3:
        r3, f3, e3, xh3 = measure_inference(args, rank, world, dev, 1024, B, T, F, peaks, "configs3")
        blk = {"metric": METRIC, "value": r3["value"], "unit": UNIT, "ms_per_step": r3["ms_per_step"], "config": r3["config"],
               "e2e": r3["e2e"], "roofline": r3["roofline"], "gpu_launches": r3["gpu_launches"], "scaling_note":
               "BASELINE configs[3] itself: 1024 services sharded over 8 GPUs (128 services = 256 experts per GPU)"}
        if rank == 0 and not args.no_parity:
            try:
                blk["mae_vs_reference"] = parity_block(args, 1024, B, T, F, f3, e3, xh3.numpy(), world)
            except Exception as exc:
                blk["mae_vs_reference"] = {"unavailable": repr(exc)}
        blk3 = blk
    res, final, e2e_final, x_host = measure_inference(args, rank, world, dev, S, B, T, F, peaks, "main")
    line = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": N, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": res["config"], "clocks": res["clocks"], "e2e": res["e2e"], "gpu_launches": res["gpu_launches"],
        "roofline": res["roofline"], "batch_windows_per_sec": res["batch_windows_per_sec"],
        "batches_in_flight": res["batches_in_flight"],
    }
    if blk4 is not None:
        line["configs4"] = blk4
    if blk3 is not None:
        line["configs3"] = blk3
    if rank == 0:
        M = 2 * S
        cs = None
        if world == 1 and not args.no_cpu_baseline:
            from deeprest_b200 import synth
            cs = cpu_sample(synth.weights(WSEED, M, F), M, F, x_host.numpy(), args.cpu_seconds)
            line["cpu_baseline"] = {
                "value": S * cs["n"] / cs["seconds"], "unit": UNIT, "cores": cs["cores"], "kind": cs["kind"],
                "sample": (f"first {cs['n']} of {B} windows x all {M} experts in chunks of {cs['chunk']} windows, eval + no_grad "
                           f"({cs['seconds']:.1f} s; {cs['what']}; BASELINE.md §3)"),
                "operating_points": {"chunk_of_16_windows_best": S * cs["chunk"] / cs["best_chunk_s"],
                                     "one_window_per_call_as_estimate_py_evaluates": S * 1 / cs["one_window_s"], "unit": UNIT}}
        if not args.no_parity:
            try:
                line["mae_vs_reference"] = parity_block(args, S, B, T, F, final, e2e_final, x_host.numpy(), world, cs)
            except Exception as exc:
                line["mae_vs_reference"] = {"unavailable": repr(exc)}
    # ---- BASELINE configs[2]: the training step (one GPU) ----
    if world == 1 and not args.no_train:
        try:
            line["train"] = train_block(args, dev, peaks)
        except Exception as exc:
            import traceback
            log("train block failed:\n" + traceback.format_exc())
            line["train"] = {"unavailable": repr(exc)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if args.gpus != 1 or world != 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
