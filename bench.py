#!/usr/bin/env python
"""bench.py — forecast windows/sec of the DeepRest estimator hot path on N B200s.

A "step" is one pass of the hot path (QuantileRNN.forward, eval mode, fp32) over one batch
of synthetic trace windows.  At N=1 the workload is BASELINE.json configs[1]:
64 services (M=128 experts) x 1024 windows x seq_len 288, F=64.  For N>1 the services are
sharded by service ID, 64 services per GPU (weak scaling), with the one all-reduce of the
cross-expert sum S and the all-gather of the forecasts the path needs.

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
    ap.add_argument("--gather", default="auto", choices=["auto", "kernel", "copy", "nccl"],
                    help="multi-GPU forecast gather: K2 peer stores / DMA-engine 2-D peer copies / NCCL all-gather")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the bounded CPU sample")
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


def cpu_sample(blob, M, F, x, budget_s):
    """Time the reference-algorithm CPU port on a bounded sample: all M experts, the first n windows."""
    import torch
    from oracle.qrnn_torch_cpu import TorchCpuPort
    cores = host_cores()
    torch.set_num_threads(cores)
    port = TorchCpuPort(blob, M, F)
    log(f"cpu baseline: calibrating on 1 window x {M} experts with {cores} threads")
    t0 = time.perf_counter()
    ref = port.forward(x[:1])
    t1 = time.perf_counter() - t0
    n = int(max(1, min(16, budget_s / max(t1, 1e-3), x.shape[0])))
    dt = t1
    if n > 1:
        log(f"cpu baseline: 1 window took {t1:.2f}s -> timing {n} windows")
        t0 = time.perf_counter()
        ref = port.forward(x[:n])
        dt = time.perf_counter() - t0
    log(f"cpu baseline: {n} windows in {dt:.2f}s")
    return {"n": n, "seconds": dt, "out": ref, "cores": cores, "torch": torch.__version__}


# --------------------------------------------------------------------------- reference arm
def run_reference(args, rank):
    if rank != 0:
        return
    from deeprest_b200 import synth
    N = args.gpus
    S = args.services or 64 * N
    M, B, T, F = 2 * S, args.windows, args.seq_len, args.features
    blob = synth.weights(WSEED, M, F)
    x = synth.windows(XSEED, min(B, 16), T, F)
    import torch
    from oracle.qrnn_torch_cpu import TorchCpuPort
    cores = host_cores()
    torch.set_num_threads(cores)
    port = TorchCpuPort(blob, M, F)
    log(f"reference arm: {cores} threads, calibrating")
    total_steps = args.steps + args.warmup
    per_step_budget = 150.0 / max(total_steps, 1)
    # The reference's cost is linear in windows and in T, but quadratic in the expert count (its stack/mean), so one
    # full window can already exceed the budget when many GPUs' worth of experts run on one host: calibrate on a short
    # prefix of one window, then bound the sample in windows and, only if one full window does not fit, in time steps.
    Tc = min(T, 8)
    t0 = time.perf_counter(); port.forward(x[:1, :Tc]); tc = time.perf_counter() - t0
    t_full_window = tc * T / Tc
    log(f"reference arm: 1 window x {M} experts x {Tc} steps = {tc:.2f}s -> full window ~{t_full_window:.1f}s")
    if t_full_window <= per_step_budget:
        Ts, n = T, int(max(1, min(x.shape[0], per_step_budget / t_full_window)))
    else:
        n, Ts = 1, int(max(Tc, min(T, T * per_step_budget / t_full_window)))
    xs = np.ascontiguousarray(x[:n, :Ts])
    for _ in range(args.warmup):
        port.forward(xs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        port.forward(xs)
    dt = (time.perf_counter() - t0) / args.steps
    value = S * n / dt * (Ts / T)                    # windows of the full seq_len per second (linear in T)
    sample = (f"{n} of {B} windows x all {M} experts x {Ts} of {T} time steps per step (reference algorithm, torch "
              f"{torch.__version__} CPU; cost is linear in windows and time steps, value scaled by {Ts}/{T})")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[1]-shaped: {S} services x {B} windows x seq_len {T}, F={F}, fp32 inference",
                   "services": S, "experts": M, "windows_per_step": n, "steps_per_window_sampled": Ts, "seq_len": T, "features": F},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch
    from deeprest_b200 import QuantileRNN, layout, synth
    N = args.gpus
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    S = args.services or 64 * N
    M, B, T, F = 2 * S, args.windows, args.seq_len, args.features
    M_loc = M // world
    lo, hi = rank * M_loc, (rank + 1) * M_loc

    blob = synth.weights(WSEED, M, F, experts=(lo, hi))
    x_host = torch.empty((B, T, F), dtype=torch.float32, pin_memory=True)
    x_host.numpy()[...] = synth.windows(XSEED, B, T, F)
    out_host = torch.empty((B, T, M, layout.Q), dtype=torch.float32, pin_memory=True)

    model = QuantileRNN(input_size=F, num_metrics=M, engine=args.engine, device=local_rank,
                        process_group=pg, rank=rank, world=world).eval()
    model.load_blob(blob)
    model.gather_mode = args.gather
    x_dev = x_host.to(dev)
    log(f"model ready: M={M} (local {M_loc}) B={B} T={T} F={F}")

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

    # ---- device-resident throughput (`value`) ----
    for i in range(max(args.warmup, 3)):
        out = model(x_dev)
        if i == 0:
            torch.cuda.synchronize()
            log(f"first forward done (engine {model.last_engine})")
    barrier()
    log("warm-up done")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        for _ in range(2):                    # nvidia-smi needs a moment to start: keep the GPU busy meanwhile
            out = model(x_dev)
        barrier()
        model.profile(True)
        launches0 = model.launch_count
        t_wall0 = time.time()
        ev0.record()
        for _ in range(args.steps):
            out = model(x_dev)
        ev1.record()
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
    log(f"device-resident: {ms_step:.2f} ms/step, recurrence kernel {gru_ms_sum / args.steps:.2f} ms/step in "
        f"{n_prof // max(args.steps, 1)} launch(es), head kernel {head_ms_sum / args.steps:.2f} ms/step")

    # ---- end to end through the public API with host buffers (`e2e`) ----
    h2d = x_host.numel() * 4
    if world == 1:
        x_np, out_np = x_host.numpy(), out_host.numpy()
        d2h = out_host.numel() * 4

        def e2e_step():
            model(x_np, out=out_np)            # C-ABI dr_forward: H2D + kernels + D2H, synchronous
    else:
        d2h = out_host.numel() * 4

        def e2e_step():
            x_dev.copy_(x_host, non_blocking=True)
            o = model(x_dev)
            out_host.copy_(o, non_blocking=True)
            torch.cuda.synchronize()
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    e2e_value = S * B / e2e_s
    log(f"e2e: {e2e_s * 1e3:.2f} ms/step")

    # ---- roofline of the dominant kernel (the bi-GRU recurrence) ----
    peaks = measured_peaks()
    # per STEP (a step may run as several chunk launches; n_prof counts launches)
    gru_ms = gru_ms_sum / max(args.steps, 1)
    head_ms = head_ms_sum / max(args.steps, 1)
    flops = algorithmic_flops(M_loc, B, T, F)
    achieved = flops / (gru_ms * 1e-3) / 1e12 if gru_ms > 0 else 0.0
    traffic = ncu_traffic()
    roofline = {
        "bound": "tensor", "kernel": f"bi-GRU recurrence ({model.last_engine} engine)",
        "achieved": achieved, "peak": peaks["tensor_tflops"], "unit": "TFLOP/s",
        "frac": achieved / peaks["tensor_tflops"], "traffic": traffic.get(model.last_engine),
        "peak_source": peaks["source"], "kernel_ms": gru_ms, "head_kernel_ms": head_ms,
        "kernel_share_of_step": gru_ms / ms_step if ms_step else None,
        "launches_per_step": n_prof // max(args.steps, 1),
        "algorithmic_flops_per_launch": flops,
        "algorithmic_hbm_bytes_per_launch": algorithmic_hbm_bytes(M_loc, M, B, T, F),
        "hbm_gbs_if_ideal_bytes": algorithmic_hbm_bytes(M_loc, M, B, T, F) / (gru_ms * 1e-3) / 1e9 if gru_ms > 0 else None,
        "note": ("fp32 parity needs split-fp16 operands: the tcgen05 engine issues 3 tensor passes per "
                 "algorithmic FLOP, so frac <= 1/3 by construction; the FFMA engine runs on CUDA cores")
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": (f"BASELINE configs[{1 if (world == 1 and S == 64) else 3 if (world == 8 and S == 1024) else 1}]"
                                + ("" if (world == 1 and S == 64) or (world == 8 and S == 1024) else "-shaped, weak-scaled")
                                + f": {S} services x {B} windows x seq_len {T}, F={F}, fp32 inference, {S // world} services per GPU"),
                   "services": S, "experts": M, "windows": B, "seq_len": T, "features": F,
                   "parallelism": f"expert-shard x{world}" if world > 1 else "single GPU",
                   "batch_windows_per_sec": B / (ms_step * 1e-3), "engine": model.last_engine,
                   "l2": (f"inputs+outputs per GPU (x {B * T * F * 4 / 1e6:.0f} MB, S {B * T * 256 * 4 / 1e6:.0f} MB, forecasts "
                          f"{B * T * M * 3 * 4 / 1e6:.0f} MB) exceed the 126 MB L2; no flush between steps needed")},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }

    # ---- CPU baseline + MAE vs the reference algorithm (rank 0, N=1 only) ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        full_blob = blob
        cs = cpu_sample(full_blob, M, F, x_host.numpy(), args.cpu_seconds)
        n = cs["n"]
        ours = out.detach().cpu().numpy()[:n]
        err = np.abs(ours - cs["out"])
        line["cpu_baseline"] = {
            "value": S * n / cs["seconds"], "unit": UNIT, "cores": cs["cores"], "kind": "port",
            "sample": f"first {n} of {B} windows x all {M} experts, one call ({cs['seconds']:.2f} s; reference algorithm "
                      f"restated on torch {cs['torch']} CPU incl. its O(M^2) stack/mean, oracle/qrnn_torch_cpu.py)"}
        line["mae_vs_reference"] = {"mae": float(err.mean()), "max_abs": float(err.max()),
                                    "allclose_rtol1e-4_atol1e-6": bool(np.all(err <= 1e-6 + 1e-4 * np.abs(cs["out"]))),
                                    "windows_compared": n}
    if rank == 0:
        print(json.dumps(line), flush=True)
    model.close()
    if world > 1:
        import torch.distributed as dist
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
