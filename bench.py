#!/usr/bin/env python
"""Benchmark of the PINN residual-loss + gradient hot path (BASELINE.json metric:
collocation points/sec per train step).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C2]

One "step" = one Adam train step (software.py:412-414): residual loss over ALL
collocation + boundary points, its gradient, and the Adam update.  N>1 is launched
by torchrun, one rank per GPU, weak scaling (each rank holds the workload's point
count; one fused NCCL allreduce per step).  Rank 0 prints ONE JSON line.
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
sys.path.insert(0, ROOT)

METRIC = "collocation points/sec per train step (residual+grad)"
UNIT = "points/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--n-col", type=int, default=None)
    ap.add_argument("--cpu-points", type=int, default=16384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--flush-mb", type=int, default=256)
    ap.add_argument("--e2e-steps", type=int, default=None)
    return ap.parse_args()


# ----------------------------------------------------------------------------- CPU baseline (oracle)
def cpu_reference_run(wl, n_cpu: int, steps: int, warmup: int):
    """Times the float64 torch.func restatement of the reference's train step
    (grad(loss_fun) + Adam, software.py:387-393) on the host cores."""
    import torch

    from oracle import reference_oracle as O
    from pinn_based_online_pde_calculator_b200.workloads import init_params, make_points, unflatten

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net = wl.net
    frac = n_cpu / wl.n_col
    x_col, x_bd, u_bd = make_points(wl)
    x_col = x_col[:n_cpu]
    nb = [max(1, int(round(len(a) * frac))) for a in x_bd]
    x_bd = [a[:k] for a, k in zip(x_bd, nb)]
    u_bd = [a[:k] for a, k in zip(u_bd, nb)]
    params = [[torch.tensor(W, dtype=torch.float64), torch.tensor(b, dtype=torch.float64)]
              for W, b in unflatten(net, init_params(net))]
    limit = [torch.tensor(net.lb, dtype=torch.float64), torch.tensor(net.ub, dtype=torch.float64)]
    f_u = O.sol_pred_create(limit, net.scl, net.epsil, act_s=net.act_first, feature_map=net.feature_map,
                            hidden_act=("tanh", "sin")[net.act_hidden])
    names = {1: ("x",), 2: ("x", "y"), 3: ("x", "y", "t")}[net.d_in]
    if net.d_in == 2 and "u_t" in wl.expr and "u_y" not in wl.expr:
        names = ("x", "t")
    residual = None if net.feature_map == "polar" else O.make_gov_eqn_expr(wl.expr, names)
    lossf = O.loss_create(f_u, torch.tensor([wl.lw, 0.0], dtype=torch.float64), 1.0, residual=residual)
    data = dict(x_col=torch.tensor(x_col, dtype=torch.float64),
                cond_bd=[[torch.tensor(a, dtype=torch.float64) for a in x_bd],
                         [torch.tensor(a, dtype=torch.float64)[:, None] for a in u_bd]])
    lossf.ref = float(lossf(params, data)[1][0])
    st = O.AdamState(params)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        params, info, st = O.adam_minimizer(lossf, params, data, 1e-3, st)
        times.append(time.perf_counter() - t0)
    t = np.array(times[warmup:])
    return dict(n_cpu=n_cpu, n_bd=nb, cores=cores, total_s=float(t.sum()), median_s=float(np.median(t)),
                pts_per_s_total=float(n_cpu * len(t) / t.sum()), pts_per_s_median=float(n_cpu / np.median(t)))


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """SM clock / throttle-reason sampling DURING the timed region: NVML in a thread (2 ms period),
    falling back to `nvidia-smi -lms 100` when pynvml is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines, self.samples, self.stop_flag = gpu_index, None, [], [], False
        self.nvml = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.th.join(timeout=1)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": float(self.sm_max), "reasons": ["no samples"]}
            n = self.nvml
            bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            reasons = sorted(k for k, b in bits.items() if any(r & b for _, _, r in self.samples))
            sm = [s for s, _, _ in self.samples]
            return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(self.sm_max), "power_w_max": float(max(p for _, p, _ in self.samples)),
                    "samples": len(sm), "source": "nvml", "reasons": reasons}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [s.strip() for s in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); pw.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = [s for s in sm if s >= 0.5 * max(sm)]
        return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "source": "nvidia-smi", "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- main
def main():
    args = parse_args()
    from pinn_based_online_pde_calculator_b200.workloads import init_params, make_points, make_workload

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = make_workload(args.workload, args.n_col)
    config = {"workload": f"{wl.name}: {wl.description}", "kernel": os.environ.get("PINN_B200_KERNEL", "auto"), "n_col_per_gpu": wl.n_col, "n_bd_per_gpu": sum(wl.n_bd),
              "equation": wl.expr, "network": f"{wl.net.n_hidden}x{wl.net.width}", "parallelism": f"dp{world}",
              "l2": f"flushed between timed steps ({args.flush_mb} MB memset)", "optimizer": "adam lr=1e-3"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(wl, args.cpu_points, args.steps, args.warmup)
        sample = (f"oracle float64 torch.func nested-vjp train step on {r['n_cpu']} of {wl.n_col} collocation points "
                  f"+ {sum(r['n_bd'])} boundary points, all host threads")
        v = r["pts_per_s_total"]
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * r["total_s"] / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    import torch

    from pinn_based_online_pde_calculator_b200 import PinnEngine
    from pinn_based_online_pde_calculator_b200.engine import fma_peak_tflops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    eng = PinnEngine(wl.net, wl.eq, n_bc=len(wl.n_bd), device=local_rank)
    stream = torch.cuda.Stream(device=local_rank)
    eng.set_stream(stream.cuda_stream)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local_rank}")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(PinnEngine.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        eng.init_nccl(bytes(idt.cpu().numpy().tobytes()), rank, world)

    eng.set_params(init_params(wl.net))
    x_col, x_bd, u_bd = make_points(wl, rank=rank)
    # pinned host copies (e2e path) and device-resident copies (kernel-throughput path)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h_col, h_bd, h_ub = pin(x_col), [pin(a) for a in x_bd], [pin(a) for a in u_bd]
    dev = f"cuda:{local_rank}"
    d_col, d_bd, d_ub = h_col.to(dev), [a.to(dev) for a in h_bd], [a.to(dev) for a in h_ub]
    torch.cuda.synchronize()

    def set_counts():
        if world > 1:
            eng.set_global_counts(wl.n_col * world, [n * world for n in wl.n_bd])

    eng.set_points(d_col, d_bd, d_ub)
    set_counts()
    eng.set_loss(wl.lw, 1.0)
    _, info0 = eng.loss_grad(want_grad=False)
    eng.set_loss(wl.lw, float(info0[0]))  # lref = initial loss (software.py:739)
    eng.adam_init()
    lr = 1e-3
    flush = torch.empty(args.flush_mb << 20, dtype=torch.uint8, device=dev) if args.flush_mb > 0 else None

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- device-resident throughput ("value")
    for _ in range(max(3, args.warmup)):
        eng.adam_steps(1, lr, want_rows=False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    with torch.cuda.stream(stream):
        for i in range(args.steps):
            if flush is not None:
                flush.fill_(i & 0xFF)
            ev[i][0].record(stream)
            eng.adam_steps(1, lr, want_rows=False)
            ev[i][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = wl.n_col * world / (ms_per_step * 1e-3)
    _, info1 = eng.loss_grad(want_grad=False)

    # ---------------- end-to-end through the public API with HOST buffers ("e2e")
    e2e_steps = args.e2e_steps or args.steps
    np_col, np_bd, np_ub = h_col.numpy(), [a.numpy() for a in h_bd], [a.numpy() for a in h_ub]
    for _ in range(3):
        eng.set_points(np_col, np_bd, np_ub)
        set_counts()
        eng.adam_steps(1, lr, want_rows=True)
    # pipelined: the H2D copy of step i+1 (copy stream) overlaps the compute of step i (engine stream)
    eng.prefetch_points(np_col, np_bd, np_ub)
    eng.commit_points()
    eng.adam_steps(1, lr, want_rows=True)
    barrier()
    t0 = time.perf_counter()
    eng.prefetch_points(np_col, np_bd, np_ub)      # H2D of step 0's inputs from pinned host memory
    for i in range(e2e_steps):
        eng.commit_points()                         # swap the staged inputs in (engine stream, no host sync)
        if i + 1 < e2e_steps:
            eng.prefetch_points(np_col, np_bd, np_ub)   # H2D of the NEXT step's inputs, overlapped
        rows = eng.adam_steps(1, lr, want_rows=True)    # D2H of the step's loss_info (synchronises)
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = wl.n_col * world * e2e_steps / e2e_s
    h2d = 4 * (x_col.size + sum(a.size for a in x_bd) + sum(a.size for a in u_bd))
    d2h = 8 * eng.n_info

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (collocation jet kernel)
    eng.set_points(d_col, d_bd, d_ub)
    set_counts()
    col_ms, bc_ms = eng.time_kernels(reps=5, flush_bytes=args.flush_mb << 20)
    fl = wl.flops_per_point()
    flops_launch = fl["col"] * wl.n_col
    achieved = flops_launch / (col_ms * 1e-3) / 1e12
    fma_peak = fma_peak_tflops(local_rank, 0)
    fma2_peak = fma_peak_tflops(local_rank, 1)
    hmma_peak = fma_peak_tflops(local_rank, 9)   # mma.sync m16n8k8 TF32, register operands
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(wl.name)
    except Exception:
        pass
    alg_bytes = 4.0 * (wl.net.d_in + wl.eq.n_aux) * wl.n_col   # coordinates + hoisted per-point columns
    tensor = eng.kernel == "mma_3xtf32"
    # forward GEMM: 3 TF32 MMAs per product; data- and weight-gradient GEMMs: 1 TF32 + 2 bf16 MMAs of half the cost
    # (= 2 TF32 equivalents): 7 TF32-MMA equivalents per 3 algorithmic products
    peak = hmma_peak * 3.0 / 7.0 if tensor else fma_peak
    roofline = {
        "bound": "tensor" if tensor else "fp32",
        "kernel": ("jet_mma_kernel<train>" if tensor else "jet_mlp_kernel<train>") + " (collocation term)",
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "peak_source": ("measured in this run: mma.sync TF32 (HMMA.1688) rate x 3/7 -- split-precision products for the "
                        "1e-5 parity bar: the forward GEMM costs three TF32 MMAs per product, the two backward GEMMs one TF32 "
                        "+ two bf16 MMAs (half cost each); achieved counts ALGORITHMIC flops"
                        if tensor else
                        "FFMA microbenchmark measured in this run (pinn_fma_peak); the path is fp32-FMA bound, "
                        "not HBM-bound (SURVEY.md section 8d)"),
        "fp32_ffma_peak": fma_peak, "frac_of_fp32_ffma_peak": achieved / fma_peak,
        "hmma_tf32_peak": hmma_peak, "fma2_peak": fma2_peak,
        "algorithmic_flops_per_launch": flops_launch, "kernel_ms": col_ms, "bc_kernel_ms": bc_ms,
        "channels_algorithmic": fl["K"], "channels_executed": fl["K_exec"],
        "executed_flops_per_launch": fl["col_exec"] * wl.n_col,
        "executed_note": "Laplacian-type operators are propagated as ONE combined second-order channel, so the kernel "
                         "executes K_exec/K of the SURVEY's algorithmic MACs; frac uses the algorithmic count",
        "kernel_share_of_step": (col_ms + bc_ms) / ms_per_step,
        "traffic": traffic,
        "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (col_ms * 1e-3) / 1e9,
                "peak_gbs": peaks.get("hbm_gbs", 6650.0), "peak_source": "measured" if peaks else "fallback"},
        "frac_of_bf16_tensor_peak": achieved / peaks.get("bf16_tflops", 1590.0),
    }

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(wl, args.cpu_points, 3, 1)
        cpu = {"value": r["pts_per_s_median"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"oracle float64 torch.func nested-vjp train step, median of 3 after 1 warm-up, on {r['n_cpu']} "
                         f"of {wl.n_col} collocation + {sum(r['n_bd'])} boundary points"}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (split-precision tensor-core products: 3xTF32 forward, TF32 + 2 bf16 correction terms backward; fp32 accumulate)" if eng.kernel == "mma_3xtf32" else "f32",
        "data": "synthetic", "config": config,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps,
                "note": "host buffers through PinnEngine.prefetch_points/commit_points + adam_steps: the H2D copy of step i+1 runs on a copy stream under the compute of step i; loss_info read back every step"},
        "gpu_launches": 7 * args.steps,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "loss_first": float(info0[0]), "loss_last": float(info1[0]), "wall_s_timed_region": t_wall,
    }
    print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
