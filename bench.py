#!/usr/bin/env python
"""Benchmark of the PINN residual-loss + gradient hot path (BASELINE.json metric:
collocation points/sec per train step).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C2]

One "step" = one Adam train step (software.py:412-414): residual loss over ALL
collocation + boundary points, its gradient, and the Adam update.  N>1 is launched
by torchrun, one rank per GPU, weak scaling (each rank holds the workload's point
count; one fused NCCL allreduce per step).  Rank 0 prints ONE JSON line: the C2 headline
(value, e2e, roofline, cpu_baseline, clocks) plus `configs` (the other BASELINE.json
configs C1, C3, C4, C5 at their named sizes), `time_to_l2` (C1, Adam -> L-BFGS, GPU and
CPU oracle; N=1) and, for N>1, `strong` (C2 1M and C5 8M points in TOTAL, sharded).
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
    ap.add_argument("--cpu-points", type=int, default=131072)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--flush-mb", type=int, default=256)
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-extra", action="store_true", help="headline workload only: skip the other BASELINE configs, time-to-L2 and strong scaling")
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


# ----------------------------------------------------------------------------- helpers
def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def kept_profile(name):
    """numbers kept from the committed ncu captures (profiles/roofline_traffic.json), per workload"""
    try:
        v = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(name)
        return v if isinstance(v, dict) else ({"dram_bytes_per_launch": v} if v is not None else None)
    except Exception:
        return None


def roofline_of(wl, eng, col_ms, bc_ms, ms_per_step, extra_peaks=None):
    """Roofline of the dominant kernel (collocation jet kernel).  The path is a dense contraction, so the bound
    is the tensor pipe.  The 1e-5 parity bar forces split-precision products (3 MMAs of TF32 cost per
    algorithmic product: SURVEY.md section 7), so the ceiling the fraction is quoted against is the MEASURED
    dense bf16 rate / 6 (= TF32 rate / 3)."""
    peaks, src = load_peaks()
    fl = wl.flops_per_point()
    flops_launch = fl["col"] * wl.n_col
    achieved = flops_launch / (col_ms * 1e-3) / 1e12
    peak = peaks["bf16_tflops"] / 6.0
    kept = kept_profile(wl.name) or {}
    alg_bytes = 4.0 * (wl.net.d_in + wl.eq.n_aux) * wl.n_col
    r = {
        "bound": "tensor",
        "kernel": {"tc_bf16x3": "jet_tc_kernel<train> (tcgen05 kind::f16, bf16x3 split)", "mma_3xtf32": "jet_mma_kernel<train> (mma.sync 3xTF32)",
                   "simt_fp32": "jet_mlp_kernel<train> (fp32 FFMA2)", "umma_3xtf32": "jet_umma_train_kernel"}[eng.kernel] + " (collocation term)",
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "peak_source": f"bf16_tflops / 6 = {peaks['bf16_tflops']:.1f} / 6, {src}: dense bf16 cuBLAS burst rate divided by the six bf16 "
                       "(= three TF32) MMAs one fp32-accurate product costs at the 1e-5 parity bar",
        "achieved_counts": "ALGORITHMIC flops 2K(2M1+M2) per point with K = one channel per derivative (SURVEY.md 8d); the kernel "
                           "executes K_exec channels (one combined second-order channel for Laplacian-type operators)",
        "channels_algorithmic": fl["K"], "channels_executed": fl["K_exec"],
        "achieved_executed": fl["col_exec"] * wl.n_col / (col_ms * 1e-3) / 1e12,
        "frac_of_bf16_dense_peak": achieved / peaks["bf16_tflops"],
        "algorithmic_flops_per_launch": flops_launch, "kernel_ms": col_ms, "bc_kernel_ms": bc_ms,
        "kernel_share_of_step": (col_ms + bc_ms) / ms_per_step if ms_per_step else None,
        "tensor_pipe_active_pct": kept.get("tensor_pipe_active_pct"), "tensor_pipe_source": kept.get("source"),
        "traffic": kept.get("dram_bytes_per_launch"),
        "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (col_ms * 1e-3) / 1e9,
                "peak_gbs": peaks.get("hbm_gbs", 6650.0)},
    }
    if extra_peaks:
        r.update({"fp32_ffma_peak": extra_peaks["ffma"], "frac_of_fp32_ffma_peak": achieved / extra_peaks["ffma"],
                  "hmma_tf32_peak": extra_peaks["hmma"], "frac_of_hmma_3xtf32_ceiling": achieved / (extra_peaks["hmma"] * 3.0 / 7.0)})
    return r


class Runner:
    """One workload on this rank's GPU: engine, device-resident points, timed Adam steps."""

    def __init__(self, wl, args, rank, world, local_rank, dist, n_col_total=None):
        import torch

        from pinn_based_online_pde_calculator_b200 import PinnEngine
        from pinn_based_online_pde_calculator_b200.workloads import init_params, make_points

        self.wl, self.args, self.rank, self.world, self.dist, self.torch = wl, args, rank, world, dist, torch
        self.dev = f"cuda:{local_rank}"
        self.eng = eng = PinnEngine(wl.net, wl.eq, n_bc=len(wl.n_bd), device=local_rank)
        self.stream = torch.cuda.Stream(device=local_rank)
        eng.set_stream(self.stream.cuda_stream)
        if world > 1:
            idt = torch.zeros(128, dtype=torch.uint8, device=self.dev)
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(PinnEngine.nccl_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            eng.init_nccl(bytes(idt.cpu().numpy().tobytes()), rank, world)
        eng.set_params(init_params(wl.net))
        self.x_col, self.x_bd, self.u_bd = make_points(wl, rank=rank)
        pin = lambda a: torch.from_numpy(a).pin_memory()
        self.h_col, self.h_bd, self.h_ub = pin(self.x_col), [pin(a) for a in self.x_bd], [pin(a) for a in self.u_bd]
        self.d_col, self.d_bd, self.d_ub = self.h_col.to(self.dev), [a.to(self.dev) for a in self.h_bd], [a.to(self.dev) for a in self.h_ub]
        torch.cuda.synchronize()
        self.set_device_points()
        eng.set_loss(wl.lw, 1.0)
        _, self.info0 = eng.loss_grad(want_grad=False)
        eng.set_loss(wl.lw, float(self.info0[0]))  # lref = initial loss (software.py:739)
        eng.adam_init()
        self.flush = torch.empty(args.flush_mb << 20, dtype=torch.uint8, device=self.dev) if args.flush_mb > 0 else None

    def set_counts(self):
        if self.world > 1:
            self.eng.set_global_counts(self.wl.n_col * self.world, [n * self.world for n in self.wl.n_bd])

    def set_device_points(self):
        self.eng.set_points(self.d_col, self.d_bd, self.d_ub)
        self.set_counts()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(self, steps, warmup, lr=1e-3, sampler=None):
        """W warm-up steps, then exactly `steps` steps, each bracketed by CUDA events on the engine stream, an L2
        flush between them; barrier + synchronize on both sides; max over ranks.  Returns ms per step."""
        torch, eng = self.torch, self.eng
        for _ in range(max(3, warmup)):
            eng.adam_steps(1, lr, want_rows=False)
        self.barrier()
        if sampler is not None:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(self.stream):
            for i in range(steps):
                if self.flush is not None:
                    self.flush.fill_(i & 0xFF)
                ev[i][0].record(self.stream)
                eng.adam_steps(1, lr, want_rows=False)
                ev[i][1].record(self.stream)
        self.barrier()
        self.wall = time.perf_counter() - t0
        total_ms = self.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
        return total_ms / steps

    def close(self):
        self.eng.close()
        del self.d_col, self.d_bd, self.d_ub, self.flush
        self.torch.cuda.empty_cache()


def time_to_l2_c1(local_rank, cpu: bool, n_adam=2000, n_lbfgs=300, thresholds=(1e-2, 1e-3)):
    """BASELINE metric, second half: wall time of the reference's Adam -> L-BFGS schedule on C1 (1D Poisson,
    u* = x(1-x)) until the relative L2 error on the 111-point test grid first drops below each threshold; the
    CPU oracle (float64, all host cores) runs the identical Adam schedule next to it."""
    from pinn_based_online_pde_calculator_b200 import PinnEngine
    from pinn_based_online_pde_calculator_b200.workloads import init_params, make_points, make_workload, unflatten

    wl = make_workload("C1")
    x_col, x_bd, u_bd = make_points(wl)
    grid = np.linspace(0, 1, 111, dtype=np.float32)[:, None]
    exact = grid[:, 0] * (1 - grid[:, 0])
    eng = PinnEngine(wl.net, wl.eq, n_bc=len(x_bd), device=local_rank)
    eng.set_params(init_params(wl.net))
    eng.set_points(x_col, x_bd, u_bd)
    eng.set_loss(wl.lw, 1.0)
    eng.set_loss(wl.lw, float(eng.loss_grad(want_grad=False)[1][0]))
    l2 = lambda: float(np.linalg.norm(eng.eval(grid)[0] - exact) / np.linalg.norm(exact))
    eng.adam_init()
    eng.adam_steps(1, 1e-3, want_rows=False)   # graph capture outside the clock, then restart from the initial point
    eng.set_params(init_params(wl.net))
    eng.adam_init()
    hits, hist = {}, []
    t0 = time.perf_counter()

    def check(phase, k):
        e = l2()
        t = time.perf_counter() - t0
        hist.append((round(t, 4), phase, k, e))
        for thr in thresholds:
            if e < thr and thr not in hits:
                hits[thr] = {"gpu_s": t, "phase": phase, "step": k}

    for k in range(0, n_adam, 50):
        eng.adam_steps(50, 1e-3, want_rows=False)
        check("adam", k + 50)
        if len(hits) == len(thresholds):
            break
    res = None
    if len(hits) < len(thresholds):
        for k in range(0, n_lbfgs, 25):
            res, _ = eng.lbfgs(25, 1e-10)
            check("lbfgs", k + 25)
            if len(hits) == len(thresholds) or res["failed"] or res["converged"]:
                break
    out = {"workload": "C1: 1D Poisson, 3x20 tanh, 1k points; Adam lr=1e-3 (L2 checked every 50 steps) then L-BFGS (every 25 iterations)",
           "final_rel_l2": hist[-1][3], "total_gpu_s": hist[-1][0],
           "thresholds": {f"{thr:g}": hits.get(thr) for thr in thresholds}}
    eng.close()
    if cpu:
        import torch

        from oracle import reference_oracle as O

        torch.set_num_threads(os.cpu_count() or 1)
        net = wl.net
        params = [[torch.tensor(W, dtype=torch.float64), torch.tensor(b, dtype=torch.float64)] for W, b in unflatten(net, init_params(net))]
        limit = [torch.tensor(net.lb, dtype=torch.float64), torch.tensor(net.ub, dtype=torch.float64)]
        f_u = O.sol_pred_create(limit, net.scl, net.epsil, act_s=net.act_first, feature_map=net.feature_map)
        lossf = O.loss_create(f_u, torch.tensor([wl.lw, 0.0], dtype=torch.float64), 1.0, residual=O.make_gov_eqn_expr(wl.expr, ("x",)))
        data = dict(x_col=torch.tensor(x_col, dtype=torch.float64), cond_bd=[[torch.tensor(a, dtype=torch.float64) for a in x_bd],
                                                                            [torch.tensor(a, dtype=torch.float64)[:, None] for a in u_bd]])
        lossf.ref = float(lossf(params, data)[1][0])
        st = O.AdamState(params)
        g = torch.tensor(grid, dtype=torch.float64)
        chits = {}
        t0 = time.perf_counter()
        for k in range(n_adam):
            params, info, st = O.adam_minimizer(lossf, params, data, 1e-3, st)
            if (k + 1) % 50 == 0:
                e = float(np.linalg.norm(f_u(params, g).numpy()[:, 0] - exact) / np.linalg.norm(exact))
                for thr in thresholds:
                    if e < thr and thr not in chits:
                        chits[thr] = {"cpu_s": time.perf_counter() - t0, "step": k + 1}
                if len(chits) == len(thresholds) or time.perf_counter() - t0 > 40.0:
                    break
        out["cpu"] = {"cores": os.cpu_count(), "dtype": "f64", "kind": "port", "adam_steps_run": k + 1,
                      "thresholds": {f"{thr:g}": chits.get(thr) for thr in thresholds}}
        for thr in thresholds:
            a, b = hits.get(thr), chits.get(thr)
            if a and b:
                out["thresholds"][f"{thr:g}"]["cpu_s"] = b["cpu_s"]
                out["thresholds"][f"{thr:g}"]["speedup"] = b["cpu_s"] / a["gpu_s"]
    return out


def reference_problem_r0(local_rank, cpu: bool, gpu_steps=4000, cpu_budget_s=12.0):
    """The reference's OWN problem size (pinn_app/software.py:1142-1201: 5,200 collocation + 200 boundary points, 6x60
    network with the polar feature map): steady-state time of one Adam step through the CUDA-graph replay, an L-BFGS leg
    (per objective evaluation, device-resident loop) and the CPU oracle (float64, all host cores) on the identical step."""
    from pinn_based_online_pde_calculator_b200 import PinnEngine
    from pinn_based_online_pde_calculator_b200.workloads import init_params, make_points, make_workload, unflatten

    wl = make_workload("R0")
    x_col, x_bd, u_bd = make_points(wl)
    eng = PinnEngine(wl.net, wl.eq, n_bc=len(x_bd), device=local_rank)
    eng.set_params(init_params(wl.net))
    eng.set_points(x_col, x_bd, u_bd)
    eng.set_loss(wl.lw, 1.0)
    eng.set_loss(wl.lw, float(eng.loss_grad(want_grad=False)[1][0]))
    eng.adam_init()
    eng.adam_steps(100, 1e-3, want_rows=False)
    t0 = time.perf_counter()
    rows = eng.adam_steps(gpu_steps, 1e-3)          # rows come back: the wall clock includes the final synchronisation
    adam_wall = time.perf_counter() - t0
    adam_dev_ms = eng.last_ms()
    t0 = time.perf_counter()
    res, ev = eng.lbfgs(200, 1e-12)
    lb_wall = time.perf_counter() - t0
    out = {"workload": "R0: the reference's own smoke problem, 5,200 + 200 points, 6x60 tanh, polar feature map (software.py:1142-1201)",
           "kernel": eng.kernel, "adam": {"steps": gpu_steps, "us_per_step_device": 1e3 * adam_dev_ms / gpu_steps,
                                          "us_per_step_wall": 1e6 * adam_wall / gpu_steps, "launches_per_step": eng.launches_per_adam_step(),
                                          "loss_first": float(rows[0, 0]), "loss_last": float(rows[-1, 0])},
           "lbfgs": {"iterations": int(res["iterations"]), "evaluations": int(res["evaluations"]),
                     "us_per_evaluation_wall": 1e6 * lb_wall / max(1, int(res["evaluations"])), "host_syncs": eng.lbfgs_host_syncs(),
                     "final_loss": float(res["final_loss"])}}
    eng.close()
    if cpu:
        import torch

        from oracle import reference_oracle as O

        torch.set_num_threads(os.cpu_count() or 1)
        net = wl.net
        params = [[torch.tensor(W, dtype=torch.float64), torch.tensor(b, dtype=torch.float64)] for W, b in unflatten(net, init_params(net))]
        limit = [torch.tensor(net.lb, dtype=torch.float64), torch.tensor(net.ub, dtype=torch.float64)]
        f_u = O.sol_pred_create(limit, net.scl, net.epsil, act_s=net.act_first, feature_map=net.feature_map)
        lossf = O.loss_create(f_u, torch.tensor([wl.lw, 0.0], dtype=torch.float64), 1.0)   # residual: the reference's gov_eqn (polar Laplacian)
        data = dict(x_col=torch.tensor(x_col, dtype=torch.float64), cond_bd=[[torch.tensor(a, dtype=torch.float64) for a in x_bd],
                                                                            [torch.tensor(a, dtype=torch.float64)[:, None] for a in u_bd]])
        lossf.ref = float(lossf(params, data)[1][0])
        st = O.AdamState(params)
        params, _, st = O.adam_minimizer(lossf, params, data, 1e-3, st)   # warm-up
        n, t0 = 0, time.perf_counter()
        while n < 200 and time.perf_counter() - t0 < cpu_budget_s:
            params, info, st = O.adam_minimizer(lossf, params, data, 1e-3, st)
            n += 1
        cpu_ms = 1e3 * (time.perf_counter() - t0) / max(1, n)
        out["cpu"] = {"cores": os.cpu_count(), "dtype": "f64", "kind": "port", "adam_steps_run": n, "ms_per_step": cpu_ms,
                      "speedup_adam_step": cpu_ms * 1e3 / out["adam"]["us_per_step_wall"]}
    return out


# ----------------------------------------------------------------------------- main
def main():
    args = parse_args()
    from pinn_based_online_pde_calculator_b200.workloads import make_workload

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = make_workload(args.workload, args.n_col)

    def config_of(w, mode="weak"):
        return {"workload": f"{w.name}: {w.description}", "kernel": os.environ.get("PINN_B200_KERNEL", "auto"),
                "n_col_per_gpu": w.n_col, "n_bd_per_gpu": sum(w.n_bd), "equation": w.expr,
                "network": f"{w.net.n_hidden}x{w.net.width}", "parallelism": f"dp{world}",
                "l2": f"flushed between timed steps ({args.flush_mb} MB memset)", "optimizer": "adam lr=1e-3"}

    config = config_of(wl)

    if args.impl == "reference":
        if rank != 0:
            return
        n_cpu = min(args.cpu_points, wl.n_col)
        r = cpu_reference_run(wl, n_cpu, args.steps, args.warmup)
        sample = (f"oracle float64 torch.func nested-vjp train step on {r['n_cpu']} of {wl.n_col} collocation points "
                  f"+ {sum(r['n_bd'])} boundary points per step, all host threads")
        v = r["pts_per_s_total"]
        config["cpu_sample_points"] = r["n_cpu"]
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * r["total_s"] / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    import torch

    from pinn_based_online_pde_calculator_b200.engine import fma_peak_tflops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    R = Runner(wl, args, rank, world, local_rank, dist)
    eng, lr = R.eng, 1e-3

    # ---------------- device-resident throughput ("value")
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_per_step = R.timed_steps(args.steps, args.warmup, lr, sampler)
    t_wall = R.wall
    clocks = sampler.stop() if rank == 0 else None
    value = wl.n_col * world / (ms_per_step * 1e-3)
    _, info1 = eng.loss_grad(want_grad=False)

    # ---------------- end-to-end through the public API with HOST buffers ("e2e")
    e2e_steps = args.e2e_steps or args.steps
    np_col, np_bd, np_ub = R.h_col.numpy(), [a.numpy() for a in R.h_bd], [a.numpy() for a in R.h_ub]
    for _ in range(3):
        eng.set_points(np_col, np_bd, np_ub)
        R.set_counts()
        eng.adam_steps(1, lr, want_rows=True)
    # pipelined: the H2D copy of step i+1 (copy stream) overlaps the compute of step i (engine stream)
    eng.prefetch_points(np_col, np_bd, np_ub)
    eng.commit_points()
    eng.adam_steps(1, lr, want_rows=True)
    R.barrier()
    t0 = time.perf_counter()
    eng.prefetch_points(np_col, np_bd, np_ub)      # H2D of step 0's inputs from pinned host memory
    for i in range(e2e_steps):
        eng.commit_points()                         # swap the staged inputs in (engine stream, no host sync)
        if i + 1 < e2e_steps:
            eng.prefetch_points(np_col, np_bd, np_ub)   # H2D of the NEXT step's inputs, overlapped
        rows = eng.adam_steps(1, lr, want_rows=True)    # D2H of the step's loss_info (synchronises)
    R.barrier()
    e2e_s = R.max_over_ranks(time.perf_counter() - t0)
    e2e_value = wl.n_col * world * e2e_steps / e2e_s
    h2d = 4 * (R.x_col.size + sum(a.size for a in R.x_bd) + sum(a.size for a in R.u_bd))
    d2h = 8 * eng.n_info
    launches_per_step = eng.launches_per_adam_step()

    # ---------------- roofline of the dominant kernel (collocation jet kernel), rank 0
    roofline = None
    if rank == 0:
        R.set_device_points()
        col_ms, bc_ms = eng.time_kernels(reps=5, flush_bytes=args.flush_mb << 20)
        extra = {"ffma": fma_peak_tflops(local_rank, 0), "hmma": fma_peak_tflops(local_rank, 9)}
        roofline = roofline_of(wl, eng, col_ms, bc_ms, ms_per_step, extra)
    kernel_name = eng.kernel
    R.close()

    # ---------------- the other BASELINE.json configs at their named sizes (weak: every rank holds the full size)
    configs = []
    if not args.no_extra:
        for name, steps, warm in (("C1", 50, 5), ("C3", 5, 3), ("C4", 3, 3), ("C5", 2, 3)):
            if name == wl.name:
                continue
            w = make_workload(name)
            Rx = Runner(w, args, rank, world, local_rank, dist)
            ms = Rx.timed_steps(steps, warm, lr)
            rec = {"workload": f"{w.name}: {w.description}", "network": f"{w.net.n_hidden}x{w.net.width}", "kernel": Rx.eng.kernel,
                   "n_col_per_gpu": w.n_col, "n_gpus": world, "steps": steps, "warmup": max(3, warm), "ms_per_step": ms,
                   "value": w.n_col * world / (ms * 1e-3), "unit": UNIT}
            if rank == 0:
                Rx.set_device_points()
                cm, bm = Rx.eng.time_kernels(reps=2 if name in ("C4", "C5") else 5, flush_bytes=args.flush_mb << 20)
                rr = roofline_of(w, Rx.eng, cm, bm, ms)
                rec["roofline"] = {k: rr[k] for k in ("kernel", "achieved", "peak", "unit", "frac", "achieved_executed", "kernel_ms",
                                                      "kernel_share_of_step", "tensor_pipe_active_pct", "traffic")}
            configs.append(rec)
            Rx.close()

    # ---------------- strong scaling (N > 1): fixed TOTAL point counts sharded over the ranks
    strong = None
    if world > 1 and not args.no_extra:
        from pinn_based_online_pde_calculator_b200.engine import shard_range

        strong = {}
        for name, total, steps, warm in (("C2", 1_000_000, 20, 5), ("C5", 8_000_000, 3, 3)):
            b, e = shard_range(total, rank, world)
            w = make_workload(name, e - b)
            w.n_bd = [max(1, n // world) for n in w.n_bd]
            Rx = Runner(w, args, rank, world, local_rank, dist)
            Rx.eng.set_global_counts(total, [n * world for n in w.n_bd])
            Rx.set_counts = lambda Rx=Rx, total=total, w=w: Rx.eng.set_global_counts(total, [n * world for n in w.n_bd])
            ms = Rx.timed_steps(steps, warm, lr)
            strong[f"{name}_{total // 1_000_000}M_total"] = {"n_col_total": total, "n_col_per_gpu": e - b, "n_gpus": world, "ms_per_step": ms,
                                                            "value": total / (ms * 1e-3), "unit": UNIT, "steps": steps,
                                                            "kernel": Rx.eng.kernel}
            Rx.close()

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    ttl2 = None
    ref_problem = None
    if world == 1 and not args.no_extra:
        ttl2 = time_to_l2_c1(local_rank, cpu=not args.no_cpu_baseline)
        ref_problem = reference_problem_r0(local_rank, cpu=not args.no_cpu_baseline)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = min(args.cpu_points, wl.n_col)
        r = cpu_reference_run(wl, n_cpu, 3, 1)
        cpu = {"value": r["pts_per_s_median"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"oracle float64 torch.func nested-vjp train step, median of 3 after 1 warm-up, on {r['n_cpu']} "
                         f"of {wl.n_col} collocation + {sum(r['n_bd'])} boundary points"}
        if n_cpu > 16384:  # flatness in N: the same step on a 16,384-point sample
            r2 = cpu_reference_run(wl, 16384, 3, 1)
            cpu["flatness"] = {"n_16384": r2["pts_per_s_median"], f"n_{n_cpu}": r["pts_per_s_median"]}
        config["cpu_sample_points"] = n_cpu

    dtypes = {"tc_bf16x3": "f32 (tcgen05 bf16x3 split products, 6 bf16 MMAs per product, fp32 accumulate in TMEM)",
              "mma_3xtf32": "f32 (split-precision tensor-core products: 3xTF32 forward, TF32 + 2 bf16 correction terms backward; fp32 accumulate)",
              "simt_fp32": "f32", "umma_3xtf32": "f32 (3xTF32 tcgen05)"}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtypes[kernel_name], "data": "synthetic", "config": config,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps,
                "note": "host buffers through PinnEngine.prefetch_points/commit_points + adam_steps: the H2D copy of step i+1 runs on a copy stream under the compute of step i; loss_info read back every step"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "configs": configs, "time_to_l2": ttl2, "reference_problem": ref_problem, "strong": strong,
        "loss_first": float(R.info0[0]), "loss_last": float(info1[0]), "wall_s_timed_region": t_wall,
    }
    print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
