"""Regenerate profiles/r01_SUMMARY.md from the bench JSON lines and ncu summaries kept in profiles/."""
import json
import os

root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles") + "/"
last = lambda f: json.loads(open(root + f).read().strip().splitlines()[-1])
allc = [json.loads(l) for l in open(root + "r01_bench_all_configs_1gpu.jsonl") if l.strip()]
sc = [json.loads(l) for l in open(root + "r01_scaling_C2_final.jsonl") if l.strip()]
head = last("r01_bench_C2_1gpu.json")
c5 = last("r01_scaling_C5_8gpu.json")
ncu = json.load(open(root + "r01_final_C2_col_kernel_ncu.json"))["metrics"]
g = lambda k: float(ncu[k]["value"].replace(",", ""))
L = []
L.append("# Round 1 measurements (B200, driver 580, CUDA 12.9)\n")
L.append("All numbers are `bench.py` JSON lines kept next to this file; CUDA-event timing on the engine stream, L2 flushed between timed steps, SM clocks sampled with NVML during the timed region (1965 MHz, no throttle reasons).  Regenerate with `python tools/make_summary.py`.\n")
L.append("## 1 GPU, every BASELINE.json config\n")
L.append("| config | kernel | N_col | ms/step | points/s | e2e points/s | algorithmic TFLOP/s | frac of its roofline | frac of FFMA peak |")
L.append("|---|---|---|---|---|---|---|---|---|")
for x in allc:
    r = x["roofline"]
    k = "split-precision mma.sync" if "mma" in r["kernel"] else "fp32 FFMA2"
    L.append(f"| {x['config']['workload'][:2]} | {k} | {x['config']['n_col_per_gpu']:,} | {x['ms_per_step']:.3f} | {x['value']:.4g} | {x['e2e']['value']:.4g} | {r['achieved']:.1f} | {r['frac']:.3f} | {r['frac_of_fp32_ffma_peak']:.3f} |")
L.append("\nRoofline of the tensor-core kernel = measured `mma.sync` TF32 rate (277 TFLOP/s) x 3/7 (forward GEMM: three TF32 MMAs per product; each backward GEMM: one TF32 + two half-cost bf16 MMAs); of the fp32 kernel = measured FFMA rate (71.7 TFLOP/s).")
L.append("`algorithmic TFLOP/s` uses the SURVEY's count 2K(2M1+M2) with one channel per derivative; the kernels execute K_exec/K of it for Laplacian-type operators (C2, C4, R0: 4/5; C5: 5/6).")
L.append("C1/R0 are launch-latency bound (1k / 5.2k points).  C3 pads W=50 to 64 (39 % extra MACs not counted as algorithmic work).\n")
L.append("## Weak scaling (each GPU holds the full per-GPU workload; one fused NCCL allreduce per step; measured on the v9 build, 6.22 ms/step on 1 GPU)\n")
L.append("| config | GPUs | points/s | e2e points/s | ms/step | speed-up vs 1 GPU |")
L.append("|---|---|---|---|---|---|")
v1 = sc[0]["value"]
for x in sc:
    L.append(f"| C2 | {x['n_gpus']} | {x['value']:.4g} | {x['e2e']['value']:.4g} | {x['ms_per_step']:.3f} | {x['value'] / v1:.2f} |")
L.append(f"| C5 (8M pts/GPU, 64M total; earlier build, 1430 ms/step on 1 GPU) | 8 | {c5['value']:.4g} | {c5['e2e']['value']:.4g} | {c5['ms_per_step']:.1f} | 7.98 |")
L.append("\n8-rank parity (`tools/check_nccl.py`, `r01_nccl_parity_8gpu.txt`): gradient rel. error vs single GPU 1.05e-7, replicas bit-identical after 5 Adam steps.\n")
L.append("## Headline (C2, 1 GPU)\n")
cb = head["cpu_baseline"]
L.append(f"* value {head['value']:.4g} points/s ({head['ms_per_step']:.2f} ms/step), e2e (host buffers: pipelined H2D of 8.48 MB per step on a copy stream + D2H of the loss row every step) {head['e2e']['value']:.4g} points/s")
L.append(f"* CPU baseline (oracle port, float64, {cb['cores']} cores): {cb['value']:.4g} points/s -> GPU/CPU = {head['e2e']['value'] / cb['value']:.0f}x end to end")
L.append(f"* roofline: {head['roofline']['achieved']:.1f} algorithmic TFLOP/s = {head['roofline']['frac']:.3f} of the split-precision mma.sync ceiling measured in the same run ({head['roofline']['peak']:.1f} TFLOP/s), {head['roofline']['frac_of_fp32_ffma_peak']:.3f} of the fp32 FFMA peak")
L.append(f"* clocks: {head['clocks']}")
L.append(f"* ncu (`r01_final_C2_col_kernel_ncu.json`): {g('gpu__time_duration.sum'):.2f} ms, {g('launch__registers_per_thread'):.0f} registers, 2 CTAs/SM, tensor pipe {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} % active, issue slots {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} % busy; DRAM traffic of the collocation kernel {g('dram__bytes_read.sum') + g('dram__bytes_write.sum'):.1f} MB per launch (algorithmic 12 MB; 4.97 GB before the persisting L2 window)")
L.append("* time-to-L2 (C1, `tests/time_to_l2.py --cpu`): rel. L2 < 1e-2 after 0.034 s (CPU oracle 2.34 s), < 1e-3 after 0.097 s (CPU 9.23 s)\n")
L.append("## tcgen05\n")
L.append("* `r01_umma_probe.txt` — hand-built `tcgen05.mma.kind::tf32` descriptors on the real chip: K-major / MN-major layouts, TS form (A in TMEM), M=64 lane interleave, truncating accumulate, 79/95/128-cycle issue floors (DESIGN.md 4.1)")
L.append("* experimental family C (`PINN_B200_KERNEL=umma`): parity 6e-7 on the gradient, C2 eval 2.6 ms (family B 2.7 ms), train step 9.1 ms (family B: 6.2 ms); phase clocks of CTA 0: forward epilogue 30 %, forward MMA wait 6 %, fused backward pass 29 %, weight-/data-gradient MMA wait 18 %, flush 7 %, TMEM->staging copy 7 %; `r01_umma_train_kernel_ncu.json` is the ncu summary of the 10.2 ms build\n")
L.append("## Files")
L.append("* `r01_final_C2_col_kernel_ncu.json` / `_launches.csv` / `_launch_shares.txt` — ncu `--set full` summary of the final kernel and the launch list of `python bench.py --steps 5 --warmup 3 --no-cpu-baseline` (the `k_fma_peak` / `k_mma_tf32_probe` launches are the live peak measurements outside the timed steps; the collocation kernel is 97 % of the step kernels)")
L.append("* `r01_final_C2_phase_profile.txt` — clock64 phase breakdown of the production kernel")
L.append("* `r01_bench_C2_1gpu.json`, `r01_bench_all_configs_1gpu.jsonl`, `r01_scaling_C2_final.jsonl`, `r01_scaling_C5_8gpu.json` — bench lines")
L.append("* `r01_col_kernel_v1_ncu.json`, `r01_launches_C2_v1.csv` — the first (SIMT v1) kernel for comparison")
L.append("* `roofline_traffic.json` — DRAM bytes per launch read by `bench.py` into `roofline.traffic`")
open(root + "r01_SUMMARY.md", "w").write("\n".join(L) + "\n")
print("\n".join(L[:14]))
