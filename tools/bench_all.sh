#!/bin/bash
# Throughput of every BASELINE.json config on one GPU (JSON lines -> gpurun_out/bench_all.jsonl)
cd "$(dirname "$0")/.."
out=gpurun_out/bench_all.jsonl
: > $out
python bench.py --workload C1 --steps 50 --warmup 10 --no-cpu-baseline 2>>gpurun_out/bench_all.err | tail -1 >> $out
python bench.py --workload R0 --steps 50 --warmup 10 --no-cpu-baseline 2>>gpurun_out/bench_all.err | tail -1 >> $out
python bench.py --workload C2 --steps 20 --warmup 5 --no-cpu-baseline 2>>gpurun_out/bench_all.err | tail -1 >> $out
PINN_B200_KERNEL=simt python bench.py --workload C2 --steps 20 --warmup 5 --no-cpu-baseline 2>>gpurun_out/bench_all.err | tail -1 >> $out
python bench.py --workload C3 --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_all.err | tail -1 >> $out
PINN_B200_KERNEL=simt python bench.py --workload C3 --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_all.err | tail -1 >> $out
python bench.py --workload C4 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>>gpurun_out/bench_all.err | tail -1 >> $out
PINN_B200_KERNEL=simt python bench.py --workload C4 --n-col 4000000 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>>gpurun_out/bench_all.err | tail -1 >> $out
python bench.py --workload C5 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>>gpurun_out/bench_all.err | tail -1 >> $out
PINN_B200_KERNEL=mma python bench.py --workload C5 --n-col 2000000 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>>gpurun_out/bench_all.err | tail -1 >> $out
python - <<'PY'
import json
for l in open("gpurun_out/bench_all.jsonl"):
    try: d = json.loads(l)
    except Exception: print("BAD", l[:200]); continue
    r = d["roofline"]
    print(f'{d["config"]["workload"][:2]} kernel={r["kernel"][:14]:14s} n={d["config"]["n_col_per_gpu"]:>9d} ms/step={d["ms_per_step"]:10.3f} pts/s={d["value"]:.4g} e2e={d["e2e"]["value"]:.4g} '
          f'alg_TF={r["achieved"]:.1f} frac={r["frac"]:.3f} frac_ffma={r["frac_of_fp32_ffma_peak"]:.3f} loss {d["loss_first"]:.3g}->{d["loss_last"]:.3g}')
PY
