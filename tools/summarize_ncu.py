"""Summarise an .ncu-rep (one kernel) + a launch-list csv into profiles/ (json + csv)."""
import collections
import csv
import json
import subprocess
import sys

rep, launches, tag, kernel_desc = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum', 'smsp__inst_executed.sum']
keep += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
out = {h: {"unit": u, "value": v} for h, u, v in zip(hdr, units, vals) if h in keep}
json.dump({"kernel": kernel_desc, "capture": "ncu --set full --clock-control none --import-source on (one launch)", "metrics": out},
          open(f"profiles/{tag}_ncu.json", "w"), indent=1)
# launch list: strip ncu banner lines, aggregate shares
lines = [l for l in open(launches) if l.startswith('"')]
open(f"profiles/{tag}_launches.csv", "w").writelines(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
for r in csv.DictReader(lines):
    n = r["Kernel Name"].split("(")[0][:70]
    agg[n][0] += 1
    agg[n][1] += float(r["Metric Value"])
tot = sum(v[1] for v in agg.values())
with open(f"profiles/{tag}_launch_shares.txt", "w") as f:
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        line = f"{n:72s} launches {c:3d} total {t / 1e6:9.3f} ms share {100 * t / tot:5.1f}%"
        print(line)
        f.write(line + "\n")
d = dict(zip(hdr, vals))
print({k: d.get(k) for k in ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
                             'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active']})
