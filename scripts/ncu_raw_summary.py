#!/usr/bin/env python
"""Compact summary of `ncu -i x.ncu-rep --page raw --csv` dumps (one kernel per file):
    python scripts/ncu_raw_summary.py gpurun_out/r02_ncu_*.csv > profiles/r02_ncu_summary.txt"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_static", "static smem / CTA"), ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
    ("launch__occupancy_limit_registers", "CTAs/SM allowed by registers"), ("launch__occupancy_limit_shared_mem", "CTAs/SM allowed by smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads / instruction"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"), ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instructions"),
]
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    if len(rows) < 3:
        print(f"== {f}: empty"); continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"== {f}\n   {vals[col['Kernel Name']][:110]}")
    for k, label in KEYS:
        if k in col:
            print(f"   {label:32s} {vals[col[k]]} {units[col[k]]}")
    st = [(float(vals[i].replace(",", "") or 0), h) for i, h in enumerate(hdr)
          if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    st.sort(reverse=True)
    print("   top stalls (warps per issue):", ", ".join(
        f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in st[:5]))
