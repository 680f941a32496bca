#!/bin/bash
# ncu launch list (gpu__time_duration + instruction count + FP64 pipe) of a command, aggregated per (kernel, grid).
#   TAG=x bash scripts/launch_list.sh python bench.py --only C3x100 --steps 2 --warmup 3 --no-cpu-baseline
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -c ${NLAUNCH:-4000} --csv --log-file gpurun_out/${TAG:-ll}_launches.csv "$@" > gpurun_out/${TAG:-ll}_ncu.log 2>&1
python - gpurun_out/${TAG:-ll}_launches.csv <<'PY'
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ik = hdr.index("Kernel Name"); im = hdr.index("Metric Name"); iv = hdr.index("Metric Value"); ii = hdr.index("ID"); ig = hdr.index("Grid Size")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[ii], {"k": (r[ik][:52], r[ig])})[r[im]] = float(r[iv].replace(",", ""))
agg = collections.OrderedDict()
for a in per.values():
    g = agg.setdefault(a["k"], [0, 0.0, 0.0, 0.0])
    g[0] += 1; g[1] += a["gpu__time_duration.sum"]; g[2] += a["smsp__inst_executed.sum"]
    g[3] += a["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
tot = sum(g[1] for g in agg.values())
for (k, grid), g in agg.items():
    if g[1] / tot > 0.002:
        print(f"{k:52s} {grid:>15s} n={g[0]:4d} mean {g[1]/g[0]/1e6:8.3f} ms  share {g[1]/tot:5.3f}  {g[2]/g[0]/1e9:6.3f} Ginst  fp64 {g[3]/g[0]:5.1f} %")
PY
