#!/bin/bash
# launch lists (per-kernel device time) for the extended-source configs; outputs -> gpurun_out/
set -x
mkdir -p gpurun_out
TAG=${TAG:-r01}
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_c4_${TAG}.csv python scripts/bench_configs.py --only C4 > gpurun_out/ncu_c4_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/launches_c3_${TAG}.csv python scripts/bench_configs.py --only C3 > gpurun_out/ncu_c3_${TAG}.log 2>&1
ls -la gpurun_out
