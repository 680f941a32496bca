#!/bin/bash
# ncu evidence for the dominant kernel (run under gpurun, one GPU).  Outputs -> gpurun_out/
set -x
mkdir -p gpurun_out
TAG=${TAG:-r01}
KREGEX=${KREGEX:-ea_kernel}
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
# launch list of the bench command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
# full capture of the top kernel
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 3 -c 1 \
    -o gpurun_out/prof_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out
