#!/bin/bash
# ncu --set full of the top extended-source kernels on config C4 (one capture each); -> gpurun_out/
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for K in k_limb_walk k_refine_solve k_contours; do
  ncu --set full --clock-control none --import-source on -k regex:${K} -s 1 -c 1 -o gpurun_out/prof_ext_${K} -f \
      python scripts/bench_configs.py --only C4 > gpurun_out/ncu_ext_${K}.log 2>&1
done
ls -la gpurun_out | tail -8
