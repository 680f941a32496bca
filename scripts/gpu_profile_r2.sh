#!/bin/bash
# Round-2 ncu evidence (run under gpurun, one GPU).  Outputs -> gpurun_out/ (summaries are copied to profiles/ here).
#  1. launch list of the bench command itself (cold-cache, serialised: compare SHARES, not absolute times)
#  2. one `ncu --set full` capture per kernel VERDICT r1 asked for
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv \
    --log-file gpurun_out/launches_${TAG}_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
cap() {  # cap <name> <demangled-name regex> <target> [skip]
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s ${4:-1} -c 1 \
      -o gpurun_out/prof_${TAG}_$1 -f python scripts/profile_targets.py $3 > gpurun_out/ncu_${TAG}_$1.log 2>&1
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_$1.csv 2>/dev/null
}
cap ea_comp10 'ea_kernel<10, true' comp10
cap ea_deg5 'ea_kernel<5, false' deg5
cap ps_map 'ps_kernel<2, false, 0' map
cap k_tracks 'k_tracks<10' c4 0
cap k_refine_select 'k_refine_select<10' c4 5
cap k_contours 'k_contours<10, false' c4 0
cap k_contours_grad 'k_contours<10, true' c4grad 0
cap k_limb_walk 'k_limb_walk<3' c4 0
cap k_ld_pq 'k_ld_pq<2' c3 1
ls -la gpurun_out | tail -30
