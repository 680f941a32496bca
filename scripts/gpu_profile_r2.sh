#!/bin/bash
# Round-2 ncu evidence (run under gpurun, one GPU).  Outputs -> gpurun_out/ (summaries are copied to profiles/ here).
#   bash scripts/gpu_profile_r2.sh [bench] [solver] [ext]
#  bench   launch list of the bench command itself (cold-cache, serialised: compare SHARES, not absolute times)
#  solver  one `ncu --set full` capture of ea_kernel<10,true>, ea_kernel<5,false>, ps_kernel<2,false,0>
#  ext     the same for the kernels of family 3 on C4 / C3
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02}
WHAT=${*:-bench solver ext}
cap() {  # cap <name> <function base name> <target> [skip]
  ncu --set full --clock-control none --import-source on --kernel-name-base function -k "$2" -s ${4:-1} -c 1 \
      -o gpurun_out/prof_${TAG}_$1 -f python scripts/profile_targets.py $3 > gpurun_out/ncu_${TAG}_$1.log 2>&1
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_$1.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page details --csv > gpurun_out/${TAG}_ncu_$1_details.csv 2>/dev/null
  # gpurun_out/ travels back only below 64 MiB: keep the CSV pages, drop the 25 MB report (KEEP_REP=1 keeps it)
  [ -n "$KEEP_REP" ] || rm -f gpurun_out/prof_${TAG}_$1.ncu-rep
}
for w in $WHAT; do
  case $w in
    bench)
      ncu --metrics gpu__time_duration.sum --clock-control none -c ${NLAUNCH:-8000} --csv \
          --log-file gpurun_out/launches_${TAG}_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline ${BENCH_ONLY:+--only $BENCH_ONLY} > gpurun_out/ncu_bench_${TAG}.log 2>&1 ;;
    solver)
      cap ea_comp10 ea_kernel comp10
      cap ea_deg5 ea_kernel deg5
      cap ps_map ps_kernel map ;;
    ext)
      cap k_limb_walk k_limb_walk c4 0
      cap k_round_select k_round_select c4 1
      cap k_round_solve k_round_solve c4 1
      cap k_sweep k_sweep c4 0
      cap k_open_compact k_open_compact c4 0
      cap k_contours_grad k_contours c4grad 0
      cap k_tracks k_tracks c4grad 0
      cap k_ld_pq k_ld_pq c3 1 ;;
  esac
done
ls -la gpurun_out | tail -30
