#!/bin/bash
# ncu --set full capture (raw page + source page with CUDA/SASS correlation) of family-3 kernels on C4.
#   LIB=build_variants/x.so KERNELS="k_walk_refine k_open_staged" TAG=dev4 bash scripts/ext_prof.sh
mkdir -p gpurun_out
TAG=${TAG:-dev}
[ -n "$LIB" ] && export CAUSTICS_B200_LIB=$PWD/$LIB
for k in $KERNELS; do
  ncu --set full --clock-control none --import-source on --kernel-name-base function -k $k -s 0 -c 1 \
      -o gpurun_out/prof_${TAG}_$k -f python scripts/profile_targets.py ${NCU_TARGET:-c4} > gpurun_out/ncu_${TAG}_$k.log 2>&1
  ncu -i gpurun_out/prof_${TAG}_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_$k.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_$k.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/${TAG}_src_$k.csv 2>/dev/null
  rm -f gpurun_out/prof_${TAG}_$k.ncu-rep
done
