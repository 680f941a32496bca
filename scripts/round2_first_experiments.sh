#!/bin/bash
# First GPU experiment queued for the next round: predicted warm starts in the extended-source
# solver phases, guarded by the distance to the nearest other root.  Build the variants HERE (no GPU needed):
#     bash scripts/round2_first_experiments.sh build
# then on the GPU box (one call, ~2 min):
#     gpurun --timeout 600 -- 'bash scripts/round2_first_experiments.sh run'
# Output: C3 / C4 timings per variant and how the extended-source GPU tests fare with the predicting library.
set -e
cd "$(dirname "$0")/.."
case "$1" in
  build)
    rm -rf build_variants
    bash scripts/build_variant.sh a_base
    bash scripts/build_variant.sh b_predict_sep -DCB200_EXT_PREDICT=1 -DCB200_EXT_PREDICT_SEP=1e-2
    bash scripts/build_variant.sh c_predict_step -DCB200_EXT_PREDICT=1 -DCB200_EXT_PREDICT_MAXSTEP2=1e-6
    rm -rf build_variants/obj_*
    ;;
  run)
    mkdir -p gpurun_out
    python scripts/variant_bench_ext.py 2>&1 | tee gpurun_out/round2_predict_timings.txt
    for v in b_predict_sep c_predict_step; do
      echo "== extended-source GPU tests with $v"
      CAUSTICS_B200_LIB=$PWD/build_variants/$v.so timeout 400 python -m pytest tests/test_gpu_extended.py -q 2>&1 | tail -8
    done 2>&1 | tee gpurun_out/round2_predict_parity.txt
    ;;
  *) echo "usage: $0 build|run"; exit 2;;
esac
