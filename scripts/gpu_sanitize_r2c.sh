#!/bin/bash
# compute-sanitizer over the kernels changed in the last session of round 2 (under gpurun) + the randomised differential run
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  N=${N:-300} timeout ${TMO:-400} $S --tool $tool python scripts/sanitize_target.py > gpurun_out/r02c_$tool.log 2>&1
  echo "== $tool"; grep -E "^plain|^grad|^ld|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02c_$tool.log | tr '\n' ';'; echo
done
python tests/fuzz_extended.py 12 24 0 0 > gpurun_out/r02c_fuzz_extended.txt 2>&1; tail -1 gpurun_out/r02c_fuzz_extended.txt
python tests/fuzz_extended.py 12 24 1 1 >> gpurun_out/r02c_fuzz_extended.txt 2>&1; tail -1 gpurun_out/r02c_fuzz_extended.txt
