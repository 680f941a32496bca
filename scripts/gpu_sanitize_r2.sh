#!/bin/bash
# compute-sanitizer over the round-2 extended-source kernels + the randomised differential run (under gpurun).
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
K="degenerate or small_batch_variants or contour_invariants or binary_other_sampling or test_triple"
timeout 700 $S --tool memcheck python -m pytest tests/test_gpu_extended.py -q -x -k "$K" > gpurun_out/r02_memcheck.log 2>&1
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r02_memcheck.log | tail -3
timeout 500 $S --tool racecheck python -m pytest tests/test_gpu_extended.py -q -x -k "degenerate or contour_invariants or binary_limb_darkened" > gpurun_out/r02_racecheck.log 2>&1
grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/r02_racecheck.log | tail -3
timeout 300 $S --tool synccheck python -m pytest tests/test_gpu_extended.py -q -x -k "degenerate or binary_limb_darkened" > gpurun_out/r02_synccheck.log 2>&1
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r02_synccheck.log | tail -3
# a large-batch call under memcheck (thread-per-source limb walk, sweep, persistent open pass): 20 000 triple-lens sources
timeout 600 $S --tool memcheck python - > gpurun_out/r02_memcheck_big.log 2>&1 <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, ".")
import caustics_b200 as cb
w = torch.from_numpy(np.linspace(-2, 2, 20000) + 0.1j).cuda()
m = cb.mag_extended_source(w, 1e-2, nlenses=3, npts_limb=200, s=0.9, q=0.2, q3=0.1, r3=0.8, psi=1.0)
print("finite", bool(torch.isfinite(m).all()), float(m.max()))
PY
grep -E "finite|ERROR SUMMARY" gpurun_out/r02_memcheck_big.log | tail -2
python tests/fuzz_extended.py 12 24 0 0 > gpurun_out/r02_fuzz_extended.txt 2>&1; tail -4 gpurun_out/r02_fuzz_extended.txt
python tests/fuzz_extended.py 12 24 1 1 >> gpurun_out/r02_fuzz_extended.txt 2>&1; tail -4 gpurun_out/r02_fuzz_extended.txt
