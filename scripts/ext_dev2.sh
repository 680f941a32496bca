#!/bin/bash
# Dev loop for kernel family 3 (run under gpurun): parity tests, tangent probe, per-kernel times per build variant,
# batch-size sweep.
mkdir -p gpurun_out
TAG=${TAG:-dev3}
python -m pytest tests/test_gpu_extended.py tests/test_gpu_round2.py -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1
tail -8 gpurun_out/${TAG}_tests.log
python scripts/tangent_probe.py > gpurun_out/${TAG}_tangent.jsonl 2>&1
cat gpurun_out/${TAG}_tangent.jsonl | cut -c1-600
TARGET=c4 bash scripts/variant_kernel_times.sh 2>&1 | tee gpurun_out/${TAG}_variants.log
MASKS="${MASKS:--1,32}" python scripts/variant_sweep_ext.py > gpurun_out/${TAG}_sweep.jsonl 2> gpurun_out/${TAG}_sweep.err
cat gpurun_out/${TAG}_sweep.jsonl
python scripts/bench_configs.py --only C3,C4 2>&1 | cut -c1-300
