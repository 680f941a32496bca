#!/bin/bash
# Run on the GPU box (via gpurun): build, GPU tests, smoke, short bench.  Outputs -> gpurun_out/
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -${TAIL:-25}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -2 > gpurun_out/bench_latest.json; cat gpurun_out/bench_latest.json
