#!/bin/bash
# Run on an N-GPU box (gpurun --gpus N -- 'bash scripts/gpu_multi.sh N'): the 2-GPU tests, the PCIe probe and
# bench.py at 1..N GPUs (the driver's scaling run).  Outputs -> gpurun_out/
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
[ -n "$SKIP_TESTS" ] || timeout 300 python -m pytest tests/test_gpu_round2.py -q -x 2>&1 | tail -5
[ -n "$SKIP_PROBE" ] || timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    scripts/pcie_probe.py 2> gpurun_out/pcie_probe_n$N.err | tail -1 > gpurun_out/pcie_probe_n$N.json
for k in ${KS:-1 2 4 8}; do
  [ $k -le $N ] || continue
  if [ $k -eq 1 ]; then
    timeout 400 python bench.py --steps ${STEPS:-20} --warmup 3 2> gpurun_out/scale_n$k.err | tail -1 > gpurun_out/scale_n$k.json
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2951$k \
        bench.py --gpus $k --steps ${STEPS:-20} --warmup 3 2> gpurun_out/scale_n$k.err | tail -1 > gpurun_out/scale_n$k.json
  fi
  tail -c 300 gpurun_out/scale_n$k.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/scale_n$k.json"))
    print("N=$k headline", d["value"], "e2e", d["e2e"]["value"])
    for key, v in d["configs"].items():
        print("  ", key, "%.4g" % v["value"], "ms", "%.3f" % v["ms_per_step"], "no-gather ms", v.get("ms_no_gather"),
              "e2e", "%.4g" % v.get("e2e", {}).get("value", 0), "same", v.get("matches_single_gpu"), v.get("max_rel_dev_vs_single_gpu"))
except Exception as e:
    print("N=$k: no line", e)
PY
done
[ -n "$SKIP_PROBE" ] || python -c "
import json; d=json.load(open('gpurun_out/pcie_probe_n$N.json'))
for k,v in d['results'].items(): print(k, '%.1f GB/s per rank, %.1f aggregate' % (v['per_rank_gbs'], v['aggregate_gbs']))
print('\n'.join(d['topo']))
"
