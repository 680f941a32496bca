#!/bin/bash
# Dev loop for kernel family 3 (run under gpurun): parity tests, batch-size sweep, per-kernel launch times.
mkdir -p gpurun_out
TAG=${TAG:-dev}
python -m pytest tests/test_gpu_extended.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1
tail -15 gpurun_out/${TAG}_tests.log
MASKS="${MASKS:--1}" python scripts/variant_sweep_ext.py > gpurun_out/${TAG}_sweep.jsonl 2> gpurun_out/${TAG}_sweep.err
cat gpurun_out/${TAG}_sweep.jsonl
python scripts/bench_configs.py --only C3,C4 2>&1 | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_c4.csv \
    python scripts/profile_targets.py c4 > gpurun_out/${TAG}_ncu_c4.log 2>&1
python - <<'PY'
import csv, collections, os
tag = os.environ.get("TAG", "dev")
rows = list(csv.reader(l for l in open(f"gpurun_out/{tag}_launches_c4.csv") if l.startswith('"')))
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(",", "")); v = v / 1e3 if r[iu] in ("ns", "nsecond") else v
    a = agg.setdefault(r[ik][:60], [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in agg.items():
    print(f"{k:60s} n={n:4d} total={t/1e3:9.3f} ms")
PY

if [ -n "$NCU_KERNELS" ]; then
  for k in $NCU_KERNELS; do
    ncu --set full --clock-control none --import-source on --kernel-name-base function -k $k -s 0 -c 1 \
        -o gpurun_out/prof_${TAG}_$k -f python scripts/profile_targets.py ${NCU_TARGET:-c4} > gpurun_out/ncu_${TAG}_$k.log 2>&1
    ncu -i gpurun_out/prof_${TAG}_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_$k.csv 2>/dev/null
    ncu -i gpurun_out/prof_${TAG}_$k.ncu-rep --page source --csv > gpurun_out/${TAG}_src_$k.csv 2>/dev/null
    rm -f gpurun_out/prof_${TAG}_$k.ncu-rep
  done
fi
