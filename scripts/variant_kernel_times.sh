#!/bin/bash
# Per-kernel launch times (ncu, serialised) and the C4 step (CUDA events) for every build in build_variants/.
#   TARGET=c4 bash scripts/variant_kernel_times.sh
mkdir -p gpurun_out
T=${TARGET:-c4}
for lib in build_variants/*.so; do
  name=$(basename $lib .so)
  echo "=== $name"
  CAUSTICS_B200_LIB=$PWD/$lib python scripts/bench_configs.py --only C4 2>&1 | cut -c1-200
  CAUSTICS_B200_LIB=$PWD/$lib ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/vk_${name}.csv python scripts/profile_targets.py $T > /dev/null 2>&1
  python - "$name" <<'PY'
import csv, collections, sys
name = sys.argv[1]
rows = list(csv.reader(l for l in open(f"gpurun_out/vk_{name}.csv") if l.startswith('"')))
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(",", "")); v = v / 1e3 if r[iu] in ("ns", "nsecond") else v
    a = agg.setdefault(r[ik][:50], [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in agg.items():
    print(f"   {k:50s} n={n:3d} mean={t/n/1e3:9.3f} ms")
PY
done
