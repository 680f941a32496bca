#!/bin/bash
# Per-kernel times of one C4 call (ncu, serialised) for the variant masks in $MASKS, then the batch-size sweep.
#   LIB=build_variants/x.so MASKS="-1 16" TAG=dev bash scripts/ext_times.sh
mkdir -p gpurun_out
[ -n "$LIB" ] && export CAUSTICS_B200_LIB=$PWD/$LIB
for m in ${MASKS:--1}; do
  echo "=== mask $m"
  EXT_MASK=$m ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active \
      --clock-control none -c 40 --csv --log-file gpurun_out/${TAG:-dev}_mask$m.csv python scripts/profile_targets.py ${TARGET:-c4} > /dev/null 2>&1
  python - gpurun_out/${TAG:-dev}_mask$m.csv <<'PY'
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ik = hdr.index("Kernel Name"); im = hdr.index("Metric Name"); iv = hdr.index("Metric Value"); ii = hdr.index("ID")
agg = collections.OrderedDict()
for r in rows[1:]:
    a = agg.setdefault((r[ii], r[ik][:44]), {})
    a[r[im]] = float(r[iv].replace(",", ""))
seen = set()
for (i, k), a in agg.items():
    if k in seen: continue
    seen.add(k)
    print(f"   {k:44s} {a['gpu__time_duration.sum']/1e6:8.3f} ms  {a['smsp__inst_executed.sum']/1e9:6.2f} Ginst  fp64 {a['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']:5.1f} %")
PY
done
[ -n "$SWEEP" ] && MASKS="$SWEEP" python scripts/variant_sweep_ext.py
