#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per (kernel, grid):
    python scripts/launch_summary.py gpurun_out/launches_x.csv > profiles/x_summary.txt"""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ik, im, iv, ii, ig = (h.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Grid Size"))
agg = collections.OrderedDict()
ours = re.compile(r"ea_kernel|ps_|k_[a-z_]+|fp64_peak|match_tracks|jvp|vjp|lc_|traj|peer")
n = 0
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    n += 1
    name = re.sub(r"^void |\(anonymous namespace\)::|<unnamed>::|cb200::", "", r[ik]).split("(")[0]
    key = (name[:46], r[ig]) if ours.search(name) else ("[torch setup / comparison kernels]", "")
    a = agg.setdefault(key, [0, 0.0, int(r[ii]), int(r[ii])])
    a[0] += 1
    a[1] += float(r[iv].replace(",", "")) / 1e6
    a[3] = int(r[ii])
print(f"{n} launches in total (cold-cache, serialised under ncu: compare SHARES, not absolute times)\n")
print(f"{'kernel':46s} {'grid':>16s} {'launches':>9s} {'total ms':>10s} {'mean us':>10s}  first..last ID")
for (k, g), a in agg.items():
    print(f"{k:46s} {g:>16s} {a[0]:9d} {a[1]:10.3f} {a[1] / a[0] * 1e3:10.1f}  {a[2]}..{a[3]}")
