#!/usr/bin/env python
"""Per-kernel device times of one profile target in a NORMAL run (CUPTI activity records through torch.profiler:
kernels run back to back with warm caches, unlike the serialised, cache-flushed launches under ncu).

    python scripts/kernel_times.py <target of scripts/profile_targets.py>
"""
import os
import runpy
import sys

import torch
from torch.profiler import ProfilerActivity, profile

here = os.path.dirname(os.path.abspath(__file__))
sys.argv = [os.path.join(here, "profile_targets.py")] + sys.argv[1:]
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    runpy.run_path(sys.argv[0], run_name="__main__")
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in ev:
    a = agg.setdefault(e.name[:70], [0, 0.0])
    a[0] += 1
    a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    if t / tot > 0.003:
        print(f"{k:70s} n={n:4d} mean {t / n / 1e3:9.3f} ms  share {t / tot:5.3f}")
print(f"total kernel time {tot / 1e3:.3f} ms")
