#!/bin/bash
# ncu evidence for the map-walk kernel (config C5 through CAUSTICS_FLAG_GRID_WALK).  Outputs -> gpurun_out/
set -x
mkdir -p gpurun_out
cat > /tmp/walk_once.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import caustics_b200 as cb
dx = 3.0 / 9999
for walk in (False, True):
    for _ in range(2):
        cb.mag_point_source_map(-1.5, -1.5, dx, dx, 10_000, 10_000, rows=(4000, 5000), walk=walk, s=0.9, q=0.2)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:ps_grid_walk -s 1 -c 1 -o gpurun_out/prof_walk -f python /tmp/walk_once.py > gpurun_out/ncu_walk.log 2>&1
ncu -i gpurun_out/prof_walk.ncu-rep --page raw --csv > gpurun_out/ncu_walk_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_walk.log
ls -la gpurun_out
