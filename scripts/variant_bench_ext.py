#!/usr/bin/env python
"""C3/C4 timing for every library build in build_variants/ (experiments)."""
import glob, os, subprocess, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for lib in sorted(glob.glob(os.path.join(ROOT, "build_variants", "*.so"))):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "bench_configs.py"), "--only", "C3,C4"],
                       env=dict(os.environ, CAUSTICS_B200_LIB=lib), capture_output=True, text=True)
    print(os.path.basename(lib))
    for line in r.stdout.strip().splitlines():
        print("   ", line[:170])
    if r.returncode: print(r.stderr[-300:])
