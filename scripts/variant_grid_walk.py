#!/usr/bin/env python
"""Times the magnification-map entry (config C5) cold and as warm-started walks for several run lengths,
with and without extrapolation (CAUSTICS_B200_GRID_RUN / _GRID_EXTRAP are read at every launch)."""
import json
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import caustics_b200 as cb  # noqa: E402


def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    out = {}
    dx = 3.0 / 9999
    for name, nl, hp, rows in (("binary", 2, dict(s=0.9, q=0.2), (4000, 6000)),
                               ("triple", 3, dict(s=0.9, q=0.2, q3=0.1, r3=0.8, psi=1.0), (5800, 6300))):
        n = 10_000 * (rows[1] - rows[0])
        kw = dict(nlenses=nl, rows=rows, **hp)
        os.environ.pop("CAUSTICS_B200_GRID_RUN", None); os.environ.pop("CAUSTICS_B200_GRID_EXTRAP", None)
        ms = t(lambda: cb.mag_point_source_map(-1.5, -1.5, dx, dx, 10_000, 10_000, walk=False, **kw))
        out[f"{name} cold"] = {"ms": ms, "evals_per_s": n / ms * 1e3}
        for ex in (1, 0):
            for run in (8, 16, 32, 64, 128):
                os.environ["CAUSTICS_B200_GRID_RUN"] = str(run); os.environ["CAUSTICS_B200_GRID_EXTRAP"] = str(ex)
                ms = t(lambda: cb.mag_point_source_map(-1.5, -1.5, dx, dx, 10_000, 10_000, walk=True, **kw))
                out[f"{name} walk run={run} extrap={ex}"] = {"ms": ms, "evals_per_s": n / ms * 1e3}
    import numpy as np
    wd = torch.from_numpy(np.linspace(-2, 2, 1_000_000) + 0.1j).cuda()
    C2P = dict(a=0.698, e1=0.02809, e2=0.9687, r3=-0.0197 - 0.95087j)
    L = cb._lib.lib()
    mag = torch.empty(wd.numel(), dtype=torch.float64, device="cuda")
    for name, lens_c in (("path binary C1 lens", cb.point_source._c_lens(2, cb.lens_params(2, s=0.9, q=0.2)[1], **cb.lens_params(2, s=0.9, q=0.2)[0])),
                         ("path triple C2 lens", cb.point_source._c_lens(3, 0.0, **C2P))):
        n = wd.numel()
        os.environ.pop("CAUSTICS_B200_PATH_RUN", None); os.environ.pop("CAUSTICS_B200_GRID_EXTRAP", None)
        ms = t(lambda: cb._lib.check(L.caustics_mag_point_source(wd.data_ptr(), mag.data_ptr(), None, n, lens_c, 2500, 0, 0, None)))
        out[f"{name} n=1e6 cold"] = {"ms": ms, "evals_per_s": n / ms * 1e3}
        for run in (0, 2, 4, 8, 16, 32):
            if run: os.environ["CAUSTICS_B200_PATH_RUN"] = str(run)
            ms = t(lambda: cb._lib.check(L.caustics_mag_point_source(wd.data_ptr(), mag.data_ptr(), None, n, lens_c, 2500, 0, 8, None)))
            out[f"{name} n=1e6 walk run={run or 'auto'}"] = {"ms": ms, "evals_per_s": n / ms * 1e3}
    for k, v in out.items():
        print(f"{k:34s} {v['ms']:9.3f} ms  {v['evals_per_s']:.4g} evals/s")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "variant_grid_walk.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
