"""GPU tier: the magnification-map entry (BASELINE config C5) against golden values from the REFERENCE's own
mag_point_source (tests/golden/map_golden.npz, tests/golden/make_golden_map.py)."""
import numpy as np
import pytest
import torch

from conftest import TRIPLE_HP

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb(built_lib):
    import caustics_b200
    assert torch.cuda.is_available()
    return caustics_b200


@pytest.mark.parametrize("key,nl,hp", [("b", 2, dict(s=0.9, q=0.2)), ("t", 3, TRIPLE_HP)])
def test_map_vs_reference_golden(cb, key, nl, hp):
    """the map entry, cold and walked, against the REFERENCE's own mag_point_source on two caustic-crossing patches
    of the C5 grid (tests/golden/map_golden.npz: the reference's Python + compiled solver); tolerance as in
    tests/test_hostsim_extended.py::test_grid_walk_vs_reference_map"""
    import os
    from conftest import ROOT
    g = np.load(os.path.join(ROOT, "tests", "golden", "map_golden.npz"))
    x0, y0, dx, dy, nx, r0, r1 = g[key + "_spec"]
    nx, r0, r1 = int(nx), int(r0), int(r1)
    want = g[key + "_mag"]
    # (conditioning term doubled here: the host-compiled device code sits at 0.4 of the CPU test's bound)
    tol = 1e-10 + 10 * np.abs(want / g[key + "_mag_comp"] - 1) + (2e-15 if nl == 2 else 4e-14) * want**2
    for walk in (False, True):
        for comp in (False, True):
            got = cb.mag_point_source_map(x0, y0, dx, dy, nx, r1, nlenses=nl, rows=(r0, r1), walk=walk,
                                          roots_compensated=comp, **hp).cpu().numpy()
            rel = np.abs(got / want - 1)
            assert (rel <= tol).all() and np.median(rel) < 1e-13, (walk, comp, rel.max())
