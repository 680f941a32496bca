"""Golden vectors for the magnification map (BASELINE config C5) from the REFERENCE's own Python (through
oracle/refshim.py) and its compiled solver -> tests/golden/map_golden.npz.  Build container only:
    python tests/golden/make_golden_map.py

Two 41 x 40 patches of the 10^4 x 10^4 grid on [-1.5, 1.5]^2, each with a fold of the lens's caustic running
through it: the reference's mag_point_source (point_source.py:1762-1830) on the explicit grid, plain and with
roots_compensated=True."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
warnings.filterwarnings("ignore")
from oracle import refshim  # noqa: E402

C = refshim.install()
A = refshim.arr

dx = 3.0 / 9999
out = {}
for key, nl, hp, col0, r0 in (("b", 2, dict(s=0.9, q=0.2), 4290, 5300),
                              ("t", 3, dict(s=0.9, q=0.2, q3=0.1, r3=0.8, psi=1.0), 4350, 6180)):
    ix, iy = np.meshgrid(np.arange(40), np.arange(r0, r0 + 41))
    x0 = -1.5 + col0 * dx
    w = (x0 + ix * (2 * dx)) + 1j * (-1.5 + iy * dx)
    out[f"{key}_spec"] = np.array([x0, -1.5, 2 * dx, dx, 40, r0, r0 + 41])
    out[f"{key}_mag"] = np.asarray(C.mag_point_source(A(w.reshape(-1).copy()), nlenses=nl, **hp)).reshape(w.shape)
    out[f"{key}_mag_comp"] = np.asarray(C.mag_point_source(A(w.reshape(-1).copy()), nlenses=nl, roots_compensated=True, **hp)).reshape(w.shape)
    print(key, out[f"{key}_mag"].max(), np.abs(out[f"{key}_mag"] / out[f"{key}_mag_comp"] - 1).max())
np.savez_compressed(os.path.join(HERE, "map_golden.npz"), **out)
