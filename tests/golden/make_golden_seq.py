"""Golden vectors for _images_point_source_sequential from the REFERENCE's own Python (through
oracle/refshim.py) and its compiled solver -> tests/golden/seq_golden.npz.  Build container only:
    python tests/golden/make_golden_seq.py"""
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
from caustics import point_source as PS  # noqa: E402  (the reference module, under the NumPy shim)

out = {}
# a source-limb-like circle crossing the caustic (binary) and a straight trajectory (triple)
th = np.linspace(-np.pi, np.pi, 60)
paths = {
    "b": (2, dict(a=0.45, e1=1 / 1.2), 0.11 + 0.1j + 0.05 * np.exp(1j * th)),
    "t": (3, dict(a=0.698, e1=0.02809, e2=0.9687, r3=-0.0197 - 0.95087j), np.linspace(-0.4, 0.4, 48) + 0.1j),
}
for k, (nl, p, w) in paths.items():
    z, m = PS._images_point_source_sequential(A(w), nlenses=nl, **p)
    out[f"{k}_w"] = w
    out[f"{k}_z"] = np.asarray(A(z), dtype=np.complex128)
    out[f"{k}_mask"] = np.asarray(A(m)).astype(bool)
    print(k, out[f"{k}_z"].shape, out[f"{k}_mask"].sum(axis=0)[:12])
np.savez_compressed(os.path.join(HERE, "seq_golden.npz"), **out)
