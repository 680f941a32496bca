"""Picklable oracle calls for process pools in the GPU tests.  Imports NumPy and the oracle only, so
that `spawn`ed workers start quickly and never touch torch/CUDA (forking a process that holds a CUDA
context and helper threads can deadlock)."""
import multiprocessing as mp
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np

from oracle import extended

HP2 = dict(s=0.9, q=0.2)


def c4_hp():
    a, e1, e2, r3c = 0.698, 0.02809, 0.9687, -0.0197 - 0.95087j      # the C2/C4 lens, low-level
    q = e2 / e1
    return dict(s=2 * a, q=q, q3=q / e1 - 1 - q, r3=abs(r3c), psi=float(np.angle(r3c)))


def oracle_ld_c3(x):
    return extended.mag_extended_source(x, 1e-2, 2, 200, True, 0.7, 100, **HP2)


def oracle_c4(x):
    return extended.mag_extended_source(x, 1e-2, 3, 200, **c4_hp())


def pool_map(fn, xs, chunksize=8):
    with ProcessPoolExecutor(min(16, os.cpu_count() or 1), mp_context=mp.get_context("spawn")) as ex:
        return np.array(list(ex.map(fn, xs, chunksize=chunksize)))
