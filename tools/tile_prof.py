"""Cycle breakdown of ldlt_tile_kernel (library built with -DTILE_PROF, B200IPM_LIB points at it)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyipm_b200 import _lib
rng = np.random.default_rng(0)
n = 128
B = rng.standard_normal((n, n)); A = B @ B.T / n + np.eye(n)
F = _lib.DenseLDLT(n)
print(F.factor(A))
