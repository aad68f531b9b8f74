"""Cycle breakdown of the tile kernel (library compiled with -DTILE_PROF): python tools/probe/tile_prof.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pyipm_b200 import _lib
rng = np.random.default_rng(0)
n = 128
B = rng.standard_normal((n, n))
A = B @ B.T / n + np.eye(n)
F = _lib.DenseLDLT(n)
print(F.factor(A))
F.close()
