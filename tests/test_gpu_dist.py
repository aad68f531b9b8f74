"""CUDA tile backend of the block-cyclic LDL^T (pyipm_b200/dist_ldlt.py) on one GPU (1x1 grid): the same code
path the multi-GPU run takes minus the NCCL broadcasts (those are covered on CPU with gloo)."""
import numpy as np
import pytest

from pyipm_b200 import problems
from pyipm_b200.dist_ldlt import BlockCyclicLDLT, CudaTileOps

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('n,m', [(256, 32), (768, 128), (2048, 256)])
def test_block_cyclic_ldlt_cuda_backend(n, m):
    K, rhs = problems.make_dense_kkt(n, m, seed=n)
    F = BlockCyclicLDLT(n, (1, 1), CudaTileOps(0), block=256)
    F.load(K)
    inertia = F.factor()
    assert inertia == (n - m, m, 0)
    X = F.solve(rhs, nrefine=1)
    res = np.max(np.abs(K @ X - rhs)) / (np.max(np.abs(K)) * np.max(np.abs(X)))
    assert res < 1e-13, res
    Xref = np.linalg.solve(K, rhs)
    assert np.max(np.abs(X - Xref)) / np.max(np.abs(Xref)) < 1e-6
