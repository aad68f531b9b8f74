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


def _nccl_worker(rank, world, port, n, m, out):
    import os
    import torch
    import torch.distributed as dist
    from pyipm_b200.dist_ldlt import choose_grid
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        K, rhs = problems.make_dense_kkt(n, m, seed=n)
        F = BlockCyclicLDLT(n, choose_grid(world), CudaTileOps(rank), block=256)
        F.load(K)
        inertia = F.factor()
        X = F.solve(rhs, nrefine=1)
        if rank == 0:
            np.savez(out, X=X, inertia=np.array(inertia))
    finally:
        dist.destroy_process_group()


def test_block_cyclic_ldlt_two_ranks_nccl(tmp_path):
    """The multi-GPU path proper: 2 ranks, NCCL panel broadcasts, look-ahead pipeline (skipped on a 1-GPU box)."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    n, m = 2048, 256
    out = str(tmp_path / 'res.npz')
    mp.spawn(_nccl_worker, args=(2, port, n, m, out), nprocs=2, join=True)
    K, rhs = problems.make_dense_kkt(n, m, seed=n)
    res = np.load(out)
    assert tuple(res['inertia']) == (n - m, m, 0)
    Xref = np.linalg.solve(K, rhs)
    assert np.max(np.abs(res['X'] - Xref)) / np.max(np.abs(Xref)) < 1e-6
    r = np.max(np.abs(K @ res['X'] - rhs)) / (np.max(np.abs(K)) * np.max(np.abs(res['X'])))
    assert r < 1e-13
