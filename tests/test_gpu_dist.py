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


def test_pipeline_config4_size_chain_kernels_under_saturating_updates():
    """Config-4 size (order 16384) on one rank: the look-ahead pipeline factors block column k + 1 on the high-priority chain
    stream while the trailing update of panel k saturates every SM, so the CTAs of the chain's small kernels start at
    different times.  Regression test of the 4-CTA mini step (its CTAs overwrite in place what their siblings read; a
    cluster barrier now separates the loads from the stores): the scaled residual after one refinement sweep was 1e-11
    with the race, 4e-18 without; the factor must also be reproducible run to run."""
    import torch
    n, m = 16384, 2048
    nh = n - m
    g = torch.Generator(device='cuda')
    g.manual_seed(16384)
    W = torch.randn(nh, nh, dtype=torch.float64, device='cuda', generator=g)
    K = torch.zeros(n, n, dtype=torch.float64, device='cuda')
    K[:nh, :nh] = W @ W.t() / nh
    del W
    K[:nh, :nh].diagonal().add_(10.0 ** (8.0 * torch.rand(nh, dtype=torch.float64, device='cuda', generator=g) - 4.0))
    J = torch.randn(nh, m, dtype=torch.float64, device='cuda', generator=g)
    K[:nh, nh:] = J
    K[nh:, :nh] = J.t()
    K[nh:, nh:].diagonal().fill_(-1e-8)
    rhs = torch.randn(4, n, dtype=torch.float64, device='cuda', generator=g)
    F = BlockCyclicLDLT(n, (1, 1), CudaTileOps(0), block=256)
    F.load_device(K)
    assert F.factor() == (nh, m, 0)
    first = [p.clone() for p in F.panels[:8]]
    X = F.solve_device(rhs, nrefine=1)
    res = float((rhs - F.matvec(X)).abs().max() / (K.abs().max() * X.abs().max()))
    assert res < 1e-15, res
    assert F.factor() == (nh, m, 0)
    for a, b in zip(first, F.panels[:8]):
        assert torch.equal(a, b)
