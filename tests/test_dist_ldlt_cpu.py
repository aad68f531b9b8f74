"""Host-side logic of the multi-GPU 2-D block-cyclic LDL^T (pyipm_b200/dist_ldlt.py) on CPU: gloo backend,
world_size 2 (grids 1x2 and 2x1) plus the degenerate 1x1 grid, with the reference tile backend."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyipm_b200.dist_ldlt import BlockCyclicLDLT, choose_grid
from tests.ref_tile_ops import RefTileOps


def kkt_matrix(n, m, seed):
    rng = np.random.default_rng(seed)
    nh = n - m
    W = rng.standard_normal((nh, nh))
    K = np.zeros((n, n))
    K[:nh, :nh] = W @ W.T / nh + np.diag(10.0 ** rng.uniform(-2, 2, nh))
    J = rng.standard_normal((nh, m))
    K[:nh, nh:] = J
    K[nh:, :nh] = J.T
    K[np.arange(nh, n), np.arange(nh, n)] = -1e-6
    return K, rng.standard_normal((n, 3))


def _worker(rank, world, port, grid, n, b, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        K, rhs = kkt_matrix(n, n // 4, 5)
        F = BlockCyclicLDLT(n, grid, RefTileOps(b), block=b)
        F.load(K)
        inertia = F.factor()
        X = F.solve(rhs, nrefine=1)
        if rank == 0:
            np.savez(out, X=X, inertia=np.array(inertia))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize('grid', [(1, 2), (2, 1)])
def test_block_cyclic_ldlt_world2_gloo(grid, tmp_path):
    n, b = 64, 8
    out = str(tmp_path / 'res.npz')
    mp.spawn(_worker, args=(2, _free_port(), grid, n, b, out), nprocs=2, join=True)
    K, rhs = kkt_matrix(n, n // 4, 5)
    res = np.load(out)
    w = np.linalg.eigvalsh(K)
    assert tuple(res['inertia']) == (int(np.sum(w > 0)), int(np.sum(w < 0)), 0)
    Xref = np.linalg.solve(K, rhs)
    assert np.max(np.abs(res['X'] - Xref)) / np.max(np.abs(Xref)) < 1e-9


def test_block_column_cyclic_ldlt_world3_gloo(tmp_path):
    """1 x 3 grid: eight block columns over three ranks (ragged ownership), the look-ahead owner changes every step."""
    n, b = 64, 8
    out = str(tmp_path / 'res.npz')
    mp.spawn(_worker, args=(3, _free_port(), (1, 3), n, b, out), nprocs=3, join=True)
    K, rhs = kkt_matrix(n, n // 4, 5)
    res = np.load(out)
    w = np.linalg.eigvalsh(K)
    assert tuple(res['inertia']) == (int(np.sum(w > 0)), int(np.sum(w < 0)), 0)
    Xref = np.linalg.solve(K, rhs)
    assert np.max(np.abs(res['X'] - Xref)) / np.max(np.abs(Xref)) < 1e-9


def test_block_column_cyclic_deferred_bulk_update(tmp_path, monkeypatch):
    """B200IPM_DIST_DEFER=1 (default from 8 ranks on): the owner of the next block column postpones its bulk update of
    panel k until its own chain work is issued -- the order of updates per block column must stay the panel order."""
    monkeypatch.setenv('B200IPM_DIST_DEFER', '1')
    n, b = 96, 8
    out = str(tmp_path / 'res.npz')
    mp.spawn(_worker, args=(3, _free_port(), (1, 3), n, b, out), nprocs=3, join=True)
    K, rhs = kkt_matrix(n, n // 4, 5)
    res = np.load(out)
    w = np.linalg.eigvalsh(K)
    assert tuple(res['inertia']) == (int(np.sum(w > 0)), int(np.sum(w < 0)), 0)
    Xref = np.linalg.solve(K, rhs)
    assert np.max(np.abs(res['X'] - Xref)) / np.max(np.abs(Xref)) < 1e-9
    # ... and on one rank, where deferring serialises everything but must still be right
    F = BlockCyclicLDLT(n, (1, 1), RefTileOps(b), block=b)
    F.load(K)
    assert F.factor() == tuple(res['inertia'])
    X = F.solve(rhs, nrefine=1)
    assert np.max(np.abs(X - Xref)) / np.max(np.abs(Xref)) < 1e-9


def test_block_cyclic_ldlt_single_rank():
    n, b = 48, 8
    K, rhs = kkt_matrix(n, 12, 9)
    F = BlockCyclicLDLT(n, (1, 1), RefTileOps(b), block=b)
    F.load(K)
    inertia = F.factor()
    w = np.linalg.eigvalsh(K)
    assert inertia == (int(np.sum(w > 0)), int(np.sum(w < 0)), 0)
    X = F.solve(rhs, nrefine=1)
    Xref = np.linalg.solve(K, rhs)
    assert np.max(np.abs(X - Xref)) / np.max(np.abs(Xref)) < 1e-9


def test_choose_grid():
    assert choose_grid(1) == (1, 1) and choose_grid(2) == (1, 2) and choose_grid(4) == (1, 4) and choose_grid(8) == (1, 8)
