"""Block-column-cyclic LDL^T on N ranks against the same pipeline on ONE rank (same kernels, same order of updates per block
column => the factors must agree to rounding): per-column max deviation of L, repeated runs, residual of the solve.
torchrun --nproc-per-node N tools/check_dist.py [n]   (test infrastructure)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from pyipm_b200.dist_ldlt import BlockCyclicLDLT, CudaTileOps, choose_grid
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
singles = [dist.new_group(ranks=[r]) for r in range(world)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m = n // 8; nh = n - m
g = torch.Generator(device='cuda'); g.manual_seed(16384)
W = torch.randn(nh, nh, dtype=torch.float64, device='cuda', generator=g)
K = torch.zeros(n, n, dtype=torch.float64, device='cuda')
K[:nh, :nh] = W @ W.t() / nh
del W
K[:nh, :nh].diagonal().add_(10.0 ** (8.0 * torch.rand(nh, dtype=torch.float64, device='cuda', generator=g) - 4.0))
J = torch.randn(nh, m, dtype=torch.float64, device='cuda', generator=g)
K[:nh, nh:] = J; K[nh:, :nh] = J.t(); K[nh:, nh:].diagonal().fill_(-1e-8)
rhs = torch.randn(8, n, dtype=torch.float64, device='cuda', generator=g)
F1 = BlockCyclicLDLT(n, (1, 1), CudaTileOps(local), block=256, group=singles[rank])
F1.load_device(K)
F1.factor()
ref = [p.clone() if p is not None else None for p in F1.panels]
X1 = F1.solve_device(rhs, nrefine=1)
r1 = float((rhs - F1.matvec(X1)).abs().max() / (K.abs().max() * X1.abs().max()))
FN = BlockCyclicLDLT(n, choose_grid(world), CudaTileOps(local), block=256)
FN.load_device(K)
for trial in range(3):
    FN.factor()
    dev = [float((FN.panels[k] - ref[k]).abs().max()) if ref[k] is not None else 0.0 for k in range(len(ref))]
    XN = FN.solve_device(rhs, nrefine=1)
    X0 = FN.solve_device(rhs, nrefine=0)
    rN = float((rhs - FN.matvec(XN)).abs().max() / (K.abs().max() * XN.abs().max()))
    r0 = float((rhs - FN.matvec(X0)).abs().max() / (K.abs().max() * X0.abs().max()))
    bad = [(k, d) for k, d in enumerate(dev) if d > 0]
    if rank == 0 or bad:
        print('rank %d trial %d: cols with deviation %d (first %s, max %.3e); resid N %.3e (unrefined %.3e) single %.3e; |X| %.3e dX %.3e'
              % (rank, trial, len(bad), bad[:3], max(dev), rN, r0, r1, float(XN.abs().max()), float((XN - X1).abs().max())), flush=True)
dist.barrier()
dist.destroy_process_group()
