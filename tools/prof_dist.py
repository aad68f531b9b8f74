"""Per-column timeline of the block-column-cyclic LDL^T (config 4) on the ranks of one node: where does a step of the chain
go?  torchrun --nproc-per-node N tools/prof_dist.py [n]   (measurement infrastructure)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from pyipm_b200.dist_ldlt import BlockCyclicLDLT, CudaTileOps, choose_grid
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
m = n // 8; nh = n - m
g = torch.Generator(device='cuda'); g.manual_seed(16384)
W = torch.randn(nh, nh, dtype=torch.float64, device='cuda', generator=g)
K = torch.zeros(n, n, dtype=torch.float64, device='cuda')
K[:nh, :nh] = W @ W.t() / nh
del W
K[:nh, :nh].diagonal().add_(10.0 ** (8.0 * torch.rand(nh, dtype=torch.float64, device='cuda', generator=g) - 4.0))
J = torch.randn(nh, m, dtype=torch.float64, device='cuda', generator=g)
K[:nh, nh:] = J; K[nh:, :nh] = J.t(); K[nh:, nh:].diagonal().fill_(-1e-8)
F = BlockCyclicLDLT(n, choose_grid(world), CudaTileOps(local), block=256)
F.load_device(K)
for _ in range(3):
    F.factor()
F.profile = True
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); print('inertia', F.factor()) if rank == 0 else F.factor(); e1.record(); torch.cuda.synchronize()
ev = F.profile_events
base = ev[0][5]
def T(a, b):
    return a.elapsed_time(b) if a is not None and b is not None else float('nan')
tot = dict(diag=0.0, panel=0.0, bcast=0.0, look=0.0, rest=0.0)
lines = []
for (k, owner, t0, t1, t2, tb0, tb1, tu0, tu1, tu2) in ev:
    d, p_, bc = T(t0, t1), T(t1, t2), T(tb0, tb1)
    lk, rs = T(tu0, tu1), T(tu1 if tu1 is not None else tu0, tu2)
    for key, v in (('diag', d), ('panel', p_), ('bcast', bc), ('look', lk), ('rest', rs)):
        if v == v:
            tot[key] += v
    lines.append('%3d own %d | chain start %7.2f diag %6.3f panel %6.3f | bcast %7.2f..%7.2f (%6.3f) | upd start %7.2f look %6.3f rest %6.3f end %7.2f'
                 % (k, owner, T(base, t0), d, p_, T(base, tb0), T(base, tb1), bc, T(base, tu0), lk, rs, T(base, tu2)))
for r in range(world):
    if r == rank:
        print('== rank %d: factor %.2f ms; sums (ms): %s' % (rank, e0.elapsed_time(e1), {k: round(v, 2) for k, v in tot.items()}))
        if rank == 0:
            print('\n'.join(lines[:24]))
    if world > 1:
        dist.barrier()
if world > 1:
    dist.destroy_process_group()
