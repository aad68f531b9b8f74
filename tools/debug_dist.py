import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyipm_b200 import problems, _lib
from pyipm_b200.dist_ldlt import BlockCyclicLDLT, CudaTileOps
from tests.ref_tile_ops import RefTileOps
n, m = 2048, 256
K, rhs = problems.make_dense_kkt(n, m, seed=n)
F1 = _lib.DenseLDLT(n); print('single-GPU ldlt_factor inertia', F1.factor(K))
Fc = BlockCyclicLDLT(n, (1, 1), CudaTileOps(0), block=256); Fc.load(K); print('cuda backend', Fc.factor())
Fr = BlockCyclicLDLT(n, (1, 1), RefTileOps(256), block=256); Fr.load(K); print('ref backend', Fr.factor())
for k in range(n // 256):
    cc = Fc.ops.counts(Fc.diags[k]); cr = Fr.ops.counts(Fr.diags[k])
    # compare the Schur-complement diagonal blocks through their D eigen-signs and the panel norms
    pn_c = float(Fc.panels[k].abs().max()) if Fc.panels[k] is not None else 0.0
    pn_r = float(Fr.panels[k].abs().max()) if Fr.panels[k] is not None else 0.0
    print(k, cc, cr, 'max|L| cuda %.3e ref %.3e' % (pn_c, pn_r))
