set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dist.py tests/test_gpu_engine.py -q -x > gpurun_out/r2_t13.log 2>&1; tail -5 gpurun_out/r2_t13.log
timeout 300 python tools/prof_solve.py 2>&1 | tail -1
B200IPM_LDLT_TMA=0 timeout 300 python tools/prof_solve.py 2>&1 | tail -1
timeout 300 python tools/trace_factor.py 3 2>&1 | tail -10
timeout 600 python bench.py --workload c4 --steps 3 --warmup 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c4 1gpu factor_ms', d['factor_ms'], d['solve_ms_8rhs_1refine'], d['inertia'], d['scaled_residual_inf'])"
python - <<'PY'
import numpy as np, torch, time, sys
sys.path.insert(0,'.')
from pyipm_b200 import _lib
import ctypes as C
n=16384
g=torch.Generator(device='cuda'); g.manual_seed(1)
W=torch.randn(n,n,dtype=torch.float64,device='cuda',generator=g); K=W@W.t()/n; del W
K.diagonal().add_(1.0)
F=_lib.DenseLDLT(n)
inertia=(C.c_int*3)(); rc=C.c_double()
for it in range(3):
    torch.cuda.synchronize(); t0=time.perf_counter()
    _lib.check(F.lib.b200ipm_ldlt_factor(F.h, C.c_void_p(K.data_ptr()), n, 1, inertia, C.byref(rc)))
    torch.cuda.synchronize(); print('native single-GPU ldlt_factor n=16384 (incl. 2 device copies of 2 GB): %.1f ms' % ((time.perf_counter()-t0)*1e3), list(inertia))
PY
