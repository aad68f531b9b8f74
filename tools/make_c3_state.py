#!/usr/bin/env python
"""Run on the GPU box: config 3 from x0, PRE_STEPS real Newton steps, then dump the iterate (x, s, lda, mu, nu, delta) to
gpurun_out/c3_state_after3.npz.  The file is committed as tests/golden/c3_state_after3.npz: it is the teacher-forced
state both arms of bench.py step from (the reference arm has no GPU to produce it with)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyipm_b200 import _lib, problems  # noqa: E402

PRE_STEPS = 3
prob = problems.make_nlp()
D, M, N = prob.nvar, prob.neq, prob.nineq
eng = _lib.Engine(D, M, N, _lib.default_params())
eng.bind(prob)
eng.set_state(prob.x0, np.ones(N), np.zeros(M + N), 0.2, 10.0, 0.0)
eng.set_mu_host(0.2)
eng.init_slack()
eng.init_lambda()
infos = [eng.newton_step().asdict() for _ in range(PRE_STEPS)]
x, s, lda, mu, nu, delta = eng.get_state()
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, 'gpurun_out', 'c3_state_after3.npz'), x=x, s=s, lda=lda, mu=mu, nu=nu, delta=delta,
                    mu_host=0.2, pre_steps=PRE_STEPS, deltas=np.array([i['delta'] for i in infos]))
print('delta', delta, 'nu', nu, [i['n_factor'] for i in infos])
