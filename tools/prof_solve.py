"""Time the triangular solve (profile_kernel 4) and the factorisation (3) at config 3."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyipm_b200 import _lib, problems
prob = problems.make_nlp()
eng = _lib.Engine(prob.nvar, prob.neq, prob.nineq, _lib.default_params())
eng.bind(prob)
eng.set_state(prob.x0, np.ones(prob.nineq), np.zeros(prob.neq + prob.nineq), 0.2, 10.0, 1.0)
eng.init_slack(); eng.init_lambda()
for which, name in ((4, 'solve'), (3, 'factor')):
    ms, wk = eng.profile_kernel(which, reps=10)
    print(name, 'ms', ms, 'env grouped', os.environ.get('B200IPM_SOLVE_GROUPED'))
