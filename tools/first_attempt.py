import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyipm_b200 import _lib, problems
prob = problems.make_nlp()
eng = _lib.Engine(prob.nvar, prob.neq, prob.nineq, _lib.default_params(flags=int(sys.argv[1]) if len(sys.argv) > 1 else None)); eng.bind(prob)
eng.set_state(prob.x0, np.ones(prob.nineq), np.zeros(prob.neq + prob.nineq), 0.2, 10.0, 0.0); eng.set_mu_host(0.2)
eng.init_slack(); eng.init_lambda()
for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6):
    info = eng.newton_step()
    print(i, 'n_factor', info.n_factor, 'first attempt neg/zero', info.n_neg_first, info.n_zero_first, 'abandoned', info.abandoned_first, 'cert', info.cert_used, 'spec', info.n_spec, info.spec_used, 'tc', info.tc_syrk, 'final neg', info.n_neg, 'delta %.3e' % info.delta,
          'ms factor %.2f solve %.2f total %.2f' % (info.ms_factor, info.ms_solve, info.ms_total), 'kkt', ['%.2e' % v for v in info.kkt_norm], 'bt', info.n_backtracks, 'soc', info.soc_tried, info.soc_accepted, 'search ms %.2f' % info.ms_search, 'alpha %.3f' % info.alpha_s)
