"""Device-side timeline of ONE factorisation at config 3 (b200ipm_trace_*): where does the serial chain spend its time?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyipm_b200 import _lib, problems

prob = problems.make_nlp()
eng = _lib.Engine(prob.nvar, prob.neq, prob.nineq, _lib.default_params()); eng.bind(prob)
eng.set_state(prob.x0, np.ones(prob.nineq), np.zeros(prob.neq + prob.nineq), 0.2, 10.0, 0.0); eng.set_mu_host(0.2)
eng.init_slack(); eng.init_lambda()
for _ in range(3):
    eng.newton_step()
eng.state_save()
which = int(sys.argv[1]) if len(sys.argv) > 1 else 3
if which == 3:
    eng.profile_kernel(3, reps=2)
    _lib.trace_start()
    ms, _ = eng.profile_kernel(3, reps=1)
else:
    eng.state_restore(); eng.newton_step()
    eng.state_restore()
    _lib.trace_start()
    info = eng.newton_step(); ms = info.ms_factor
ids, blk, t0, t1 = _lib.trace_dump()
names = {1: 'tile', 2: 'panel', 3: 'mini', 4: 'upd64', 5: 'dmma'}
base = t0.min()
print('factor ms %.3f, %d records, span %.3f ms' % (ms, len(ids), (t1.max() - base) / 1e6))
for k in names:
    m = ids == k
    if m.any():
        d = (t1[m] - t0[m]) / 1e3
        print('%-6s n=%4d  dur us: mean %.1f  median %.1f  min %.1f  max %.1f   busy %.3f ms' % (names[k], m.sum(), d.mean(), np.median(d), d.min(), d.max(), d.sum() / 1e3))
# chain: tile kernels in start order
m = ids == 1
order = np.argsort(t0[m])
ts, te = t0[m][order], t1[m][order]
per = np.diff(ts) / 1e3
print('tile start-to-start period us: mean %.1f median %.1f ; gap (prev tile end -> next tile start) mean %.1f median %.1f' %
      (per.mean(), np.median(per), ((ts[1:] - te[:-1]) / 1e3).mean(), np.median((ts[1:] - te[:-1]) / 1e3)))
print('first 24 tile periods:', np.round(per[:24], 1))
print('last 16 tile periods:', np.round(per[-16:], 1))
# print the first ~40 events in time order
o = np.argsort(t0)
for i in o[:60]:
    print('%8.1f %8.1f  %-6s blk %d' % ((t0[i] - base) / 1e3, (t1[i] - base) / 1e3, names.get(int(ids[i]), '?'), blk[i]))
