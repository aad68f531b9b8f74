"""Device-side timeline of the factorisation(s) at config 3 (b200ipm_trace_*): where does the serial chain spend its
time?   python tools/trace_factor.py 3 -> one factorisation alone;   ... 0 -> the reghess of a full Newton step
(foreground candidate + background delta = 0 test), reported per factorisation."""
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
    for _ in range(2):
        eng.state_restore(); eng.newton_step()
    eng.state_restore()
    _lib.trace_start()
    info = eng.newton_step(); ms = info.ms_factor
ids, blk, t0, t1, tag = _lib.trace_dump()
names = {1: 'tile', 2: 'panel', 3: 'mini', 4: 'upd64', 5: 'dmma', 6: 'oz'}
base = t0.min()
print('factor ms %.3f, %d records, span %.3f ms' % (ms, len(ids), (t1.max() - base) / 1e6))
for tg in np.unique(tag[ids == 1]):
    sel = tag == tg
    tm = sel & (ids == 1)
    print('== factorisation with control block %x: first tile at %.1f us, last tile ends %.1f us' %
          (int(tg), (t0[tm].min() - base) / 1e3, (t1[tm].max() - base) / 1e3))
    for k in names:
        m = sel & (ids == k)
        if m.any():
            d = (t1[m] - t0[m]) / 1e3
            print('   %-6s n=%4d  dur us: mean %.1f  median %.1f  min %.1f  max %.1f   busy %.3f ms' %
                  (names[k], m.sum(), d.mean(), np.median(d), d.min(), d.max(), d.sum() / 1e3))
    order = np.argsort(t0[tm])
    ts, te = t0[tm][order], t1[tm][order]
    per = np.diff(ts) / 1e3
    print('   tile start-to-start period us: mean %.1f median %.1f ; gap (tile end -> next tile start) mean %.1f median %.1f' %
          (per.mean(), np.median(per), ((ts[1:] - te[:-1]) / 1e3).mean(), np.median((ts[1:] - te[:-1]) / 1e3)))
    print('   periods:', ' '.join('%.0f' % v for v in per))
if len(sys.argv) > 2:
    o = np.argsort(t0)
    for i in o[:int(sys.argv[2])]:
        print('%8.1f %8.1f  %-6s blk %d tag %x' % ((t0[i] - base) / 1e3, (t1[i] - base) / 1e3, names.get(int(ids[i]), '?'), blk[i], int(tag[i]) & 0xfffff))
if len(sys.argv) > 4:
    lo, hi = float(sys.argv[3]), float(sys.argv[4])
    o = np.argsort(t0)
    for i in o:
        a, b = (t0[i] - base) / 1e3, (t1[i] - base) / 1e3
        if a >= lo and a <= hi:
            print('%8.1f %8.1f  %-6s blk %d' % (a, b, names.get(int(ids[i]), '?'), blk[i]))
