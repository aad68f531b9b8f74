"""Sweep the SM budget of the bulk trailing updates (foreground / background LDL^T) at config 3."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyipm_b200 import _lib, problems

prob = problems.make_nlp()
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for fg, bg in [(96, 48), (0, 48), (132, 48), (120, 48), (108, 48), (84, 48), (72, 48)]:
    os.environ['B200IPM_SIDE_FG'] = str(fg)
    os.environ['B200IPM_SIDE_BG'] = str(bg)
    eng = _lib.Engine(prob.nvar, prob.neq, prob.nineq, _lib.default_params(flags=flags))
    eng.bind(prob)
    eng.set_state(prob.x0, np.ones(prob.nineq), np.zeros(prob.neq + prob.nineq), 0.2, 10.0, 0.0)
    eng.set_mu_host(0.2)
    eng.init_slack(); eng.init_lambda()
    for _ in range(3):
        eng.newton_step()
    eng.state_save()
    for _ in range(3):
        eng.state_restore(); eng.newton_step()
    ts, tf = [], []
    t0 = time.perf_counter()
    for _ in range(10):
        eng.state_restore()
        i = eng.newton_step()
        ts.append(i.ms_total); tf.append(i.ms_factor)
    wall = (time.perf_counter() - t0) / 10 * 1e3
    print('fg=%3d bg=%3d  ms_total %.3f  ms_factor %.3f  wall/step %.3f ms  n_factor %d' % (fg, bg, np.mean(ts), np.mean(tf), wall, i.n_factor), flush=True)
    eng.close()
