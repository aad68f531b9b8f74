"""Pins the L-BFGS restatement of the oracle (oracle/pyipm_numpy.py: lbfgs_init / lbfgs_dir / lbfgs_update and the hooks
in newton_step/solve) against traces of the UNMODIFIED reference run with lbfgs=4 (tests/golden/make_golden.py lbfgs;
pyipm.py:993-1371, 1702-1713; unit_tests.py:49)."""
import numpy as np
import pytest

from oracle.pyipm_numpy import OracleIPM
from tests.util import EXAMPLES, get_problem, load_golden

LBFGS_CASES = EXAMPLES + ['qp_small', 'nlp_small', 'nlp_eqonly']


def lbfgs_callables(prob):
    return {k: v for k, v in prob.callables().items() if not k.startswith('d2')}


@pytest.mark.parametrize('name', LBFGS_CASES)
def test_oracle_lbfgs_matches_reference_trajectory(name):
    g = load_golden('lbfgs_' + name)
    prob, x0, gts = get_problem(name)
    tr = []
    o = OracleIPM(x0=np.array(x0), Ftol=1.0E-8, verbosity=-1, lbfgs=4, trace=tr, **lbfgs_callables(prob))
    with np.errstate(all='ignore'):
        x, s, lda, fval, kkt = o.solve()
    assert len(tr) == int(g['nsteps'])
    assert o.signal == int(g['sol0_signal'])
    np.testing.assert_allclose(x, g['sol0_x'], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(lda, g['sol0_lda'], rtol=1e-8, atol=1e-10)
    for k, st in enumerate(tr):
        np.testing.assert_allclose(st['x'], g['st_x'][k], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(st['g'], g['st_g'][k], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(st['dz'], g['st_dz'][k], rtol=1e-7, atol=1e-9)
        assert st['reg']['lbfgs_m'] == int(g['st_m'][k])
        np.testing.assert_allclose(st['reg']['zeta'], g['st_zeta'][k], rtol=1e-9)
        np.testing.assert_allclose(st['x_new'], g['st_x_new'][k], rtol=1e-9, atol=1e-11)
    if gts is not None:
        assert min(np.linalg.norm(x - gt) for gt in gts) <= 1.0E-3     # unit_tests.py:51,405-409
