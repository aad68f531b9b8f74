"""L-BFGS mode (pyipm.py:993-1371, hooks 1702-1713) on the device, through the drop-in class, against
traces of the UNMODIFIED reference run with lbfgs=4 (tests/golden/ref_lbfgs_*.npz, unit_tests.py:49) and against the
CPU oracle on larger problems."""
import numpy as np
import pytest

from oracle.pyipm_numpy import OracleIPM
from pyipm_b200 import IPM, problems
from tests.util import EXAMPLES, get_problem, load_golden

pytestmark = pytest.mark.gpu

LBFGS_CASES = EXAMPLES + ['qp_small', 'nlp_small', 'nlp_eqonly']


def lbfgs_callables(prob):
    return {k: v for k, v in prob.callables().items() if not k.startswith('d2')}


def check_against_trace(p, x, s, lda, g_nsteps, g_signal, g_x, g_lda, m_trace, zeta_trace, x_new_trace):
    assert p.signal == g_signal
    assert p.iter_count == g_nsteps
    for k, lg in enumerate(p.step_log):
        assert lg['lbfgs_m'] == int(m_trace[k]), (k, lg['lbfgs_m'], int(m_trace[k]))
        assert abs(lg['lbfgs_zeta'] - zeta_trace[k]) <= 1e-6 * abs(zeta_trace[k]), (k, lg['lbfgs_zeta'], zeta_trace[k])
    assert np.linalg.norm(x - g_x) <= 1e-6 * (1.0 + np.linalg.norm(g_x))
    if g_lda.size:
        assert np.linalg.norm(lda - g_lda) <= 1e-5 * (1.0 + np.linalg.norm(g_lda))


@pytest.mark.parametrize('name', LBFGS_CASES)
def test_lbfgs_full_solve_matches_reference(name):
    """Lowered problems: every inner iteration is b200ipm_lbfgs_step.  Same signal, same iteration count, same number of
    stored pairs and the same zeta at every step, final (x, lda) as the reference's."""
    g = load_golden('lbfgs_' + name)
    prob, x0, gts = get_problem(name)
    p = IPM(x0=np.array(x0), f=prob, Ftol=1.0E-8, lbfgs=4, verbosity=-1)
    x, s, lda, fval, kkt = p.solve()
    check_against_trace(p, x, s, lda, int(g['sol0_nsteps']), int(g['sol0_signal']), g['sol0_x'], g['sol0_lda'], g['st_m'],
                        g['st_zeta'], g['st_x_new'])
    if gts is not None:
        assert min(np.linalg.norm(x - gt) for gt in gts) <= 1.0E-3        # unit_tests.py:51,405-409


@pytest.mark.parametrize('name', ['example1', 'example4', 'example7', 'example10', 'example5'])
def test_lbfgs_callable_mode_matches_reference(name):
    """Opaque host callables without second derivatives (the reference's precompiled-function mode with lbfgs=4):
    b200ipm_lbfgs_update(gradx_old) + b200ipm_lbfgs_direction per step."""
    g = load_golden('lbfgs_' + name)
    prob, x0, _ = get_problem(name)
    p = IPM(x0=np.array(x0), Ftol=1.0E-8, lbfgs=4, verbosity=-1, **lbfgs_callables(prob))
    x, s, lda, fval, kkt = p.solve()
    assert p.signal == int(g['sol0_signal'])
    assert np.linalg.norm(x - g['sol0_x']) <= 1e-5 * (1.0 + np.linalg.norm(g['sol0_x']))


def test_lbfgs_direction_teacher_forced_midsize():
    """A problem large enough for the DMMA Schur-complement product and a multi-panel factorisation (D = 96, M = 12,
    N = 96): the whole trajectory against the CPU oracle, and every direction against the oracle's at 1e-7."""
    prob = problems.make_nlp(D=96, M=12, N=96, seed=51)
    tr = []
    o = OracleIPM(x0=prob.x0.copy(), Ftol=1.0E-8, verbosity=-1, lbfgs=4, trace=tr, niter=3, miter=8,
                  **lbfgs_callables(prob))
    with np.errstate(all='ignore'):
        xo, so, lo, fo, ko = o.solve()
    p = IPM(x0=prob.x0.copy(), f=prob, Ftol=1.0E-8, lbfgs=4, verbosity=-1, niter=3, miter=8)
    x, s, lda, fval, kkt = p.solve()
    assert p.iter_count == len(tr) and p.signal == o.signal
    for k, (st, lg) in enumerate(zip(tr, p.step_log)):
        assert lg['lbfgs_m'] == st['reg']['lbfgs_m'], k
        assert abs(lg['lbfgs_zeta'] - st['reg']['zeta']) <= 1e-6 * abs(st['reg']['zeta']), k
        assert lg['n_backtracks'] == st['search']['n_backtracks'], k
    assert np.linalg.norm(x - xo) <= 1e-6 * (1.0 + np.linalg.norm(xo))
    assert np.linalg.norm(lda - lo) <= 1e-5 * (1.0 + np.linalg.norm(lo))
