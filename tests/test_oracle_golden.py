"""Pins the CPU oracle (oracle/pyipm_numpy.py) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py) and against the reference's own known answers
(unit_tests.py:104-235 ground truths, tolerance Stol=1e-3 at unit_tests.py:51,405-409; README.md:115-121)."""
import numpy as np
import pytest

from oracle.pyipm_numpy import OracleIPM
from tests.util import ALL_GOLDEN, EXAMPLES, get_problem, load_golden

# Same NumPy/SciPy calls in the same order => we demand agreement to a few ulps, not just "close".
RTOL = 1e-12
ATOL = 1e-13


def run_oracle(name, n_solves=1):
    prob, x0, gts = get_problem(name)
    trace = []
    o = OracleIPM(x0=np.array(x0), Ftol=1.0E-8, verbosity=-1, trace=trace, **prob.callables())
    sols = [o.solve() for _ in range(n_solves)]
    return o, trace, sols, gts


@pytest.mark.parametrize('name', ALL_GOLDEN)
def test_oracle_matches_reference_trajectory(name):
    g = load_golden(name)
    n_solves = 2 if name == 'example7' else 1
    o, trace, sols, _ = run_oracle(name, n_solves)
    assert len(trace) == int(g['nsteps'])
    for i, (x, s, lda, fval, kkt) in enumerate(sols):
        pre = 'sol%d_' % i
        np.testing.assert_allclose(x, g[pre + 'x'], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(s, g[pre + 's'], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(lda, g[pre + 'lda'], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(fval, g[pre + 'fval'], rtol=RTOL, atol=ATOL)
        for j in range(4):
            np.testing.assert_allclose(np.atleast_1d(kkt[j]), g[pre + 'kkt%d' % (j + 1)], rtol=1e-9, atol=1e-12)
    assert o.signal == int(g['sol%d_signal' % (n_solves - 1)])
    # per-Newton-step quantities (teacher-forcing data for the GPU tests)
    for k, st in enumerate(trace):
        np.testing.assert_allclose(st['x'], g['st_x'][k], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(st['s'], g['st_s'][k], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(st['lda'], g['st_lda'][k], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(st['g'], g['st_g'][k], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(st['dz'], g['st_dz'][k], rtol=1e-9, atol=1e-12)
        assert st['mu'] == g['st_mu'][k]
        assert st['mu_host'] == g['st_mu_host'][k]
        np.testing.assert_allclose(st['nu_after'], g['st_nu'][k], rtol=RTOL)
        assert st['delta'] == g['st_delta'][k]
        assert st['reg']['n_eig'] == int(g['st_n_eig'][k])
        assert st['reg']['nneg0'] == int(g['st_nneg_first'][k])
        if 'alpha_smax' in st:
            np.testing.assert_allclose(st['alpha_smax'], g['st_alpha_smax'][k], rtol=0, atol=1e-15)
            np.testing.assert_allclose(st['alpha_lmax'], g['st_alpha_lmax'][k], rtol=0, atol=1e-15)
        np.testing.assert_allclose(st['x_new'], g['st_x_new'][k], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(st['lda_new'], g['st_lda_new'][k], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize('name', EXAMPLES)
def test_oracle_reaches_reference_ground_truth(name):
    """The reference's own pass criterion: ||x_gt - x||_2 <= 1e-3 for any listed minimiser."""
    o, trace, sols, gts = run_oracle(name)
    x = sols[0][0]
    assert min(np.linalg.norm(x - gt) for gt in gts) <= 1.0E-3
    assert o.signal in (1, 2)


def test_oracle_example7_readme_transcript():
    """README.md:115-121: x ~ 1/3, f = 0.0370370370369, lda ~ (-0.1111, ~0, ~0, ~0), KKT <= ~1.2e-6 scale."""
    o, trace, sols, _ = run_oracle('example7')
    x, s, lda, fval, kkt = sols[0]
    assert np.allclose(x, 1.0 / 3.0, atol=1e-3) and np.allclose(s, 1.0 / 3.0, atol=1e-3)
    assert abs(-fval - 0.0370370370369) < 1e-6
    assert abs(lda[0] + 1.0 / 9.0) < 1e-3 and np.all(np.abs(lda[1:]) < 1e-3)
    assert all(np.linalg.norm(k) <= 1e-4 for k in kkt)


def test_oracle_full_kkt_matrix_matches_reference():
    """hess(): the reference's full K x K matrix before/after reghess (pyipm.py:768-814, 1373-1406)."""
    for name in ('example7', 'example4', 'nlp_small'):
        g = load_golden(name)
        prob, x0, _ = get_problem(name)
        o = OracleIPM(x0=np.array(x0), Ftol=1.0E-8, verbosity=-1, **prob.callables())
        o.nvar = prob.nvar
        o.compile()
        for k in range(int(g['nsteps'])):
            o.mu_dev = o.float_dtype(g['st_mu'][k])
            H = o.hess(g['st_x'][k], g['st_s'][k], g['st_lda'][k])
            np.testing.assert_allclose(H, g['st_Hfull'][k], rtol=RTOL, atol=ATOL)
            assert np.array_equal(H, H.T)


def test_oracle_rcond_branch_matches_reference():
    """The `rcond <= eps` branch of reghess (pyipm.py:1381-1389: eq-block regularisation) on the rank-deficient
    fixture: 4 Newton steps of the unmodified reference (niter=1, miter=4), every one with the eq block regularised."""
    from pyipm_b200 import problems
    g = load_golden('nlp_rankdef')
    prob = problems.make_rankdef_nlp()
    tr = []
    o = OracleIPM(x0=prob.x0.copy(), Ftol=1.0E-8, verbosity=-1, niter=1, miter=4, trace=tr, **prob.callables())
    with np.errstate(all='ignore'):
        o.solve()
    assert len(tr) == int(g['nsteps']) == 4
    for k, st in enumerate(tr):
        assert st['reg']['eq_reg'] and st['reg']['rcond'] <= o.eps
        assert g['st_rcond'][k] <= o.eps
        assert st['delta'] == g['st_delta'][k]
        assert st['reg']['n_eig'] == int(g['st_n_eig'][k])
        np.testing.assert_allclose(st['g'], g['st_g'][k], rtol=1e-9, atol=1e-11)
        # the matrix reghess returns: eq block = -sqrt(eps)*eta*mu^beta, (x,x) block shifted by delta
        D, M, N = prob.nvar, prob.neq, prob.nineq
        Hreg = g['st_Hreg'][k]
        reg = np.sqrt(o.eps) * o.eta * g['st_mu_host'][k] ** o.beta
        np.testing.assert_allclose(np.diagonal(Hreg)[D + N:D + N + M], -reg, rtol=1e-12)
        # the system is singular to working precision (condition ~1e13 after the regularisation): the LU solution is
        # reproducible only to that level
        np.testing.assert_allclose(st['dz'], g['st_dz'][k], rtol=1e-4, atol=1e-6 * np.max(np.abs(g['st_dz'][k])))
