"""Batched multi-start solver (b200ipm_batch_solve_poly: one warp per instance, the whole
IPM.solve() loop of pyipm.py:1567-1863 on the device) against the reference-generated fixtures (the seed-42 starting
point of every example problem) and against the CPU oracle run instance by instance on random starting points."""
import numpy as np
import pytest

from oracle.pyipm_numpy import OracleIPM
from pyipm_b200 import problems
from pyipm_b200.batch import solve_batch
from tests.util import load_golden

pytestmark = pytest.mark.gpu


def starts(k, n, seed):
    rng = np.random.RandomState(seed)
    prob, _ = problems.example_problem(k)
    X = rng.rand(n, prob.nvar) if k == 6 else rng.randn(n, prob.nvar)
    X[0] = problems.example_x0(k)
    return X


@pytest.mark.parametrize('k', range(1, 11))
def test_batch_matches_reference_and_oracle(k):
    prob, gts = problems.example_problem(k)
    g = load_golden('example%d' % k)
    X0 = starts(k, 48, 100 + k)
    res = solve_batch(prob, X0, Ftol=1.0E-8)
    # instance 0 = the reference run of the fixture: same signal, same number of Newton steps, same solution
    assert res.signal[0] == int(g['sol0_signal'])
    assert res.iters[0] == int(g['sol0_nsteps'])
    assert np.linalg.norm(res.x[0] - g['sol0_x']) <= 1e-6 * (1.0 + np.linalg.norm(g['sol0_x']))
    if prob.neq + prob.nineq:
        assert np.linalg.norm(res.lda[0] - g['sol0_lda']) <= 1e-5 * (1.0 + np.linalg.norm(g['sol0_lda']))
    assert min(np.linalg.norm(res.x[0] - gt) for gt in gts) <= 1.0E-3
    # every other instance against the CPU oracle from the same starting point
    same = close = 0
    for b in range(1, X0.shape[0]):
        o = OracleIPM(x0=X0[b].copy(), Ftol=1.0E-8, verbosity=-1, **prob.callables())
        try:
            with np.errstate(all='ignore'):
                xo, so, lo, fo, ko = o.solve()
        except Exception:
            continue          # the reference itself fails from this start (e.g. overflow in its delta loop)
        ok = (o.signal == res.signal[b]) and (o.iter_count == res.iters[b])
        same += ok
        if ok and np.linalg.norm(res.x[b] - xo) <= 1e-5 * (1.0 + np.linalg.norm(xo)):
            close += 1
    n = X0.shape[0] - 1
    assert same >= 0.85 * n, (k, same, n)          # identical signal AND iteration count
    assert close >= 0.85 * n, (k, close, n)


def test_large_batch_example7():
    """10^4 starting points of example 7 (BASELINE config 1) in one launch."""
    prob, gts = problems.example_problem(7)
    rng = np.random.RandomState(7)
    X0 = 0.3 + 0.2 * rng.rand(10000, 3)
    res = solve_batch(prob, X0, Ftol=1.0E-8)
    conv = np.isin(res.signal, (1, 2))
    err = np.linalg.norm(res.x - gts[0][None, :], axis=1)
    assert conv.mean() >= 0.99
    assert (err[conv] <= 1e-3).mean() >= 0.99
    # a sample against the oracle: same solution everywhere; signal and iteration count may differ in borderline cases
    # (Ktol vs Ftol convergence decided by a last-digit difference), allowed for at most 10 % of the sample
    same = 0
    sample = rng.choice(10000, 40, replace=False)
    for b in sample:
        o = OracleIPM(x0=X0[b].copy(), Ftol=1.0E-8, verbosity=-1, **prob.callables())
        with np.errstate(all='ignore'):
            xo = o.solve()[0]
        same += int(o.signal == res.signal[b] and o.iter_count == res.iters[b])
        assert np.linalg.norm(res.x[b] - xo) <= 1e-4
    assert same >= 0.9 * len(sample), same
    print('10^4 solves of example 7: %.2f ms kernel time = %.0f solves/s' % (res.ms, 1e4 / (res.ms * 1e-3)))
