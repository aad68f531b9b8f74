"""Direct parity tests of every operator slot of the C ABI (include/b200ipm.h) and of the host-side mirrors in
``pyipm_b200.IPM`` against the reference-generated fixtures / the CPU oracle: ``con``/``jaco`` (pyipm.py:564-607),
``KKT`` 4-tuple (958-991), merit ``phi``/``dphi`` values (670-721), ``step`` (1408-1436), the barrier update
(1804-1814), ``cost`` (855-857), and the ``rcond <= eps`` branch of ``reghess`` (1381-1389)."""
import numpy as np
import pytest

from oracle.pyipm_numpy import OracleIPM
from pyipm_b200 import IPM, _lib, problems
from tests.util import ALL_GOLDEN, get_problem, load_golden

pytestmark = pytest.mark.gpu


def relinf(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300)


def make_engine(prob, **kw):
    eng = _lib.Engine(prob.nvar, prob.neq, prob.nineq, _lib.default_params(**kw))
    eng.bind(prob)
    return eng


def compiled_oracle(prob, x0, **kw):
    o = OracleIPM(x0=np.array(x0), verbosity=-1, **prob.callables(), **kw)
    o.nvar = prob.nvar
    o.compile()
    return o


@pytest.mark.parametrize('name', ALL_GOLDEN)
def test_con_jac_cost_kkt_slots(name):
    """b200ipm_con_jac / b200ipm_cost / b200ipm_kkt at every state of the reference run."""
    g = load_golden(name)
    prob, x0, _ = get_problem(name)
    D, M, N = prob.nvar, prob.neq, prob.nineq
    o = compiled_oracle(prob, x0)
    eng = make_engine(prob)
    for k in range(int(g['nsteps'])):
        x, s, lda = g['st_x'][k], g['st_s'][k], g['st_lda'][k]
        o.mu_dev = np.float64(g['st_mu'][k])
        eng.set_state(x, s, lda, g['st_mu'][k], g['st_nu_before'][k], g['st_delta_before'][k])
        assert abs(eng.cost() - float(prob.f(x))) <= 1e-13 * (1.0 + abs(float(prob.f(x))))
        if M + N:
            con, J = eng.con_jac()
            cref = o.con(x, s)
            assert np.max(np.abs(con - cref)) <= 1e-13 * (1.0 + np.max(np.abs(cref)))
            Jref = o.jaco(x)[:D, :]
            assert np.max(np.abs(J - Jref)) <= 1e-13 * (1.0 + np.max(np.abs(Jref)))
        k1, k2, k3, k4 = eng.kkt()
        r1, r2, r3, r4 = o.KKT(x, s, lda)
        scale = 1.0 + np.max(np.abs(prob.df(x)))       # kkt1 itself goes to zero at convergence
        assert np.max(np.abs(k1 - r1)) <= 1e-12 * scale
        if N:
            assert np.max(np.abs(k2 - r2)) <= 1e-12 * (1.0 + np.max(np.abs(r2)))
            assert np.max(np.abs(k4 - r4)) <= 1e-12 * (1.0 + np.max(np.abs(r4)))
        if M:
            assert np.max(np.abs(k3 - r3)) <= 1e-12 * (1.0 + np.max(np.abs(r3)))
    eng.close()


@pytest.mark.parametrize('name', ['example1', 'example3', 'example5', 'example7', 'nlp_small'])
def test_ipm_class_slots_shapes_and_values(name):
    """The host-side mirrors IPM.con / IPM.jaco (full (D+N) x (M+N) shape, pyipm.py:581-607) / IPM.KKT (scalar 0.0 for
    absent conditions, pyipm.py:977-989) / IPM.grad / IPM.hess / IPM.cost against the oracle."""
    g = load_golden(name)
    prob, x0, _ = get_problem(name)
    D, M, N = prob.nvar, prob.neq, prob.nineq
    o = compiled_oracle(prob, x0)
    p = IPM(x0=np.array(x0), f=prob, verbosity=-1)
    p.nvar = D
    p.compile()
    k = min(1, int(g['nsteps']) - 1)
    x, s, lda = g['st_x'][k], g['st_s'][k], g['st_lda'][k]
    p.mu_dev = float(g['st_mu'][k])
    o.mu_dev = np.float64(g['st_mu'][k])
    assert abs(p.cost(x) - float(o.cost(x))) <= 1e-13 * (1 + abs(float(o.cost(x))))
    assert relinf(p.grad(x, s, lda), o.grad(x, s, lda)) < 1e-12
    Href = o.hess(x, s, lda)
    assert np.max(np.abs(p.hess(x, s, lda) - Href)) <= 1e-12 * max(1.0, np.max(np.abs(Href)))
    if M + N:
        np.testing.assert_allclose(p.con(x, s), o.con(x, s), rtol=1e-12, atol=1e-13)
        Jg, Jr = p.jaco(x), o.jaco(x)
        assert Jg.shape == Jr.shape == ((D + N, M + N) if N else (D, M))
        np.testing.assert_allclose(Jg, Jr, rtol=1e-12, atol=1e-13)
    kg, kr = p.KKT(x, s, lda), o.KKT(x, s, lda)
    assert isinstance(kg, tuple) and len(kg) == 4
    for a, b in zip(kg, kr):
        assert np.ndim(a) == np.ndim(b)          # scalar 0.0 where the reference returns scalar 0.0
        if np.ndim(b) == 0:
            assert a == 0.0 and b == 0.0
        else:
            np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize('name', ALL_GOLDEN)
def test_merit_values_and_step_max(name):
    """a7/a8: phi(x0, s0) and dphi(x0, s0, dz) as VALUES (info.phi0 / info.dphi0 of b200ipm_newton_step, and the
    b200ipm_merit slot after b200ipm_direction) and the fraction-to-the-boundary limits of b200ipm_step_max, against
    the oracle's search() record at every step of the oracle trajectory."""
    prob, x0, _ = get_problem(name)
    tr = []
    o = OracleIPM(x0=np.array(x0), Ftol=1.0E-8, verbosity=-1, trace=tr, **prob.callables())
    with np.errstate(all='ignore'):
        o.solve()
    g = load_golden(name)
    eng = make_engine(prob)
    for k, st in enumerate(tr):
        sr = st['search']
        tol = 1e-9 * (1.0 + abs(sr['phi0']))
        eng.set_state(st['x'], st['s'], st['lda'], st['mu'], g['st_nu_before'][k], g['st_delta_before'][k])
        eng.set_mu_host(st['mu_host'])
        info = eng.newton_step()
        assert abs(info.phi0 - sr['phi0']) <= tol, (k, info.phi0, sr['phi0'])
        assert abs(info.dphi0 - sr['dphi0']) <= 1e-8 * (1.0 + abs(sr['dphi0'])), (k, info.dphi0, sr['dphi0'])
        assert abs(info.alpha_s - sr['alpha_s']) <= 1e-12 and abs(info.alpha_l - sr['alpha_l']) <= 1e-12
        # the stand-alone slots: direction, then merit / step_max at the same state with the POST-update nu
        eng.set_state(st['x'], st['s'], st['lda'], st['mu'], st['nu_after'], g['st_delta_before'][k])
        eng.direction(want_dz=False)
        phi, dphi = eng.merit()
        assert abs(phi - sr['phi0']) <= tol
        assert abs(dphi - sr['dphi0']) <= 1e-8 * (1.0 + abs(sr['dphi0']))
        if prob.nineq:
            a_s, a_l = eng.step_max()
            assert abs(a_s - st['alpha_smax']) < 1e-12 and abs(a_l - st['alpha_lmax']) < 1e-12
    eng.close()


@pytest.mark.parametrize('name', ['example5', 'example6', 'example7', 'example9', 'example10', 'qp_mid', 'nlp_mid'])
def test_update_mu_slot(name):
    """a11: b200ipm_update_mu against the formula of pyipm.py:1804-1814 at every accepted point of the reference run,
    and against the reference's own next barrier parameter wherever it changed between consecutive steps."""
    g = load_golden(name)
    prob, x0, _ = get_problem(name)
    M, N = prob.neq, prob.nineq
    eps = np.finfo(np.float64).eps
    eng = make_engine(prob)
    n = int(g['nsteps'])
    nchg = 0
    for k in range(n):
        s, lda = g['st_s_new'][k], g['st_lda_new'][k]
        eng.set_state(g['st_x_new'][k], s, lda, g['st_mu'][k], 10.0, 0.0)
        mu = eng.update_mu()
        xi = N * np.min(s * lda[M:]) / (np.dot(s, lda[M:]) + eps)
        ref = max(0.1 * np.min([0.05 * (1.0 - xi) / (xi + eps), 2.0]) ** 3 * np.dot(s, lda[M:]) / N, 0.0)
        assert abs(mu - ref) <= 1e-12 * max(ref, 1e-300), (k, mu, ref)
        if k + 1 < n and k + 1 != int(g['sol0_nsteps']) and g['st_mu_host'][k + 1] != g['st_mu_host'][k]:
            assert abs(mu - g['st_mu_host'][k + 1]) <= 1e-12 * g['st_mu_host'][k + 1]
            nchg += 1
    assert nchg >= 1
    eng.close()


@pytest.mark.parametrize('flags', [0, 1, 64, 6])
def test_reghess_rcond_branch_rank_deficient_jacobian(flags):
    """pyipm.py:1381-1389: with a rank-deficient equality Jacobian the unshifted matrix is singular (rcond <= eps), the
    reference regularises the eq-multiplier block by -sqrt(eps)*eta*mu^beta and then shifts.  Same eq_reg / delta /
    number of inertia tests on the sequential (flags 1), background-test (64), certificate (0) and tcgen05 (6) paths
    as the reference-generated fixture; the direction solves the regularised system."""
    g = load_golden('nlp_rankdef')
    prob = problems.make_rankdef_nlp()
    D, M, N = prob.nvar, prob.neq, prob.nineq
    o = compiled_oracle(prob, prob.x0)
    eng = make_engine(prob, flags=flags)
    for k in range(int(g['nsteps'])):
        x, s, lda = g['st_x'][k], g['st_s'][k], g['st_lda'][k]
        eng.set_state(x, s, lda, g['st_mu'][k], g['st_nu_before'][k], g['st_delta_before'][k])
        eng.set_mu_host(g['st_mu_host'][k])
        dz, info = eng.direction()
        assert info.eq_reg == 1, (k, info.asdict())
        assert info.delta == g['st_delta'][k], (k, info.delta, g['st_delta'][k])
        assert info.n_factor == int(g['st_n_eig'][k]), (k, info.n_factor, int(g['st_n_eig'][k]))
        assert info.n_neg == M and info.n_zero == 0
        # the direction solves the reference's regularised matrix (st_Hreg) to a residual at rounding level; the system
        # itself has condition ~1e13 (eq block = -7.8e-13 on a null direction), so dz is compared through the residual
        # and, loosely, with the reference's LU solution
        Hreg = g['st_Hreg'][k]
        y = dz.copy()
        y[D + N:] = -y[D + N:]
        r = Hreg @ y - g['st_g'][k]
        assert np.max(np.abs(r)) <= 1e-9 * max(1.0, np.max(np.abs(g['st_g'][k]))), (k, np.max(np.abs(r)))
        assert relinf(dz[:D + N], g['st_dz'][k][:D + N]) < 1e-3, (k, relinf(dz[:D + N], g['st_dz'][k][:D + N]))
    eng.close()


def test_rank_deficient_jacobian_at_tcgen05_size():
    """The same branch on a problem large enough for the default tcgen05 path and several LDL^T panels (D = 320): the
    decisions of the CPU oracle, step by step from delta = 0 and with speculation."""
    prob = problems.make_rankdef_nlp(D=320, M=48, N=320, seed=42)
    tr = []
    o = OracleIPM(x0=prob.x0.copy(), Ftol=1.0E-8, verbosity=-1, niter=1, miter=3, trace=tr, **prob.callables())
    with np.errstate(all='ignore'):
        o.solve()
    for flags in (6, 6 | 64, 1):
        eng = make_engine(prob, flags=flags)
        nu_b, de_b = 10.0, 0.0
        for k, st in enumerate(tr):
            assert st['reg']['eq_reg']
            eng.set_state(st['x'], st['s'], st['lda'], st['mu'], nu_b, de_b)
            eng.set_mu_host(st['mu_host'])
            dz, info = eng.direction()
            assert info.eq_reg == 1 and info.delta == st['delta'] and info.n_factor == st['reg']['n_eig'], \
                (flags, k, info.asdict(), st['reg'])
            nu_b, de_b = st['nu_after'], st['delta']
        eng.close()
