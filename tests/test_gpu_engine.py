"""Engine parity on the B200, through the C ABI, against the CPU oracle and the reference-generated fixtures.

Teacher-forced single Newton steps start from oracle states (tight tolerances: same inertia decision, same
delta, same number of backtracks); full solves are compared at 1e-6 and against the reference's known answers.
Tolerances follow SURVEY.md section 8(c)."""
import numpy as np
import pytest

from oracle.pyipm_numpy import OracleIPM
from pyipm_b200 import IPM, _lib, problems
from tests.util import ALL_GOLDEN, EXAMPLES, get_problem, load_golden

pytestmark = pytest.mark.gpu

DZ_RTOL = 1e-9


def oracle_trace(prob, x0, **kw):
    tr = []
    o = OracleIPM(x0=np.array(x0), Ftol=1.0E-8, verbosity=-1, trace=tr, **prob.callables(), **kw)
    with np.errstate(all='ignore'):
        o.solve()
    return o, tr


def make_engine(prob, **kw):
    eng = _lib.Engine(prob.nvar, prob.neq, prob.nineq, _lib.default_params(**kw))
    eng.bind(prob)
    return eng


def relinf(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize('name', ALL_GOLDEN)
def test_residual_and_full_kkt_matrix(name):
    """a1/a10/a3: grad(), KKT norms and the reference's full K x K hess() at every state of the reference run."""
    g = load_golden(name)
    prob, x0, _ = get_problem(name)
    eng = make_engine(prob)
    for k in range(int(g['nsteps'])):
        eng.set_state(g['st_x'][k], g['st_s'][k], g['st_lda'][k], g['st_mu'][k], g['st_nu_before'][k],
                      g['st_delta_before'][k])
        gv, nrm = eng.residual()
        ref = -g['st_g'][k]
        assert relinf(gv, ref) < 1e-12
        if 'st_Hfull' in g.files:
            H = eng.hess_full()
            Href = g['st_Hfull'][k]
            assert np.array_equal(H, H.T)
            assert np.max(np.abs(H - Href)) <= 1e-12 * max(1.0, np.max(np.abs(Href)))
    eng.close()


@pytest.mark.parametrize('name', ALL_GOLDEN)
def test_teacher_forced_newton_steps(name):
    """a3-a9: one device Newton step from each oracle state: dz, delta, #factorisations (= eigvalsh calls),
    step limits, number of backtracks, accepted point."""
    prob, x0, _ = get_problem(name)
    o, tr = oracle_trace(prob, x0)
    g = load_golden(name)
    eng = make_engine(prob)
    for k, st in enumerate(tr):
        eng.set_state(st['x'], st['s'], st['lda'], st['mu'], g['st_nu_before'][k], g['st_delta_before'][k])
        eng.set_mu_host(st['mu_host'])
        dz, dinfo = eng.direction()
        assert relinf(dz, st['dz']) < DZ_RTOL, (k, relinf(dz, st['dz']))
        assert dinfo.delta == st['delta'], (k, dinfo.delta, st['delta'])
        assert dinfo.n_factor == st['reg']['n_eig'], (k, dinfo.n_factor, st['reg']['n_eig'])
        assert dinfo.n_neg == prob.neq
        assert dinfo.resid < 1e-9 * max(1.0, np.max(np.abs(st['g'])))
        # the full step from the same state
        eng.set_state(st['x'], st['s'], st['lda'], st['mu'], g['st_nu_before'][k], g['st_delta_before'][k])
        info = eng.newton_step()
        x, s, lda, mu, nu, delta = eng.get_state()
        assert abs(nu - st['nu_after']) <= 1e-9 * abs(st['nu_after']), k
        if prob.nineq:
            assert abs(info.alpha_smax - st['alpha_smax']) < 1e-12
            assert abs(info.alpha_lmax - st['alpha_lmax']) < 1e-12
        sr = st['search']
        assert info.soc_tried == int(sr['soc_tried']), k
        assert info.soc_accepted == int(sr['soc_accepted']), k
        assert info.n_backtracks == sr['n_backtracks'], (k, info.n_backtracks, sr['n_backtracks'])
        tol = 1e-8 if not sr['soc_accepted'] else 1e-6
        assert relinf(x, st['x_new']) < tol, (k, relinf(x, st['x_new']))
        if prob.nineq:
            assert relinf(s, st['s_new']) < tol
        if prob.neq + prob.nineq:
            assert relinf(lda, st['lda_new']) < tol
        np.testing.assert_allclose(np.array(list(info.kkt_norm)), st['kkt_norms'], rtol=1e-6, atol=1e-10)
    eng.close()


@pytest.mark.parametrize('name', ALL_GOLDEN)
def test_full_solve_matches_reference(name):
    """IPM(...).solve() drop-in: final (x, s, lda, fval, kkt), signal and iteration count vs the reference run."""
    g = load_golden(name)
    prob, x0, gts = get_problem(name)
    p = IPM(x0=np.array(x0), f=prob, Ftol=1.0E-8, verbosity=-1)
    x, s, lda, fval, kkt = p.solve()
    assert p.signal == int(g['sol0_signal'])
    assert p.iter_count == int(g['sol0_nsteps'])
    for a, key in ((x, 'x'), (s, 's'), (lda, 'lda')):
        ref = g['sol0_' + key]
        assert np.linalg.norm(a - ref) <= 1e-6 * (1.0 + np.linalg.norm(ref)), key
    assert abs(fval - float(g['sol0_fval'])) <= 1e-8 * (1.0 + abs(float(g['sol0_fval'])))
    if gts is not None:
        assert min(np.linalg.norm(x - gt) for gt in gts) <= 1.0E-3    # unit_tests.py:51,405-409


def test_second_solve_quirk_mu_dev_not_reset():
    """Quirk xi (pyipm.py:1603 vs 1607): a second solve() starts from the stale mu_dev."""
    g = load_golden('example7')
    prob, x0, _ = get_problem('example7')
    p = IPM(x0=np.array(x0), f=prob, Ftol=1.0E-8, verbosity=-1)
    p.solve()
    x, s, lda, fval, kkt = p.solve()
    assert np.linalg.norm(x - g['sol1_x']) <= 1e-6
    assert np.linalg.norm(lda - g['sol1_lda']) <= 1e-6


def test_init_slack_and_lambda():
    """a12: s0 = max(ci, Ktol), lda0 = pinv(J) df with negative inequality multipliers -> Ktol."""
    for name in ('example7', 'example5', 'example8', 'qp_small', 'nlp_mid', 'nlp_eqonly'):
        prob, x0, _ = get_problem(name)
        o = OracleIPM(x0=np.array(x0), verbosity=-1, **prob.callables())
        o.nvar = prob.nvar
        o.compile()
        eng = make_engine(prob)
        eng.set_state(x0, np.ones(prob.nineq), np.zeros(prob.neq + prob.nineq), 0.2, 10.0, 0.0)
        if prob.nineq:
            eng.init_slack()
        eng.init_lambda()
        x, s, lda, _, _, _ = eng.get_state()
        lref = o.init_lambda(np.array(x0))
        if prob.nineq:
            np.testing.assert_allclose(s, o.init_slack(np.array(x0)), rtol=1e-13)
            li = lref[prob.neq:]
            li[li < 0] = 1e-4
        assert relinf(lda, lref) < 1e-7, (name, relinf(lda, lref))
        eng.close()


def test_callable_mode_matches_lowered_mode():
    """Opaque host callables (the reference's precompiled-function input mode) through set_derivs."""
    for name in ('example1', 'example5', 'example7', 'qp_small'):
        prob, x0, gts = get_problem(name)
        g = load_golden(name)
        p = IPM(x0=np.array(x0), Ftol=1.0E-8, verbosity=-1, **prob.callables())
        x, s, lda, fval, kkt = p.solve()
        assert np.linalg.norm(x - g['sol0_x']) <= 1e-6 * (1 + np.linalg.norm(g['sol0_x'])), name
        assert p.signal == int(g['sol0_signal'])


def test_mid_size_steps_vs_oracle():
    """Teacher-forced steps on problems large enough to use the DMMA kernels and several LDL^T panels."""
    for prob in (problems.make_qp(D=320, M=80, nbox=160, seed=21), problems.make_nlp(D=256, M=48, N=256, seed=22)):
        o, tr = oracle_trace(prob, prob.x0, niter=2, miter=3)
        eng = make_engine(prob)
        nu_b, de_b = 10.0, 0.0
        for k, st in enumerate(tr):
            eng.set_state(st['x'], st['s'], st['lda'], st['mu'], nu_b, de_b)
            eng.set_mu_host(st['mu_host'])
            dz, dinfo = eng.direction()
            assert dinfo.n_factor == st['reg']['n_eig'], (prob.name, k)
            assert dinfo.delta == st['delta']
            assert relinf(dz, st['dz']) < 1e-8, (prob.name, k, relinf(dz, st['dz']))
            nu_b, de_b = st['nu_after'], st['delta']
        eng.close()


def test_config3_size_properties():
    """BASELINE config 3 (D=4096, M=512, N=4096) at full size: properties that do not need the (hours-long) CPU
    oracle -- exact inertia of the accepted matrix, unreduced KKT residual of the direction, a descent
    direction for the merit function, and monotone decrease of the KKT norms over a few steps."""
    prob = problems.make_nlp()
    p = IPM(x0=prob.x0, f=prob, verbosity=-1, niter=1, miter=3)
    p.solve()
    assert len(p.step_log) == 3
    for st in p.step_log:
        assert st['n_neg'] == prob.neq and st['n_zero'] == 0
        assert st['resid'] < 1e-8
        assert st['dphi0'] < 0.0
        assert st['signal'] == 0
        assert st['alpha_s'] > 0.0 and st['n_factor'] >= 1
    # Armijo at the accepted point of the last step: phi(new) <= phi0 + alpha*eta*dphi0 (same mu, nu)
    last = p.step_log[-1]
    phi_new, _ = p.engine.merit()
    assert phi_new <= last['phi0'] + last['alpha_s'] * 1e-4 * last['dphi0'] + 1e-9 * abs(last['phi0'])


@pytest.mark.parametrize('mu', [1e-1, 1e-2, 1e-4, 1e-6, 1e-8, 1e-10])
def test_config5_mu_sweep_vs_oracle(mu):
    """BASELINE config 5 (ill-conditioned barrier sweep) on a size the CPU oracle can handle: s_i*lda_i = mu with
    25% active rows s_i = mu^0.9, so Sigma = lda/s spans up to ~1e8.  One teacher-forced step per mu: same
    delta / inertia decisions as the reference's eigvalsh test, dz to 1e-6 relative, and the
    UNREDUCED KKT residual of the refined direction small relative to the right-hand side."""
    prob = problems.make_nlp(D=192, M=32, N=192, seed=31)
    x, s, lda = problems.mu_sweep_state(prob, mu)
    o = OracleIPM(x0=x.copy(), verbosity=-1, mu=mu, **prob.callables())
    o.nvar = prob.nvar
    o.compile()
    o.mu_host, o.mu_dev, o.nu_host, o.nu_dev, o.delta, o.signal = mu, np.float64(mu), 10.0, np.float64(10.0), np.float64(0.0), 0
    tr = []
    o.trace = tr
    with np.errstate(all='ignore'):
        o.newton_step(x.copy(), s.copy(), lda.copy())
    st = tr[0]
    eng = make_engine(prob, mu=mu)
    eng.set_state(x, s, lda, mu, 10.0, 0.0)
    eng.set_mu_host(mu)
    dz, info = eng.direction()
    assert info.n_factor == st['reg']['n_eig'], (info.n_factor, st['reg']['n_eig'])
    assert info.delta == st['delta']
    assert info.n_neg == prob.neq and info.n_zero == 0
    assert relinf(dz, st['dz']) < 1e-6, relinf(dz, st['dz'])
    assert info.resid <= 1e-10 * max(1.0, np.max(np.abs(st['g'])))
    eng.close()


def test_config5_mu_sweep_full_size_properties():
    """Config 5 at the full config-3 size (no CPU oracle possible): for every mu the accepted condensed matrix has
    exactly M negative pivots and the refined direction satisfies the unreduced KKT system."""
    prob = problems.make_nlp()
    eng = make_engine(prob)
    for mu in (1e-1, 1e-4, 1e-7, 1e-10):
        x, s, lda = problems.mu_sweep_state(prob, mu)
        eng.set_state(x, s, lda, mu, 10.0, 0.0)
        eng.set_mu_host(mu)
        g, _ = eng.residual()
        dz, info = eng.direction(want_dz=False)
        assert info.n_neg == prob.neq and info.n_zero == 0, mu
        assert info.resid <= 1e-9 * max(1.0, np.max(np.abs(g))), (mu, info.resid)
    eng.close()


def test_config2_full_size_steps_vs_oracle():
    """BASELINE config 2 at full size (convex QP, D=1024, 256 eq + 1024 box inequalities, K = 3328): the first two
    Newton steps teacher-forced against the CPU oracle (each oracle step = one eigvalsh(3328) + one LU)."""
    prob = problems.make_qp()
    # The two-sided box gives EXACTLY dependent constraint gradients (dci = [E, -E]): the reference's own
    # init_lambda = pinv(J) . df (pyipm.py:729-730, numpy rcond = 1e-15) then keeps singular values of 1.6e-15 and
    # returns |lda| ~ 1e14, after which reghess multiplies delta by 10 until it overflows.  Both sides therefore
    # start from the same explicit multipliers through the reference's lda0 argument (pyipm.py:1567-1578).
    s0 = np.maximum(prob.ci(prob.x0), 1.0E-4)
    lda0 = np.concatenate([np.zeros(prob.neq), 0.2 / s0])
    o, tr = oracle_trace(prob, prob.x0, niter=1, miter=2, lda0=lda0)
    eng = make_engine(prob)
    nu_b, de_b = 10.0, 0.0
    for k, st in enumerate(tr):
        eng.set_state(st['x'], st['s'], st['lda'], st['mu'], nu_b, de_b)
        eng.set_mu_host(st['mu_host'])
        gv, _ = eng.residual()
        assert relinf(gv, -st['g']) < 1e-12
        dz, dinfo = eng.direction()
        assert dinfo.n_factor == st['reg']['n_eig'] and dinfo.delta == st['delta'] and dinfo.n_neg == prob.neq
        assert relinf(dz, st['dz']) < 1e-8, (k, relinf(dz, st['dz']))
        eng.set_state(st['x'], st['s'], st['lda'], st['mu'], nu_b, de_b)
        info = eng.newton_step()
        assert info.n_backtracks == st['search']['n_backtracks']
        assert abs(info.alpha_smax - st['alpha_smax']) < 1e-12 and abs(info.alpha_lmax - st['alpha_lmax']) < 1e-12
        x, s, lda, _, _, _ = eng.get_state()
        assert relinf(x, st['x_new']) < 1e-8 and relinf(s, st['s_new']) < 1e-8 and relinf(lda, st['lda_new']) < 1e-8
        nu_b, de_b = st['nu_after'], st['delta']
    eng.close()


def test_speculative_reghess_matches_sequential():
    """reghess (pyipm.py:1373-1406) with the delta = 0 and delta/2 attempts factored CONCURRENTLY must take exactly
    the decisions -- and produce bitwise the direction -- of the sequential loop (flags = NO_SPECULATION)."""
    prob = problems.make_nlp(D=320, M=48, N=320, seed=5)
    o, tr = oracle_trace(prob, prob.x0, niter=1, miter=4)
    eng_s = make_engine(prob, flags=0)
    eng_q = make_engine(prob, flags=1)
    nu_b, de_b = 10.0, 0.0
    n_spec = 0
    for k, st in enumerate(tr):
        out = []
        for eng in (eng_s, eng_q):
            eng.set_state(st['x'], st['s'], st['lda'], st['mu'], nu_b, de_b)
            eng.set_mu_host(st['mu_host'])
            out.append(eng.direction())
        (dz_s, i_s), (dz_q, i_q) = out
        assert i_q.n_spec == 0 and i_q.cert_used == 0
        assert i_s.n_spec == (1 if de_b > 0.0 else 0)
        assert i_s.cert_used <= i_s.n_spec      # a proof replaces the delta = 0 test only on speculated steps
        n_spec += i_s.n_spec
        assert (i_s.delta, i_s.n_factor, i_s.n_neg, i_s.n_zero) == (i_q.delta, i_q.n_factor, i_q.n_neg, i_q.n_zero)
        assert i_s.delta == st['delta'] and i_s.n_factor == st['reg']['n_eig']
        assert np.array_equal(dz_s, dz_q), (k, relinf(dz_s, dz_q))
        assert relinf(dz_s, st['dz']) < DZ_RTOL
        nu_b, de_b = st['nu_after'], st['delta']
    assert n_spec >= 1      # the nonconvex trajectory does need a shift, so speculation was exercised
    eng_s.close()
    eng_q.close()


@pytest.mark.parametrize('flags', [2, 6, 10, 6 | 128])
def test_tcgen05_syrk_path_teacher_forced(flags):
    """Engine with the d2L and condensation contractions on tcgen05 (B200IPM_FLAG_TCGEN05_SYRK, all three tile variants,
    and the condensation with all 34 slice pairs instead of 21): same reghess decisions and the same direction as the CPU oracle, to the fp64 tolerance."""
    prob = problems.make_nlp(D=320, M=48, N=320, seed=5)
    o, tr = oracle_trace(prob, prob.x0, niter=1, miter=4)
    eng = make_engine(prob, flags=flags)
    nu_b, de_b = 10.0, 0.0
    for k, st in enumerate(tr):
        eng.set_state(st['x'], st['s'], st['lda'], st['mu'], nu_b, de_b)
        eng.set_mu_host(st['mu_host'])
        dz, info = eng.direction()
        assert info.tc_syrk == 1
        assert info.delta == st['delta'] and info.n_factor == st['reg']['n_eig'] and info.n_neg == prob.neq
        assert relinf(dz, st['dz']) < DZ_RTOL, (k, relinf(dz, st['dz']))
        H = eng.hess_full()
        Href = o.hess(st['x'], st['s'], st['lda'])
        assert np.max(np.abs(H - Href)) <= 1e-12 * max(1.0, np.max(np.abs(Href)))
        eng.set_state(st['x'], st['s'], st['lda'], st['mu'], nu_b, de_b)
        info = eng.newton_step()
        assert info.n_backtracks == st['search']['n_backtracks']
        x, s, lda, _, _, _ = eng.get_state()
        assert relinf(x, st['x_new']) < 1e-8 and relinf(s, st['s_new']) < 1e-8 and relinf(lda, st['lda_new']) < 1e-8
        nu_b, de_b = st['nu_after'], st['delta']
    eng.close()


def test_tcgen05_syrk_falls_back_on_negative_multipliers():
    """A teacher-forced state with negative inequality multipliers (never produced by the IPM itself) cannot be
    represented by the unsigned operand of the tcgen05 path: the engine must notice and redo the step in fp64 DMMA."""
    prob = problems.make_nlp(D=320, M=48, N=320, seed=5)
    rng = np.random.default_rng(0)
    x = prob.x0.copy()
    s = np.maximum(prob.ci(x), 1e-2)
    lda = np.concatenate([rng.standard_normal(prob.neq), rng.standard_normal(prob.nineq)])
    out = []
    for flags in (0, 2):
        eng = make_engine(prob, flags=flags)
        eng.set_state(x, s, lda, 0.2, 10.0, 0.0)
        eng.set_mu_host(0.2)
        out.append(eng.direction())
        eng.close()
    (dz0, i0), (dz1, i1) = out
    assert i1.tc_syrk == 0 and i0.tc_syrk == 0
    assert np.array_equal(dz0, dz1) and i0.delta == i1.delta and i0.n_factor == i1.n_factor


def test_abandoned_inertia_tests_do_not_change_decisions():
    """A failed inertia test is abandoned on the device as soon as it has more than M negative pivots.  Decisions,
    delta, the number of tests and the direction must be bitwise those of the engine that completes every test
    (B200IPM_FLAG_NO_ABANDON = 16), with and without speculation."""
    prob = problems.make_nlp(D=320, M=16, N=320, seed=9)
    o, tr = oracle_trace(prob, prob.x0, niter=1, miter=3)
    # 64 = NO_CERT: the delta = 0 test really runs (in the background), 16 = NO_ABANDON, 1 = NO_SPECULATION (sequential);
    # the last engine (flags = 0) replaces the test by the negative-curvature certificate
    engs = [make_engine(prob, flags=f) for f in (64, 80, 1, 17, 0)]
    nu_b, de_b = 10.0, 0.0
    n_ab = 0
    for k, st in enumerate(tr):
        out = []
        for eng in engs:
            eng.set_state(st['x'], st['s'], st['lda'], st['mu'], nu_b, de_b)
            eng.set_mu_host(st['mu_host'])
            out.append(eng.direction())
        dz0, i0 = out[0]
        n_ab += i0.abandoned_first
        assert out[1][1].abandoned_first == 0 and out[3][1].abandoned_first == 0
        for dz, i in out[1:]:
            assert (i.delta, i.n_factor, i.n_neg, i.n_zero, i.eq_reg) == (i0.delta, i0.n_factor, i0.n_neg, i0.n_zero, i0.eq_reg)
            assert np.array_equal(dz, dz0)
        assert i0.delta == st['delta'] and i0.n_factor == st['reg']['n_eig']
        assert relinf(dz0, st['dz']) < DZ_RTOL
        nu_b, de_b = st['nu_after'], st['delta']
    assert n_ab >= 1
    for eng in engs:
        eng.close()


def test_background_inertia_test_that_passes_is_honoured():
    """Speculative reghess assumes the delta = 0 test fails and proceeds with delta/2.  On a CONVEX problem entered with
    a left-over delta > 0 the delta = 0 test passes (pyipm.py:1381): the engine must notice when it collects the
    background verdict and return what the reference's loop returns -- one test, delta unchanged, no shift applied."""
    prob = problems.make_qp(D=320, M=64, nbox=100, seed=2)
    s0 = np.maximum(prob.ci(prob.x0), 1.0E-4)
    lda0 = np.concatenate([np.zeros(prob.neq), 0.2 / s0])
    o = OracleIPM(x0=prob.x0.copy(), verbosity=-1, **prob.callables())
    o.nvar = prob.nvar
    o.compile()
    o.mu_host, o.mu_dev, o.nu_host, o.nu_dev, o.delta, o.signal = 0.2, np.float64(0.2), 10.0, np.float64(10.0), np.float64(1.0), 0
    tr = []
    o.trace = tr
    with np.errstate(all='ignore'):
        o.newton_step(prob.x0.copy(), s0.copy(), lda0.copy())
    st = tr[0]
    assert st['reg']['n_eig'] == 1 and st['delta'] == 1.0
    for flags in (0, 6, 64):     # certificate path (fp64 / tcgen05 contractions) and the real background test
        eng = make_engine(prob, flags=flags)
        eng.set_state(prob.x0, s0, lda0, 0.2, 10.0, 1.0)
        eng.set_mu_host(0.2)
        dz, info = eng.direction()
        assert info.n_factor == 1 and info.delta == 1.0 and info.n_neg == prob.neq and info.cert_used == 0
        assert relinf(dz, st['dz']) < DZ_RTOL
        # the next step from the same state does not speculate again (the previous first test passed)
        eng.set_state(prob.x0, s0, lda0, 0.2, 10.0, 1.0)
        dz2, info2 = eng.direction()
        assert info2.n_spec == 0 and info2.n_factor == 1 and np.array_equal(dz2, dz)
        eng.close()


@pytest.mark.parametrize('D,M,N', [(96, 0, 96), (96, 16, 0), (320, 48, 320)])
def test_reghess_certificate_path_all_constraint_structures(D, M, N):
    """Default engine (speculative reghess with the negative-curvature certificate) on nonconvex problems without
    equality constraints (the condensed matrix is Hb alone), without inequalities (no condensation) and with both:
    decisions and directions of the CPU oracle at every step, whether the failure of the delta = 0 test was proven
    (cert_used) or computed."""
    prob = problems.make_nlp(D=D, M=M, N=N, seed=11)
    o, tr = oracle_trace(prob, prob.x0, niter=1, miter=5)
    eng = make_engine(prob)
    nu_b, de_b = 10.0, 0.0
    n_cert = 0
    for k, st in enumerate(tr):
        eng.set_state(st['x'], st['s'], st['lda'], st['mu'], nu_b, de_b)
        eng.set_mu_host(st['mu_host'])
        dz, info = eng.direction()
        n_cert += info.cert_used
        assert info.delta == st['delta'] and info.n_factor == st['reg']['n_eig'] and info.n_neg == prob.neq, (k, info.asdict())
        assert relinf(dz, st['dz']) < DZ_RTOL, (k, relinf(dz, st['dz']))
        nu_b, de_b = st['nu_after'], st['delta']
    eng.close()
    assert n_cert >= 1     # the certificate did replace at least one delta = 0 factorisation on each trajectory


def test_callable_mode_nonconvex_delta_trace():
    """Callable mode on a nonconvex problem: the diagonal shift must be carried from step to step exactly like the
    reference's self.delta (pyipm.py:1390-1395): per-step delta and number of inertia tests of the oracle trajectory."""
    prob, x0, _ = get_problem('nlp_mid')
    o, tr = oracle_trace(prob, x0)
    p = IPM(x0=np.array(x0), Ftol=1.0E-8, verbosity=-1, **prob.callables())
    p.solve()
    assert any(st['delta'] > 0.0 for st in tr)
    n = 0
    for st, lg in zip(tr, p.step_log):
        assert lg['delta'] == st['delta'], (n, lg['delta'], st['delta'])
        assert lg['n_factor'] == st['reg']['n_eig'], (n, lg['n_factor'], st['reg']['n_eig'])
        assert lg['n_backtracks'] == st['search']['n_backtracks']
        n += 1
        if st['search']['soc_tried']:
            break      # trajectories may differ once a second-order correction is involved
    assert n >= 3


def test_callable_mode_second_order_correction_matches_oracle():
    """Callable mode now attempts the second-order correction (pyipm.py:1464-1489) like the reference: on problems whose
    oracle trajectory tries / accepts it, signal, iteration count and the final iterate agree."""
    n_tried = 0
    for name in ('nlp_small', 'nlp_mid', 'example8', 'example10'):
        prob, x0, _ = get_problem(name)
        g = load_golden(name)
        o, tr = oracle_trace(prob, x0)
        n_tried += sum(int(st['search']['soc_tried']) for st in tr)
        p = IPM(x0=np.array(x0), Ftol=1.0E-8, verbosity=-1, **prob.callables())
        x, s, lda, fval, kkt = p.solve()
        assert p.signal == int(g['sol0_signal']), name
        assert p.iter_count == int(g['sol0_nsteps']), (name, p.iter_count, int(g['sol0_nsteps']))
        assert sum(lg['soc_tried'] for lg in p.step_log) == sum(int(st['search']['soc_tried']) for st in tr)
        assert np.linalg.norm(x - g['sol0_x']) <= 1e-6 * (1 + np.linalg.norm(g['sol0_x'])), name
    assert n_tried >= 1


def test_device_callables_mode():
    """torch-CUDA callables: derivatives stay on the device (b200ipm_set_derivs with on_device = 1); same
    solution as the lowered problem."""
    import torch
    prob = problems.make_nlp(D=96, M=12, N=64, seed=61)
    dev = torch.device('cuda', 0)
    T = lambda a: torch.as_tensor(a, device=dev, dtype=torch.float64)
    Q, c, At, Ut, b, Gt, Vt, r = (T(a) for a in (prob.Q, prob.c, prob.At, prob.Ut, prob.b, prob.Gt, prob.Vt, prob.r))
    M = prob.neq
    kw = dict(
        f=lambda x: 0.5 * x @ (Q @ x) + c @ x + 0.25 * prob.q4 * (x ** 4).sum(),
        df=lambda x: Q @ x + c + prob.q4 * x ** 3,
        d2f=lambda x: Q + torch.diag(3.0 * prob.q4 * x ** 2),
        ce=lambda x: x @ At - b + 0.5 * (x @ Ut) ** 2,
        dce=lambda x: At + Ut * (x @ Ut)[None, :],
        d2ce=lambda x, lda: (Ut * lda[:M][None, :]) @ Ut.T,
        ci=lambda x: x @ Gt + r - 0.5 * (x @ Vt) ** 2,
        dci=lambda x: Gt - Vt * (x @ Vt)[None, :],
        d2ci=lambda x, lda: -(Vt * lda[M:][None, :]) @ Vt.T)
    pd = IPM(x0=prob.x0.copy(), Ftol=1.0E-8, verbosity=-1, device_callables=True, **kw)
    xd, sd, ld, fd, kd = pd.solve()
    pl = IPM(x0=prob.x0.copy(), f=prob, Ftol=1.0E-8, verbosity=-1)
    xl, sl, ll, fl, kl = pl.solve()
    assert pd.signal == pl.signal and pd.iter_count == pl.iter_count
    assert np.linalg.norm(xd - xl) <= 1e-6 * (1 + np.linalg.norm(xl))
    assert np.linalg.norm(ld - ll) <= 1e-5 * (1 + np.linalg.norm(ll))


def test_adversarial_singular_leading_tiles_at_size():
    """Pivoting is confined to 64 x 64 diagonal tiles.  Adversarial case for that design: a LINEAR objective with equality
    constraints only (example 8's structure at D = 1024, M = 256): d2L = -sum lda_e U'U has rank <= M, so with lda = 0 the
    leading D x D block of the KKT matrix is EXACTLY zero -- all 16 leading tiles singular.  The reference sees rcond <= eps
    (eq-block regularisation) and shifts; the engine must take the same decisions and produce the same direction."""
    rng = np.random.default_rng(8)
    D, M = 1024, 256
    A = rng.standard_normal((M, D)) / np.sqrt(D)
    U = rng.standard_normal((M, D)) / np.sqrt(D)
    xs = 0.5 * rng.standard_normal(D)
    ux = U @ xs
    b = A @ xs + 0.5 * ux * ux
    prob = problems.QuadProblem(np.zeros((D, D)), rng.standard_normal(D), 0.0, A=A, U=U, b=b, x0=xs + 0.05 * rng.standard_normal(D))
    tr = []
    o = OracleIPM(x0=prob.x0.copy(), verbosity=-1, niter=1, miter=2, trace=tr, lda0=np.zeros(M), **prob.callables())
    with np.errstate(all='ignore'):
        o.solve()
    assert tr[0]['reg']['eq_reg'] and tr[0]['reg']['rcond'] <= o.eps
    for flags in (6, 1):
        eng = make_engine(prob, flags=flags)
        nu_b, de_b = 10.0, 0.0
        for k, st in enumerate(tr):
            eng.set_state(st['x'], None, st['lda'], st['mu'], nu_b, de_b)
            eng.set_mu_host(st['mu_host'])
            dz, info = eng.direction()
            assert info.delta == st['delta'] and info.n_factor == st['reg']['n_eig'], (flags, k, info.asdict(), st['reg'])
            assert info.eq_reg == int(st['reg']['eq_reg'])
            assert info.n_neg == M and info.n_zero == 0
            assert relinf(dz, st['dz']) < 1e-7, (flags, k, relinf(dz, st['dz']))
            nu_b, de_b = st['nu_after'], st['delta']
        eng.close()


def test_config2_own_init_lambda_with_dependent_box_constraints():
    """Config 2's two-sided box gives EXACTLY dependent inequality gradients (dci = [E, -E]).  The reference's own
    pinv-based init_lambda blows up on it, so the parity test passes lda0 explicitly; here the engine
    starts from ITS OWN init_lambda (regularised least squares on the device): finite multipliers, and the convex QP
    converges to the same minimiser as the run started from the explicit multipliers."""
    prob = problems.make_qp(D=256, M=64, nbox=128, seed=3)
    p1 = IPM(x0=prob.x0.copy(), f=prob, Ftol=1.0E-10, Ktol=1.0E-6, verbosity=-1, niter=20)
    x1, s1, l1, f1, k1 = p1.solve()
    s0 = np.maximum(prob.ci(prob.x0), 1.0E-4)
    lda0 = np.concatenate([np.zeros(prob.neq), 0.2 / s0])
    p2 = IPM(x0=prob.x0.copy(), f=prob, lda0=lda0, Ftol=1.0E-10, Ktol=1.0E-6, verbosity=-1, niter=20)
    x2, s2, l2, f2, k2 = p2.solve()
    assert np.all(np.isfinite(l1)) and p1.signal in (1, 2) and p2.signal in (1, 2)
    assert np.linalg.norm(x1 - x2) <= 1e-4 * (1.0 + np.linalg.norm(x2))
    assert abs(f1 - f2) <= 1e-6 * (1.0 + abs(f2))


@pytest.mark.parametrize('D,M,N,cond', [(80, 30, 0, 1e6), (40, 12, 50, 1e6), (64, 16, 24, 1e3), (48, 48, 0, 1e5)])
def test_init_lambda_ill_conditioned_jacobian(D, M, N, cond):
    """a12 / ADVICE r1: lda0 = pinv(J) df (pyipm.py:729-730) for a Jacobian [dce | dci] with singular values graded over
    `cond` -- both Gram sides (C < D: J'J, C >= D: J J') of the nonstationary iterated-Tikhonov solve, against NumPy's
    minimum-norm least-squares solution.  Tolerance 1e-6 relative (the fixed six sweeps at t = 1e-7 of round 1 left the
    components below sigma_max * 3e-4 unconverged: errors of order one at cond = 1e6)."""
    rng = np.random.default_rng(D * 1000 + M + N)
    C = M + N
    r = min(D, C)
    P, _ = np.linalg.qr(rng.standard_normal((D, r)))
    Qm, _ = np.linalg.qr(rng.standard_normal((C, r)))
    sig = np.logspace(0.0, -np.log10(cond), r)
    J = (P * sig) @ Qm.T                                   # D x C
    Qh = rng.standard_normal((D, D))
    Qh = (Qh + Qh.T) / (2.0 * np.sqrt(D))
    x0 = rng.standard_normal(D)
    A = J[:, :M].T.copy()
    G = J[:, M:].T.copy() if N else None
    prob = problems.QuadProblem(Qh, rng.standard_normal(D), 0.0, A=A, b=np.zeros(M), G=G,
                                r=(np.abs(rng.standard_normal(N)) + 5.0) if N else None, x0=x0, name='illcond')
    eng = make_engine(prob)
    eng.set_state(x0, np.ones(N), np.zeros(C), 0.2, 10.0, 0.0)
    if N:
        eng.init_slack()
    eng.init_lambda()
    _, _, lda, _, _, _ = eng.get_state()
    eng.close()
    lref = np.linalg.lstsq(J, prob.df(x0), rcond=None)[0]
    if N:
        lref[M:][lref[M:] < 0] = 1e-4
    assert relinf(lda, lref) < 1e-6, relinf(lda, lref)
