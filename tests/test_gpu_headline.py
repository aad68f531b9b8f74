"""Parity at the HEADLINE configuration (BASELINE.json configs 3 and 5: D = 4096, M = 512, N = 4096, K = 12800), on the
default engine flags (tcgen05 contractions, speculative reghess with the negative-curvature certificate), against
independent fp64 computations on the host:

* residual and the full K x K ``hess()`` against the oracle's assembly (pyipm.py:610-653, 768-844);
* the direction against ``scipy.linalg.solve(K_full, -g, assume_a='gen')`` on the 12800^2 system -- literally what the
  reference's ``sym_solve_cmp`` does (pyipm.py:18-20, 911-914, 1720-1725);
* the reghess decisions (pyipm.py:1373-1406) against the inertia of the FULL matrix computed by LAPACK ``dsytrf``
  (Sylvester: inertia of the block-diagonal D), i.e. the quantity the reference gets from ``eigvalsh``;
* the tcgen05 SYRK kernel at n = 4096, K = [512, 4096].

Tolerances: SURVEY.md section 8(c)."""
import numpy as np
import pytest
import scipy.linalg
import scipy.linalg.lapack

from oracle.pyipm_numpy import OracleIPM
from pyipm_b200 import _lib, problems

pytestmark = pytest.mark.gpu


def relinf(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300)


def sytrf_inertia(A):
    """(pos, neg, zero) of a symmetric matrix from LAPACK's Bunch-Kaufman factorisation (1x1 / 2x2 blocks of D)."""
    ldu, piv, info = scipy.linalg.lapack.dsytrf(A, lower=1, overwrite_a=0)
    assert info >= 0
    n = A.shape[0]
    d = np.diagonal(ldu)
    sub = np.diagonal(ldu, -1)
    pos = neg = zero = 0
    i = 0
    while i < n:
        if piv[i] > 0:
            v = d[i]
            pos += v > 0
            neg += v < 0
            zero += v == 0
            i += 1
        else:   # 2x2 block [[d_i, e], [e, d_{i+1}]]
            a, b, c = d[i], sub[i], d[i + 1]
            hm, hd = 0.5 * (a + c), 0.5 * (a - c)
            rad = np.hypot(hd, b)
            for e in (hm + rad, hm - rad):
                pos += e > 0
                neg += e < 0
                zero += e == 0
            i += 2
    return int(pos), int(neg), int(zero)


def oracle_at(prob, mu, nu):
    o = OracleIPM(x0=prob.x0.copy(), verbosity=-1, mu=mu, **prob.callables())
    o.nvar = prob.nvar
    o.compile()
    o.mu_host, o.mu_dev, o.nu_host, o.nu_dev, o.signal = mu, np.float64(mu), nu, np.float64(nu), 0
    return o


def check_step_against_host(prob, eng, x, s, lda, mu, nu, delta_in, dz_tol, check_first_test=True):
    """One teacher-forced direction of the engine from (x, s, lda, mu, nu, delta_in) against the host computations."""
    D, M, N = prob.nvar, prob.neq, prob.nineq
    o = oracle_at(prob, mu, nu)
    eng.set_state(x, s, lda, mu, nu, delta_in)
    eng.set_mu_host(mu)
    gv, nrm = eng.residual()
    g_ref = o.grad(x, s, lda)
    assert relinf(gv, g_ref) < 1e-12
    kk = o.KKT(x, s, lda)
    np.testing.assert_allclose(nrm, [np.linalg.norm(k) for k in kk], rtol=1e-10)
    # a3: the reference's full K x K matrix (tcgen05 d2L inside)
    H = eng.hess_full()
    Href = o.hess(x, s, lda)
    assert np.array_equal(H, H.T)
    assert np.max(np.abs(H - Href)) <= 1e-12 * max(1.0, np.max(np.abs(Href)))
    W = eng.d2L()
    assert np.max(np.abs(W - Href[:D, :D])) <= 1e-12 * max(1.0, np.max(np.abs(Href[:D, :D])))
    del H, W
    # a4 + a5
    dz, info = eng.direction()
    assert info.tc_syrk == 1
    assert info.n_neg == M and info.n_zero == 0
    Hc = Href
    idx = np.arange(D)
    if check_first_test and info.n_factor >= 2:
        # the reference's first test (pyipm.py:1378-1381) on the unshifted matrix must fail on inertia
        p0, n0, z0 = sytrf_inertia(Hc)
        assert n0 != M + N or z0 != 0, (p0, n0, z0)
    if info.n_factor == 2:
        d1 = np.sqrt(np.finfo(np.float64).eps)
        assert info.delta == (d1 if delta_in == 0.0 else max(delta_in / 2.0, d1))
    elif info.n_factor == 1:
        assert info.delta == delta_in
    Hc[idx, idx] += info.delta
    if info.eq_reg:
        ie = D + N + np.arange(M)
        Hc[ie, ie] -= np.sqrt(np.finfo(np.float64).eps) * 1e-4 * mu ** 0.4
    # the accepted matrix has the reference's inertia (D + N, M + N, 0)  (pyipm.py:1399)
    assert sytrf_inertia(Hc) == (D + N, M + N, 0)
    if info.n_factor > 2:
        # every rejected shift of the delta *= 10 loop (pyipm.py:1399-1403) must indeed fail: check the last one
        Hc[idx, idx] += info.delta / 10.0 - info.delta
        pr, nr, zr = sytrf_inertia(Hc)
        assert nr != M + N or zr != 0
        Hc[idx, idx] += info.delta - info.delta / 10.0
    if info.n_factor > 2:
        assert info.delta == np.sqrt(np.finfo(np.float64).eps) * 10.0 ** (info.n_factor - 2) or delta_in != 0.0
    dz_ref = scipy.linalg.solve(Hc, -g_ref.reshape(-1, 1), assume_a='gen', overwrite_a=True).reshape(-1)
    dz_ref[D + N:] = -dz_ref[D + N:]
    err = relinf(dz, dz_ref)
    assert err < dz_tol, err
    return info, err


def test_config3_full_size_step_vs_host_lapack():
    """Config 3 at full size, default flags, a state three real Newton steps into the solve (delta > 0, so reghess
    speculates and the certificate path is the one exercised)."""
    prob = problems.make_nlp()
    D, M, N = prob.nvar, prob.neq, prob.nineq
    eng = _lib.Engine(D, M, N, _lib.default_params())
    eng.bind(prob)
    eng.set_state(prob.x0, np.ones(N), np.zeros(M + N), 0.2, 10.0, 0.0)
    eng.set_mu_host(0.2)
    eng.init_slack()
    eng.init_lambda()
    # a12 at full size: s0 = max(ci, Ktol); lda0 = least-squares multipliers
    x, s, lda, _, _, _ = eng.get_state()
    np.testing.assert_allclose(s, np.maximum(prob.ci(prob.x0), 1e-4), rtol=1e-13)
    # the very first step (delta = 0: the sequential reghess loop) ...
    info0, err0 = check_step_against_host(prob, eng, x, s, lda, 0.2, 10.0, 0.0, 1e-9)
    assert info0.n_factor >= 2 and info0.delta > 0.0
    # ... and a speculated one, three steps in
    eng.set_state(x, s, lda, 0.2, 10.0, 0.0)
    for _ in range(3):
        eng.newton_step()
    x, s, lda, mu, nu, delta = eng.get_state()
    assert delta > 0.0
    info, err = check_step_against_host(prob, eng, x, s, lda, mu, nu, delta, 1e-9)
    assert info.n_spec == 1 and info.n_factor == 2
    eng.close()


@pytest.mark.parametrize('mu', [1e-10])
def test_config5_full_size_step_vs_host_lapack(mu):
    """Config 5 (ill-conditioned barrier state, Sigma spanning ~10 decades) at full size: direction to 1e-6 relative
     against the LU solve of the unreduced 12800^2 system, inertia of the accepted matrix by dsytrf."""
    prob = problems.make_nlp()
    eng = _lib.Engine(prob.nvar, prob.neq, prob.nineq, _lib.default_params(mu=mu))
    eng.bind(prob)
    x, s, lda = problems.mu_sweep_state(prob, mu)
    check_step_against_host(prob, eng, x, s, lda, mu, 10.0, 0.0, 1e-6, check_first_test=False)
    eng.close()


@pytest.mark.parametrize('mu', [1e-1, 1e-2, 1e-4, 1e-6, 1e-8, 1e-10])
def test_config5_mu_sweep_vs_oracle_tcgen05_path(mu):
    """The config-5 sweep against the CPU oracle on a size where the default (tcgen05, D >= 256) path runs: same delta,
    same number of inertia tests, dz to 1e-6, from delta_in = 0 and from a left-over delta (speculative path)."""
    prob = problems.make_nlp(D=320, M=48, N=320, seed=31)
    x, s, lda = problems.mu_sweep_state(prob, mu)
    for delta_in in (0.0, 0.5):
        o = oracle_at(prob, mu, 10.0)
        o.delta = np.float64(delta_in)
        tr = []
        o.trace = tr
        with np.errstate(all='ignore'):
            o.newton_step(x.copy(), s.copy(), lda.copy())
        st = tr[0]
        eng = _lib.Engine(prob.nvar, prob.neq, prob.nineq, _lib.default_params(mu=mu))
        eng.bind(prob)
        eng.set_state(x, s, lda, mu, 10.0, delta_in)
        eng.set_mu_host(mu)
        dz, info = eng.direction()
        assert info.tc_syrk == 1
        assert info.n_factor == st['reg']['n_eig'], (mu, delta_in, info.n_factor, st['reg']['n_eig'])
        assert info.delta == st['delta']
        assert info.n_neg == prob.neq and info.n_zero == 0
        assert relinf(dz, st['dz']) < 1e-6, (mu, delta_in, relinf(dz, st['dz']))
        assert info.resid <= 1e-10 * max(1.0, np.max(np.abs(st['g'])))
        eng.close()


@pytest.mark.parametrize('ndiag,tol', [(8, 1.6e-15), (7, 1.6e-15), (6, 1e-12)])
def test_syrk_tcgen05_int8_config3_shape(ndiag, tol):
    """The tcgen05 SYRK at the headline shape n = 4096, K = [512, 4096]: signed weights on the first term (lda_e), 12
    decades of weights on the second (Sigma); 128x128 tiles with 8 / 7 / 6 slice-pair diagonals."""
    n, Ks = 4096, [512, 4096]
    rng = np.random.default_rng(4096 + ndiag)
    Cin = rng.standard_normal((n, n))
    dadd = rng.standard_normal(n)
    A0 = rng.standard_normal((n, Ks[0])) / 64.0
    A1 = rng.standard_normal((n, Ks[1])) / 64.0
    w0 = rng.standard_normal(Ks[0])
    w1 = 10.0 ** rng.uniform(-6, 6, Ks[1])
    terms = [(A0, w0, -1.0), (A1, w1, 1.0)]
    C, ms, err = _lib.test_syrk_i8(n, Cin, 1.0, dadd, 0.25, terms, signed_mask=1, variant=1 + 16 * ndiag)
    assert err == 0
    up = np.triu(Cin)
    ref = up + np.triu(Cin, 1).T + np.diag(dadd) + 0.25 * np.eye(n)
    ref -= (A0 * w0[None, :]) @ A0.T
    ref += (A1 * w1[None, :]) @ A1.T
    rowmax = np.sqrt(((np.abs(A0) * np.sqrt(np.abs(w0))).max(axis=1)) ** 2 + ((np.abs(A1) * np.sqrt(w1)).max(axis=1)) ** 2)
    scale = np.outer(rowmax, rowmax) * sum(Ks) + np.abs(ref) + 1.0
    assert np.array_equal(C, C.T)
    assert np.max(np.abs(C - ref) / scale) < tol
