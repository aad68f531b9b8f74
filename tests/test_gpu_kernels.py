"""Kernel-level parity on the B200 (through the C ABI): GEMV, the DMMA A*diag(w)*A' contraction and the
tile-pivoted LDL^T (factor, inertia, solve) against NumPy/SciPy on the same seeded inputs."""
import numpy as np
import pytest
import scipy.linalg

from pyipm_b200 import _lib, problems

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('rows,cols', [(3, 4), (17, 33), (128, 130), (1000, 257), (512, 2048)])
@pytest.mark.parametrize('transpose', [False, True])
def test_gemv(rows, cols, transpose):
    rng = np.random.default_rng(rows * 1000 + cols)
    A = rng.standard_normal((rows, cols))
    v = rng.standard_normal(rows if transpose else cols)
    y = _lib.test_gemv(A, v, transpose)
    ref = A.T @ v if transpose else A @ v
    np.testing.assert_allclose(y, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())


def _syrk_ref(n, Cin, beta, dadd, shift, terms):
    C = np.zeros((n, n))
    if Cin is not None:
        U = np.triu(Cin)
        C += beta * (U + np.triu(U, 1).T)       # reference quirk: only triu(d2L) is used (pyipm.py:785,843)
    if dadd is not None:
        C += np.diag(dadd)
    C += shift * np.eye(n)
    for A, w, alpha in terms:
        C += alpha * (A if w is None else A * w[None, :]) @ A.T
    return C


@pytest.mark.parametrize('n,Ks,force_simple', [(5, [3], False), (64, [20], False), (130, [70, 33], False),
                                               (130, [70, 33], True), (257, [64], False), (512, [256, 512, 100], False),
                                               (1000, [1000], False)])
def test_syrk_adat(n, Ks, force_simple):
    rng = np.random.default_rng(n + sum(Ks))
    Cin = rng.standard_normal((n, n))            # deliberately NOT symmetric
    dadd = rng.standard_normal(n)
    terms = []
    for i, K in enumerate(Ks):
        w = None if i == 2 else 10.0 ** rng.uniform(-3, 3, K)
        terms.append((rng.standard_normal((n, K)), w, [-1.0, 1.0, 0.5][i]))
    C, ms = _lib.test_syrk(n, Cin, 1.0, dadd, 0.25, terms, force_simple=force_simple)
    ref = _syrk_ref(n, Cin, 1.0, dadd, 0.25, terms)
    scale = sum(np.abs(A) @ (np.abs(A) if w is None else np.abs(A) * w).T for A, w, _ in terms) + np.abs(ref) + 1.0
    assert np.array_equal(C, C.T), 'output must be bitwise symmetric'
    assert np.max(np.abs(C - ref) / scale) < 1e-14 * max(Ks)


def _check_ldlt(A, nrhs=2, tol=1e-10, expect_inertia=True):
    n = A.shape[0]
    rng = np.random.default_rng(n)
    F = _lib.DenseLDLT(n)
    (pos, neg, zero), rcond = F.factor(A)
    w = np.linalg.eigvalsh(A)
    if expect_inertia:
        assert (pos, neg, zero) == (int(np.sum(w > 0)), int(np.sum(w < 0)), 0)
    B = rng.standard_normal((n, nrhs))
    X = F.solve(B, nrefine=2)
    Xref = np.linalg.solve(A, B)
    err = np.max(np.abs(X - Xref)) / np.max(np.abs(Xref))
    res = np.max(np.abs(A @ X - B)) / (np.max(np.abs(A)) * np.max(np.abs(X)) * n)
    F.close()
    assert res < 1e-14, res
    assert err < tol, err
    return rcond


@pytest.mark.parametrize('n', [1, 2, 7, 64, 65, 100, 257, 640, 1000])
def test_ldlt_spd(n):
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, n))
    A = B @ B.T / n + np.eye(n)
    _check_ldlt(A)


@pytest.mark.parametrize('n', [2, 10, 64, 100, 300, 777])
def test_ldlt_indefinite_random(n):
    rng = np.random.default_rng(100 + n)
    B = rng.standard_normal((n, n))
    A = (B + B.T) / np.sqrt(n)
    _check_ldlt(A, tol=1e-8)


@pytest.mark.parametrize('D,M', [(3, 1), (2, 1), (40, 10), (100, 28), (300, 64), (900, 124)])
def test_ldlt_saddle_point_inertia(D, M):
    """Condensed-KKT structure [[H, A], [A', 0]] with an indefinite H whose reduced Hessian is positive
    definite: inertia must be exactly (D, M, 0) -- the acceptance test of reghess (pyipm.py:1381)."""
    rng = np.random.default_rng(D * 7 + M)
    Aj = rng.standard_normal((D, M))
    Q, _ = np.linalg.qr(np.concatenate([Aj, rng.standard_normal((D, D - M))], axis=1))
    Z = Q[:, M:]                                   # null-space basis of Aj'
    Y = Q[:, :M]
    S = rng.standard_normal((D - M, D - M))
    H = Z @ (S @ S.T / D + np.eye(D - M)) @ Z.T - 0.5 * Y @ Y.T    # negative curvature only in range(Aj)
    K = np.zeros((D + M, D + M))
    K[:D, :D] = (H + H.T) / 2
    K[:D, D:] = Aj
    K[D:, :D] = Aj.T
    n = D + M
    F = _lib.DenseLDLT(n)
    (pos, neg, zero), _ = F.factor(K)
    assert (pos, neg, zero) == (D, M, 0)
    b = rng.standard_normal(n)
    x = F.solve(b, nrefine=2)
    F.close()
    xref = np.linalg.solve(K, b)
    assert np.max(np.abs(x - xref)) / np.max(np.abs(xref)) < 1e-9


def test_ldlt_zero_diagonal_tile():
    """Example-8-like structure (linear objective): H has exact zeros on the diagonal; inside one tile the
    Bunch-Kaufman 2x2 pivots must handle it."""
    K = np.array([[0.0, 0, 0, 2, 1.0], [0, 0, 0, -1, 2.0], [0, 0, 0, -1, 0.0], [2, -1, -1, 0, 0.0], [1, 2, 0, 0, 0.0]])
    K[0, 0] = K[1, 1] = -0.7
    F = _lib.DenseLDLT(5)
    (pos, neg, zero), _ = F.factor(K)
    w = np.linalg.eigvalsh(K)
    assert (pos, neg, zero) == (int(np.sum(w > 0)), int(np.sum(w < 0)), 0)
    b = np.arange(1.0, 6.0)
    np.testing.assert_allclose(F.solve(b), np.linalg.solve(K, b), rtol=1e-11)
    F.close()


def test_ldlt_singular_reports_zero_pivot():
    K = np.zeros((4, 4))
    K[0, 0] = 1.0
    K[1, 1] = 2.0
    F = _lib.DenseLDLT(4)
    (pos, neg, zero), rcond = F.factor(K)
    F.close()
    assert zero == 2 and pos == 2 and rcond == 0.0


# (variant word of b200ipm_test_syrk_i8, tolerance relative to the row-scale products): tile shapes 128x64 / 128x128 /
# 128x256 with 7 slice-pair diagonals, and the 128x128 shape with 8 (34 pairs, fp64-level), 7 (28) and 6 (21) diagonals
OZ_VARIANTS = [(0, 1.6e-15), (2, 1.6e-15), (1 + 16 * 8, 1.6e-15), (1 + 16 * 7, 1.6e-15), (1 + 16 * 6, 1e-12)]


@pytest.mark.parametrize('variant,tol', OZ_VARIANTS)
@pytest.mark.parametrize('n,Ks', [(128, [32]), (130, [70, 33]), (300, [70, 203]), (640, [512, 100, 40])])
def test_syrk_tcgen05_int8_matches_fp64(n, Ks, variant, tol):
    """The tcgen05 path (error-free int8 split into balanced base-256 digits, int32 TMEM accumulators, fp64
    recombination) must reproduce the fp64 product: signed weights on the first term (lda_e), weights spanning 12
    orders of magnitude on the second (Sigma / lda_i), ragged sizes, asymmetric Cin, diagonal add."""
    rng = np.random.default_rng(n + sum(Ks) + variant)
    Cin = rng.standard_normal((n, n))
    dadd = rng.standard_normal(n)
    terms = []
    for i, K in enumerate(Ks):
        if i == 0:
            w = rng.standard_normal(K)
        elif i == 1:
            w = 10.0 ** rng.uniform(-6, 6, K)
        else:
            w = None
        terms.append((rng.standard_normal((n, K)), w, [-1.0, 1.0, 0.5][i]))
    C, ms, err = _lib.test_syrk_i8(n, Cin, 1.0, dadd, 0.25, terms, signed_mask=1, variant=variant)
    assert err == 0
    ref = _syrk_ref(n, Cin, 1.0, dadd, 0.25, terms)
    # bound relative to the products of the row scales (the fixed-point split is relative to each row's largest entry)
    rowmax = np.sqrt(sum(((np.abs(A) * np.sqrt(np.abs(al * (1.0 if w is None else w)))).max(axis=1)) ** 2 for A, w, al in terms))
    scale = np.outer(rowmax, rowmax) * sum(Ks) + np.abs(ref) + 1.0
    assert np.array_equal(C, C.T), 'output must be bitwise symmetric'
    assert np.max(np.abs(C - ref) / scale) < tol
    # and against the DMMA kernel on the same inputs
    Cd, _ = _lib.test_syrk(n, Cin, 1.0, dadd, 0.25, terms)
    assert np.max(np.abs(C - Cd) / scale) < max(tol, 1e-15 * max(Ks))


def test_syrk_tcgen05_reports_unannounced_negative_weight():
    rng = np.random.default_rng(5)
    A = rng.standard_normal((128, 64))
    w = rng.standard_normal(64)
    _, _, err = _lib.test_syrk_i8(128, None, 0.0, None, 0.0, [(A, w, 1.0)], signed_mask=0)
    assert err & 2
    A[3, 5] = np.inf
    _, _, err = _lib.test_syrk_i8(128, None, 0.0, None, 0.0, [(A, np.abs(w), 1.0)], signed_mask=0)
    assert err & 1


@pytest.mark.parametrize('tc', ['1', '0'])
def test_ldlt_tcgen05_trailing_updates(tc, monkeypatch):
    """The bulk trailing updates of the factorisation on tcgen05 (int8 error-free split, 21 slice pairs, B200IPM_LDLT_TC=1,
    default for n >= 2048) against the fp64 DMMA updates (=0): same inertia of a quasi-definite KKT matrix with diagonal
    entries spanning 8 decades, solution to 1e-9 after refinement either way."""
    monkeypatch.setenv('B200IPM_LDLT_TC', tc)
    n, m = 2560, 320
    K, rhs = problems.make_dense_kkt(n, m, seed=77)
    F = _lib.DenseLDLT(n)
    (pos, neg, zero), _ = F.factor(K)
    assert (pos, neg, zero) == (n - m, m, 0)
    X = F.solve(rhs[:, :2], nrefine=2)
    F.close()
    Xref = np.linalg.solve(K, rhs[:, :2])
    assert np.max(np.abs(X - Xref)) / np.max(np.abs(Xref)) < 1e-9
    res = np.max(np.abs(K @ X - rhs[:, :2])) / (np.max(np.abs(K)) * np.max(np.abs(X)))
    assert res < 1e-14, res


@pytest.mark.parametrize('n,m', [(100, 20), (257, 31), (1000, 124), (2560, 320), (4608, 512)])
def test_ldlt_block256_solve_matches_tile_chain(n, m, monkeypatch):
    """The block-256 cluster solves (explicit inverses of the 256-row diagonal blocks, one DSMEM exchange per link;
    default) against the 64-row chain (B200IPM_SOLVE256=0): same factorisation, UNREFINED solutions of a quasi-definite
    KKT matrix agree to the conditioning-level error of either, both reach 1e-9 of LAPACK after refinement."""
    K, rhs = problems.make_dense_kkt(n, m, seed=n)
    Xref = np.linalg.solve(K, rhs[:, :2])
    raw = {}
    for flag in ('1', '0'):
        monkeypatch.setenv('B200IPM_SOLVE256', flag)
        F = _lib.DenseLDLT(n)
        (pos, neg, zero), _ = F.factor(K)
        assert (pos, neg, zero) == (n - m, m, 0)
        raw[flag] = F.solve(rhs[:, :2], nrefine=0)
        X = F.solve(rhs[:, :2], nrefine=2)
        F.close()
        assert np.max(np.abs(X - Xref)) / np.max(np.abs(Xref)) < 1e-9
    sc = np.max(np.abs(Xref))
    e1, e0 = np.max(np.abs(raw['1'] - Xref)) / sc, np.max(np.abs(raw['0'] - Xref)) / sc
    assert e1 < max(100.0 * e0, 1e-10), (e1, e0)
