"""Problem descriptors for the B200 Newton-step engine.

Aesara's symbolic autodiff (pyipm.py:473-509) is replaced by two built-in *lowerable* problem forms whose
derivatives are evaluated on the device, plus plain user callables:

* :class:`PolyProblem` -- sparse multivariate polynomials (monomial tables) with an optional elementwise
  ``c * sum(x*log(x+shift))`` term.  Covers all ten example problems of the reference
  (pyipm.py:1920-2131, unit_tests.py:96-237).
* :class:`QuadProblem` -- the dense synthetic family used by BASELINE.json configs 2/3/5
  (SURVEY.md section 8d): ``f = 1/2 x'Qx + c'x + q4/4 sum(x^4)``, ``ce = Ax + 1/2 (Ux)^2 - b``,
  ``ci = Gx - 1/2 (Vx)^2 + r``.

Both expose NumPy callables with the reference's "precompiled function" conventions (pyipm.py:216-231:
``dce`` is D x M, ``dci`` is D x N, ``d2ce(x, lda)`` / ``d2ci(x, lda)`` take the FULL multiplier vector) --
these are what the CPU oracle and the reference itself are driven with in the tests -- and a flat device
descriptor consumed by ``libb200ipm.so`` (include/b200ipm.h: b200ipm_bind_poly / b200ipm_bind_quad).
"""
from __future__ import print_function

import numpy as np


# --------------------------------------------------------------------------------------------- PolyProblem
class PolyProblem(object):
    """Sparse polynomial NLP.

    A *term* is ``(coeff, ((var, power), ...))`` with distinct ``var`` and integer ``power >= 1``; a constant
    is ``(coeff, ())``.  ``f_terms`` is a list of terms; ``ce_terms`` / ``ci_terms`` are lists (one entry per
    constraint) of lists of terms.  ``xlogx=(coeff, shift)`` adds ``coeff*sum_i x_i*log(x_i+shift)`` to f
    (example 6, pyipm.py:2027).
    """

    def __init__(self, nvar, f_terms, ce_terms=None, ci_terms=None, xlogx=None, name=None):
        self.nvar = int(nvar)
        self.f_terms = [self._norm(t) for t in f_terms]
        self.ce_terms = [[self._norm(t) for t in row] for row in (ce_terms or [])]
        self.ci_terms = [[self._norm(t) for t in row] for row in (ci_terms or [])]
        self.neq = len(self.ce_terms)
        self.nineq = len(self.ci_terms)
        self.xlogx = None if xlogx is None else (float(xlogx[0]), float(xlogx[1]))
        self.name = name

    @staticmethod
    def _norm(t):
        c, facs = t
        facs = tuple((int(v), int(p)) for v, p in facs)
        assert len(set(v for v, _ in facs)) == len(facs), 'repeated variable in a monomial'
        assert all(p >= 1 for _, p in facs)
        return float(c), facs

    # ---- scalar polynomial helpers (deterministic term order)
    @staticmethod
    def _val(terms, x):
        acc = 0.0
        for c, facs in terms:
            m = c
            for v, p in facs:
                m = m * x[v] ** p
            acc = acc + m
        return acc

    def _grad(self, terms, x):
        g = np.zeros(self.nvar)
        for c, facs in terms:
            for a, (va, pa) in enumerate(facs):
                m = c * pa * x[va] ** (pa - 1)
                for b, (vb, pb) in enumerate(facs):
                    if b != a:
                        m = m * x[vb] ** pb
                g[va] += m
        return g

    def _hess(self, terms, x):
        H = np.zeros((self.nvar, self.nvar))
        for c, facs in terms:
            for a, (va, pa) in enumerate(facs):
                # diagonal second derivative
                if pa >= 2:
                    m = c * pa * (pa - 1) * x[va] ** (pa - 2)
                    for b, (vb, pb) in enumerate(facs):
                        if b != a:
                            m = m * x[vb] ** pb
                    H[va, va] += m
                for b, (vb, pb) in enumerate(facs):
                    if b == a:
                        continue
                    m = c * pa * x[va] ** (pa - 1) * pb * x[vb] ** (pb - 1)
                    for k, (vk, pk) in enumerate(facs):
                        if k != a and k != b:
                            m = m * x[vk] ** pk
                    H[va, vb] += m
        return H

    # ---- reference-convention callables
    def f(self, x):
        x = np.asarray(x, dtype=np.float64)
        v = self._val(self.f_terms, x)
        if self.xlogx is not None:
            c, sh = self.xlogx
            v = v + c * np.sum(x * np.log(x + sh))
        return np.float64(v)

    def df(self, x):
        x = np.asarray(x, dtype=np.float64)
        g = self._grad(self.f_terms, x)
        if self.xlogx is not None:
            c, sh = self.xlogx
            g = g + c * (np.log(x + sh) + x / (x + sh))
        return g

    def d2f(self, x):
        x = np.asarray(x, dtype=np.float64)
        H = self._hess(self.f_terms, x)
        if self.xlogx is not None:
            c, sh = self.xlogx
            H = H + np.diag(c * (1.0 / (x + sh) + sh / (x + sh) ** 2))
        return H

    def ce(self, x):
        x = np.asarray(x, dtype=np.float64)
        return np.array([self._val(r, x) for r in self.ce_terms], dtype=np.float64)

    def dce(self, x):
        x = np.asarray(x, dtype=np.float64)
        return np.stack([self._grad(r, x) for r in self.ce_terms], axis=1)

    def d2ce(self, x, lda):
        x = np.asarray(x, dtype=np.float64)
        H = np.zeros((self.nvar, self.nvar))
        for j, r in enumerate(self.ce_terms):
            H = H + lda[j] * self._hess(r, x)
        return H

    def ci(self, x):
        x = np.asarray(x, dtype=np.float64)
        return np.array([self._val(r, x) for r in self.ci_terms], dtype=np.float64)

    def dci(self, x):
        x = np.asarray(x, dtype=np.float64)
        return np.stack([self._grad(r, x) for r in self.ci_terms], axis=1)

    def d2ci(self, x, lda):
        x = np.asarray(x, dtype=np.float64)
        H = np.zeros((self.nvar, self.nvar))
        for j, r in enumerate(self.ci_terms):
            H = H + lda[self.neq + j] * self._hess(r, x)
        return H

    def callables(self):
        """kwargs for OracleIPM / the reference's precompiled-function mode."""
        kw = dict(f=self.f, df=self.df, d2f=self.d2f)
        if self.neq:
            kw.update(ce=self.ce, dce=self.dce, d2ce=self.d2ce)
        if self.nineq:
            kw.update(ci=self.ci, dci=self.dci, d2ci=self.d2ci)
        return kw

    # ---- device descriptor (CSR monomial table; row 0 = f, 1..M = ce, M+1..M+N = ci)
    def descriptor(self):
        rows = [self.f_terms] + self.ce_terms + self.ci_terms
        term_row, term_coeff, term_ptr, fac_var, fac_pow = [], [], [0], [], []
        for r, terms in enumerate(rows):
            for c, facs in terms:
                term_row.append(r)
                term_coeff.append(c)
                for v, p in facs:
                    fac_var.append(v)
                    fac_pow.append(p)
                term_ptr.append(len(fac_var))
        return dict(
            nterms=len(term_row),
            term_row=np.asarray(term_row, dtype=np.int32),
            term_coeff=np.asarray(term_coeff, dtype=np.float64),
            term_ptr=np.asarray(term_ptr, dtype=np.int32),
            fac_var=np.asarray(fac_var if fac_var else [0], dtype=np.int32),
            fac_pow=np.asarray(fac_pow if fac_pow else [0], dtype=np.int32),
            xlogx_coeff=0.0 if self.xlogx is None else self.xlogx[0],
            xlogx_shift=0.0 if self.xlogx is None else self.xlogx[1],
        )


def _lin(coeffs, const=0.0):
    """terms of a linear form sum_i coeffs[i]*x_i + const (zero coefficients dropped)."""
    t = [(c, ((i, 1),)) for i, c in enumerate(coeffs) if c != 0.0]
    if const != 0.0:
        t.append((const, ()))
    return t


_EPS = float(np.finfo(np.float64).eps)


def example_problem(k):
    """The ten example problems of the reference (pyipm.py:1920-2131; same set as unit_tests.py:96-237).

    Returns (PolyProblem, ground_truths) where ground_truths is the list of acceptable minimisers."""
    s2, s3, s13 = np.sqrt(2.0), np.sqrt(3.0), np.sqrt(13.0)
    if k == 1:   # pyipm.py:1925-1926
        p = PolyProblem(2, [(1.0, ((0, 2),)), (-4.0, ((0, 1),)), (1.0, ((1, 2),)), (-1.0, ((1, 1),)),
                            (-1.0, ((0, 1), (1, 1)))])
        gt = [np.array([3.0, 2.0])]
    elif k == 2:  # pyipm.py:1943  100*(y-x^2)^2 + (1-x)^2 expanded
        p = PolyProblem(2, [(100.0, ((1, 2),)), (-200.0, ((0, 2), (1, 1))), (100.0, ((0, 4),)),
                            (1.0, ()), (-2.0, ((0, 1),)), (1.0, ((0, 2),))])
        gt = [np.array([1.0, 1.0])]
    elif k == 3:  # pyipm.py:1959-1960
        p = PolyProblem(2, _lin([-1.0, -1.0]), ce_terms=[[(1.0, ((0, 2),)), (1.0, ((1, 2),)), (-1.0, ())]])
        gt = [np.array([s2 / 2.0, s2 / 2.0])]
    elif k == 4:  # pyipm.py:1977-1978
        p = PolyProblem(2, [(-1.0, ((0, 2), (1, 1)))],
                        ce_terms=[[(1.0, ((0, 2),)), (1.0, ((1, 2),)), (-3.0, ())]])
        gt = [np.array([s2, 1.0]), np.array([-s2, 1.0]), np.array([0.0, -s3])]
    elif k == 5:  # pyipm.py:2000-2005
        p = PolyProblem(2, [(1.0, ((0, 2),)), (2.0, ((1, 2),)), (2.0, ((0, 1),)), (8.0, ((1, 1),))],
                        ci_terms=[_lin([1.0, 2.0], -10.0), _lin([1.0, 0.0]), _lin([0.0, 1.0])])
        gt = [np.array([4.0, 3.0])]
    elif k == 6:  # pyipm.py:2027-2029
        p = PolyProblem(6, [], ce_terms=[_lin([1.0] * 6, -1.0)],
                        ci_terms=[_lin([1.0 if j == i else 0.0 for j in range(6)]) for i in range(6)],
                        xlogx=(1.0, _EPS))
        gt = [np.array([1.0 / 6.0] * 6)]
    elif k == 7:  # pyipm.py:2049-2051  (BASELINE.json config 1)
        p = PolyProblem(3, [(-1.0, ((0, 1), (1, 1), (2, 1)))], ce_terms=[_lin([1.0, 1.0, 1.0], -1.0)],
                        ci_terms=[_lin([1.0 if j == i else 0.0 for j in range(3)]) for i in range(3)])
        gt = [np.array([1.0 / 3.0] * 3)]
    elif k == 8:  # pyipm.py:2070-2073
        p = PolyProblem(3, _lin([0.0, 4.0, -2.0]),
                        ce_terms=[_lin([2.0, -1.0, -1.0], -2.0), [(1.0, ((0, 2),)), (1.0, ((1, 2),)), (-1.0, ())]])
        gt = [np.array([2.0 / s13, -3.0 / s13, -2.0 + 7.0 / s13])]
    elif k == 9:  # pyipm.py:2095-2098  (x-2)^2 + 2(y-1)^2 expanded
        p = PolyProblem(2, [(1.0, ((0, 2),)), (-4.0, ((0, 1),)), (4.0, ()), (2.0, ((1, 2),)), (-4.0, ((1, 1),)),
                            (2.0, ())],
                        ci_terms=[_lin([-1.0, -4.0], 3.0), _lin([1.0, -1.0])])
        gt = [np.array([5.0 / 3.0, 1.0 / 3.0])]
    elif k == 10:  # pyipm.py:2116-2119 expanded
        p = PolyProblem(3, [(1.0, ((0, 2),)), (-2.0, ((0, 1),)), (1.0, ()),
                            (2.0, ((1, 2),)), (8.0, ((1, 1),)), (8.0, ()),
                            (3.0, ((2, 2),)), (18.0, ((2, 1),)), (27.0, ())],
                        ce_terms=[_lin([-1.0, -1.0, 1.0], -1.0)],
                        ci_terms=[[(1.0, ((2, 1),)), (-1.0, ((0, 2),))]])
        gt = [np.array([0.12288, -1.1078, 0.015100])]
    else:
        raise ValueError('example problems are numbered 1..10')
    p.name = 'example%d' % k
    return p, gt


def example_x0(k):
    """Seed-42 initial points in the draw order of unit_tests.py:8,103,118,...,234.
    Note p6 is *not* normalised in unit_tests.py (unlike pyipm.py:2025)."""
    rs = np.random.RandomState(42)
    draws = {}
    for i, (n, kind) in enumerate([(2, 'n'), (2, 'n'), (2, 'n'), (2, 'n'), (2, 'n'), (6, 'u'), (3, 'n'), (3, 'n'),
                                   (2, 'n'), (3, 'n')], start=1):
        draws[i] = rs.randn(n) if kind == 'n' else rs.rand(n)
    return draws[k].astype(np.float64)


# --------------------------------------------------------------------------------------------- QuadProblem
class QuadProblem(object):
    """Dense synthetic NLP family of BASELINE.json configs 2, 3, 5.

        f  = 1/2 x'Qx + c'x + q4/4 * sum(x^4)
        ce = A x + 1/2 (U x)^2 - b          (M rows;  U optional)
        ci = G x - 1/2 (V x)^2 + r          (N rows;  V optional)

    Jacobians follow the reference convention (D x M, D x N), so the device stores the *transposed* data
    ``At, Ut (D x M)`` and ``Gt, Vt (D x N)`` row-major, K-contiguous for the SYRK-shaped contractions:
        dce = At + Ut*diag(Ux),   dci = Gt - Vt*diag(Vx),
        d2L = Q + 3 q4 diag(x^2) - Ut diag(lda_e) Ut' + Vt diag(lda_i) Vt'.
    """

    def __init__(self, Q, c, q4=0.0, A=None, U=None, b=None, G=None, V=None, r=None, x0=None, name=None):
        self.Q = np.ascontiguousarray(Q, dtype=np.float64)
        self.c = np.ascontiguousarray(c, dtype=np.float64)
        self.q4 = float(q4)
        self.nvar = self.Q.shape[0]
        self.neq = 0 if A is None else A.shape[0]
        self.nineq = 0 if G is None else G.shape[0]
        D = self.nvar
        self.At = None if A is None else np.ascontiguousarray(A.T, dtype=np.float64)
        self.Ut = None if U is None else np.ascontiguousarray(U.T, dtype=np.float64)
        self.b = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
        self.Gt = None if G is None else np.ascontiguousarray(G.T, dtype=np.float64)
        self.Vt = None if V is None else np.ascontiguousarray(V.T, dtype=np.float64)
        self.r = None if r is None else np.ascontiguousarray(r, dtype=np.float64)
        assert self.c.shape == (D,)
        self.x0 = x0
        self.name = name

    def f(self, x):
        return np.float64(0.5 * np.dot(x, np.dot(self.Q, x)) + np.dot(self.c, x) + 0.25 * self.q4 * np.sum(x ** 4))

    def df(self, x):
        return np.dot(self.Q, x) + self.c + self.q4 * x ** 3

    def d2f(self, x):
        return self.Q + np.diag(3.0 * self.q4 * x ** 2)

    def ce(self, x):
        v = np.dot(x, self.At) - self.b
        if self.Ut is not None:
            ux = np.dot(x, self.Ut)
            v = v + 0.5 * ux * ux
        return v

    def dce(self, x):
        if self.Ut is None:
            return self.At
        return self.At + self.Ut * np.dot(x, self.Ut)[None, :]

    def d2ce(self, x, lda):
        if self.Ut is None:
            return np.zeros((self.nvar, self.nvar))
        return np.dot(self.Ut * lda[:self.neq][None, :], self.Ut.T)

    def ci(self, x):
        v = np.dot(x, self.Gt) + self.r
        if self.Vt is not None:
            vx = np.dot(x, self.Vt)
            v = v - 0.5 * vx * vx
        return v

    def dci(self, x):
        if self.Vt is None:
            return self.Gt
        return self.Gt - self.Vt * np.dot(x, self.Vt)[None, :]

    def d2ci(self, x, lda):
        if self.Vt is None:
            return np.zeros((self.nvar, self.nvar))
        return -np.dot(self.Vt * lda[self.neq:][None, :], self.Vt.T)

    def callables(self):
        kw = dict(f=self.f, df=self.df, d2f=self.d2f)
        if self.neq:
            kw.update(ce=self.ce, dce=self.dce, d2ce=self.d2ce)
        if self.nineq:
            kw.update(ci=self.ci, dci=self.dci, d2ci=self.d2ci)
        return kw


def make_qp(D=1024, M=256, nbox=512, seed=None):
    """BASELINE.json config 2 -- synthetic convex QP (SURVEY.md section 8d 'C2'); sizes scale for tests."""
    rng = np.random.default_rng(D if seed is None else seed)
    B = rng.standard_normal((D, D))
    Q = np.dot(B, B.T) / D + np.eye(D)
    c = rng.standard_normal(D)
    A = rng.standard_normal((M, D)) / np.sqrt(D)
    xs = rng.uniform(-0.5, 0.5, D)
    b = np.dot(A, xs)
    G = np.zeros((2 * nbox, D))
    G[np.arange(nbox), np.arange(nbox)] = 1.0
    G[nbox + np.arange(nbox), np.arange(nbox)] = -1.0
    r = np.ones(2 * nbox)
    x0 = xs + 0.1 * rng.standard_normal(D)
    x0[:nbox] = np.clip(x0[:nbox], -0.9, 0.9)
    return QuadProblem(Q, c, 0.0, A=A, b=b, G=G, r=r, x0=x0, name='qp_D%d_M%d_N%d' % (D, M, 2 * nbox))


def make_nlp(D=4096, M=512, N=4096, seed=None):
    """BASELINE.json config 3 -- synthetic nonconvex NLP (SURVEY.md section 8d 'C3'); sizes scale for tests."""
    rng = np.random.default_rng(D if seed is None else seed)
    sq = np.sqrt(D)
    B = rng.standard_normal((D, D))
    Q = (B + B.T) / (2.0 * sq)
    del B
    c = rng.standard_normal(D)
    A = rng.standard_normal((M, D)) / sq
    U = rng.standard_normal((M, D)) / sq
    G = rng.standard_normal((N, D)) / sq
    V = rng.standard_normal((N, D)) / sq
    xs = 0.5 * rng.standard_normal(D)
    ux = np.dot(U, xs)
    b = np.dot(A, xs) + 0.5 * ux * ux
    vx = np.dot(V, xs)
    r = np.abs(rng.standard_normal(N)) + 0.1 - (np.dot(G, xs) - 0.5 * vx * vx)
    x0 = xs + 0.05 * rng.standard_normal(D)
    return QuadProblem(Q, c, 1.0, A=A, U=U, b=b, G=G, V=V, r=r, x0=x0, name='nlp_D%d_M%d_N%d' % (D, M, N))


def make_rankdef_nlp(D=24, M=6, N=16, seed=41):
    """make_nlp with the LAST equality constraint an exact copy of the first: dce has two identical columns, the KKT
    matrix is exactly singular for every delta, so the reference's `rcond <= eps` branch (pyipm.py:1381-1389: eq-block
    regularisation -sqrt(eps)*eta*mu^beta*I) fires at every step."""
    p = make_nlp(D=D, M=M, N=N, seed=seed)
    A_, U_, b_ = p.At.T.copy(), p.Ut.T.copy(), p.b.copy()
    A_[M - 1], U_[M - 1], b_[M - 1] = A_[0], U_[0], b_[0]
    return QuadProblem(p.Q, p.c, p.q4, A=A_, U=U_, b=b_, G=p.Gt.T, V=p.Vt.T, r=p.r, x0=p.x0,
                       name='nlp_rankdef_D%d_M%d_N%d' % (D, M, N))


def mu_sweep_state(prob, mu, seed=5):
    """BASELINE.json config 5 -- ill-conditioned barrier state for a teacher-forced Newton step
    (SURVEY.md section 8d 'C5'): s_i*lda_i = mu, 25% 'active' rows with s_i = mu^0.9, rest U(0.1, 1)."""
    rng = np.random.default_rng(seed)
    N, M = prob.nineq, prob.neq
    s = rng.uniform(0.1, 1.0, N)
    act = rng.random(N) < 0.25
    s[act] = mu ** 0.9
    lda_i = mu / s
    lda_e = 0.1 * rng.standard_normal(M)
    return prob.x0.copy(), s, np.concatenate([lda_e, lda_i])


def make_dense_kkt(n=16384, m=2048, seed=None, delta_c=1e-8):
    """BASELINE.json config 4 -- dense symmetric quasi-definite KKT matrix given directly
    (SURVEY.md section 8d 'C4'): [[H, J], [J', -delta_c I]], H = WW'/n + diag(log-uniform 1e-4..1e4)."""
    rng = np.random.default_rng(n if seed is None else seed)
    nh = n - m
    W = rng.standard_normal((nh, nh))
    H = np.dot(W, W.T) / nh
    del W
    H[np.arange(nh), np.arange(nh)] += 10.0 ** rng.uniform(-4.0, 4.0, nh)
    J = rng.standard_normal((nh, m))
    K = np.zeros((n, n))
    K[:nh, :nh] = H
    K[:nh, nh:] = J
    K[nh:, :nh] = J.T
    K[np.arange(nh, n), np.arange(nh, n)] = -delta_c
    rhs = rng.standard_normal((n, 8))
    return K, rhs
