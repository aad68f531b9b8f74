"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.

NumPy/SciPy restatement of the exact-Hessian Newton-step path of jkaardal/pyipm
(`/root/reference/pyipm.py`, commit ccc74da).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module; the product
(`pyipm_b200/`) never does and fails loudly when its CUDA library is missing.

What is restated (reference file:line in every docstring below):
  * the *precompiled-function* assembly path of `IPM.compile()` (pyipm.py:512-562, 564-814), which is
    already plain NumPy in the reference except for six one-line Aesara expressions
    (pyipm.py:624-627, 673-675, 699-702, 749-751, 906-909, 911-914) that are restated here with the
    SciPy/NumPy calls Aesara forwards them to (`scipy.linalg.eigvalsh(a, b, lower=True)`,
    `scipy.linalg.solve(a, b, assume_a='gen')`);
  * `KKT` (958-991), `reghess` (1373-1406), `step` (1408-1436), `search` (1438-1565),
    `solve` (1567-1863).

Third-party dependency that really executes the arithmetic in the reference: `aesara>=2.2.6`
(setup.py:18-22, unpinned, not vendored, not installable here).  PARITY PINNING: this oracle is
pinned against outputs of the UNMODIFIED reference `pyipm.py` executed in the authoring container
through a minimal stand-in for the missing `aesara` import (`oracle/aesara_shim`, lazy expression
evaluator that forwards to the same SciPy/NumPy calls); the fixtures and the script that made them
are `tests/golden/*.npz` and `tests/golden/make_golden.py`.  `tests/test_oracle_golden.py` checks
this file against those fixtures and against the reference's own known answers
(unit_tests.py:104-235, README.md:115-121).

Callable conventions (identical to the reference's "precompiled function" input mode,
pyipm.py:216-231): f(x)->scalar, df(x)->(D,), d2f(x)->(D,D), ce(x)->(M,), dce(x)->(D,M),
d2ce(x, lda)->(D,D) [lda is the FULL multiplier vector of size M+N], ci(x)->(N,), dci(x)->(D,N),
d2ci(x, lda)->(D,D).
"""
from __future__ import print_function

import numpy as np
import scipy.linalg


class OracleIPM(object):
    """Restatement of `class IPM` (pyipm.py:23-1863): exact-Hessian branch and the L-BFGS branch (lbfgs=m)."""

    def __init__(self, x0=None, f=None, df=None, d2f=None, ce=None, dce=None, d2ce=None, ci=None, dci=None,
                 d2ci=None, lda0=None, s0=None, mu=0.2, nu=10.0, rho=0.1, tau=0.995, eta=1.0E-4,
                 beta=0.4, miter=20, niter=10, Xtol=None, Ktol=1.0E-4, Ftol=None, lbfgs=False, lbfgs_zeta=None,
                 float_dtype=np.float64, verbosity=-1, trace=None):
        # pyipm.py:311-376
        self.x0 = x0
        self.lda0 = lda0
        self.s0 = s0
        self.f = f
        self.df = df
        self.d2f = d2f
        self.ce = ce
        self.dce = dce
        self.d2ce = d2ce
        self.ci = ci
        self.dci = dci
        self.d2ci = d2ci
        self.nvar = None
        self.neq = None
        self.nineq = None
        self.eps = np.finfo(float_dtype).eps
        self.mu = mu
        self.nu = nu
        self.rho = rho
        self.tau = tau
        self.eta = eta
        self.beta = beta
        self.miter = miter
        self.niter = niter
        if Xtol:
            self.Xtol = Xtol
        else:
            self.Xtol = self.eps
        self.Ktol = Ktol
        self.Ftol = Ftol
        self.reg_coef = float_dtype(np.sqrt(self.eps))
        # pyipm.py:355-360
        self.lbfgs = lbfgs
        if self.lbfgs and lbfgs_zeta is None:
            self.lbfgs_zeta = float_dtype(1.0)
        else:
            self.lbfgs_zeta = lbfgs_zeta
        self.lbfgs_fail_max = lbfgs
        self.float_dtype = float_dtype
        # the two Aesara shared scalars (pyipm.py:363-364); kept separate from mu_host/nu_host on purpose
        # (quirk xi: mu_dev is not reset by a second solve() when nineq>0, pyipm.py:1603 vs 1607)
        self.nu_dev = self.float_dtype(self.nu)
        self.mu_dev = self.float_dtype(self.mu)
        self.verbosity = verbosity
        self.delta0 = self.reg_coef
        self.compiled = False
        # optional per-Newton-step recorder: list that receives one dict per inner iteration
        self.trace = trace
        # phase timers (seconds) for the CPU baseline split (bench.py); keys grad/hess/reghess/solve/search/kkt
        self.timers = None

    # ------------------------------------------------------------------ validate / compile
    def validate(self):
        """pyipm.py:385-408 (x_dev assertion dropped: there is no symbolic variable)."""
        assert self.f is not None
        assert (self.ce is not None) or (self.ce is None and self.dce is None and self.d2ce is None)
        assert (self.ci is not None) or (self.ci is None and self.dci is None and self.d2ci is None)
        assert self.mu > 0.0
        assert self.nu > 0.0
        assert 0.0 < self.eta < 1.0
        assert 0.0 < self.rho < 1.0
        assert 0.0 < self.tau < 1.0
        assert self.beta < 1.0
        assert self.miter >= 0 and isinstance(self.miter, int)
        assert self.niter >= 0 and isinstance(self.miter, int)
        assert self.Xtol >= self.eps
        assert self.Ktol >= self.eps
        assert self.Ftol is None or self.Ftol >= 0.0
        assert self.lbfgs >= 0 or self.lbfgs == False  # noqa: E712  (pyipm.py:405-408)
        if self.lbfgs:
            assert isinstance(self.lbfgs, int)
        assert self.lbfgs_zeta is None or self.lbfgs_zeta > 0.0

    def compile(self, nvar=None, neq=None, nineq=None):
        """Size discovery pyipm.py:414-467, then the NumPy lambdas of the precompile path."""
        if nvar is not None:
            self.nvar = nvar
        self.neq = neq
        self.nineq = nineq
        if self.ce is not None and self.neq is None:
            self.neq = np.asarray(self.ce(self.x0)).size
        elif neq is None:
            self.neq = 0
        if self.ci is not None and self.nineq is None:
            self.nineq = np.asarray(self.ci(self.x0)).size
        elif nineq is None:
            self.nineq = 0

        nvar, neq, nineq = self.nvar, self.neq, self.nineq
        eps = self.eps
        f_func, df_func, d2f_func = self.f, self.df, self.d2f

        # pyipm.py:527-562
        if neq:
            ce_func = self.ce
            dce_raw, d2ce_func = self.dce, self.d2ce

            def dce_func(x):
                return dce_raw(x).reshape((nvar, neq))
        if nineq:
            ci_raw, dci_raw, d2ci_func = self.ci, self.dci, self.d2ci

            def ci_func(x, s):
                return ci_raw(x) - s

            def dci_func(x):
                return dci_raw(x).reshape((nvar, nineq))

        # constraints, pyipm.py:565-573
        if neq or nineq:
            if neq and nineq:
                def con(x, s):
                    return np.concatenate([ce_func(x).reshape((neq,)), ci_func(x, s).reshape((nineq,))], axis=0)
            elif neq:
                def con(x, s):
                    return ce_func(x).reshape((neq,))
            else:
                def con(x, s):
                    return ci_func(x, s).reshape((nineq,))
            self.con = con

        # constraint Jacobian, pyipm.py:582-599
        if neq or nineq:
            if neq and nineq:
                def jaco_top(x):
                    return np.concatenate([dce_func(x).reshape((nvar, neq)),
                                           dci_func(x).reshape((nvar, nineq))], axis=1)
                jaco_bottom = np.concatenate([np.zeros((nineq, neq)), -np.eye(nineq)], axis=1)

                def jaco(x):
                    return np.concatenate([jaco_top(x), jaco_bottom], axis=0)
            elif neq:
                def jaco(x):
                    return dce_func(x).reshape((nvar, neq))
            else:
                def jaco(x):
                    return np.concatenate([dci_func(x).reshape((nvar, nineq)), -np.eye(nineq)], axis=0)
            self.jaco = jaco

        # gradient, pyipm.py:610-653
        if neq and nineq:
            def grad_x(x, lda):
                return df_func(x) - np.dot(dce_func(x), lda[:neq]) - np.dot(dci_func(x), lda[neq:])
        elif neq:
            def grad_x(x, lda):
                return df_func(x) - np.dot(dce_func(x), lda)
        elif nineq:
            def grad_x(x, lda):
                return df_func(x) - np.dot(dci_func(x), lda)
        else:
            def grad_x(x, lda):
                return df_func(x)

        if nineq:
            def grad_s(x, s, lda):
                # Aesara one-liner pyipm.py:624-627 (uses the shared mu_dev)
                return lda[neq:] - self.mu_dev / (s + eps)
        if neq:
            def grad_lda_eq(x):
                return ce_func(x).ravel()
        if nineq:
            def grad_lda_ineq(x, s):
                return ci_func(x, s).ravel()

        if neq and nineq:
            def grad(x, s, lda):
                return np.concatenate([grad_x(x, lda), grad_s(x, s, lda), grad_lda_eq(x), grad_lda_ineq(x, s)],
                                      axis=0)
        elif neq:
            def grad(x, s, lda):
                return np.concatenate([grad_x(x, lda), grad_lda_eq(x)], axis=0)
        elif nineq:
            def grad(x, s, lda):
                return np.concatenate([grad_x(x, lda), grad_s(x, s, lda), grad_lda_ineq(x, s)], axis=0)
        else:
            def grad(x, s, lda):
                return grad_x(x, lda)
        self.grad = grad

        # merit function, pyipm.py:671-686 (log(s) WITHOUT eps: quirk i)
        if nineq:
            def bar_func(s):
                return self.mu_dev * np.sum(np.log(s))
        if neq and nineq:
            def phi(x, s):
                return (f_func(x) + self.nu_dev * (np.sum(np.abs(ce_func(x))) + np.sum(np.abs(ci_func(x, s)))) -
                        bar_func(s))
        elif neq:
            def phi(x, s):
                return f_func(x) + self.nu_dev * np.sum(np.abs(ce_func(x)))
        elif nineq:
            def phi(x, s):
                return f_func(x) + self.nu_dev * np.sum(np.abs(ci_func(x, s))) - bar_func(s)
        else:
            def phi(x, s):
                return f_func(x)
        self.phi = phi

        # merit directional derivative, pyipm.py:697-714
        if nineq:
            def dbar_func(s, dz):
                return np.dot(self.mu_dev / (s + eps), dz[nvar:])
        if neq and nineq:
            def dphi(x, s, dz):
                return (np.dot(df_func(x), dz[:nvar]) - self.nu_dev *
                        (np.sum(np.abs(ce_func(x))) + np.sum(np.abs(ci_func(x, s)))) - dbar_func(s, dz))
        elif neq:
            def dphi(x, s, dz):
                return np.dot(df_func(x), dz[:nvar]) - self.nu_dev * np.sum(np.abs(ce_func(x)))
        elif nineq:
            def dphi(x, s, dz):
                return (np.dot(df_func(x), dz[:nvar]) - self.nu_dev * np.sum(np.abs(ci_func(x, s))) -
                        dbar_func(s, dz))
        else:
            def dphi(x, s, dz):
                return np.dot(df_func(x), dz[:nvar])
        self.dphi = dphi

        # multiplier / slack initialisation, pyipm.py:724-739
        if neq or nineq:
            def init_lambda(x):
                return np.dot(np.linalg.pinv(jaco(x)[:nvar, :]),
                              df_func(x).reshape((nvar, 1))).reshape((neq + nineq,))
            self.init_lambda = init_lambda
        if nineq:
            def init_slack(x):
                return np.max(np.concatenate([
                    ci_func(x, np.zeros((nineq,))).reshape((nineq, 1)),
                    self.Ktol * np.ones((nineq, 1))
                ], axis=1), axis=1)
            self.init_slack = init_slack

        # gradient of f (+ barrier), pyipm.py:747-756
        if nineq:
            def barrier_cost_grad(x, s):
                return np.concatenate([df_func(x), -self.mu_dev / (s + eps)], axis=0)
        else:
            def barrier_cost_grad(x, s):
                return df_func(x)
        self.barrier_cost_grad = barrier_cost_grad

        # Lagrangian Hessian and full KKT matrix, pyipm.py:768-814 (only triu(d2L) is used: quirk ii); in L-BFGS mode
        # (pyipm.py:767) these slots exist but are never called (d2f/d2ce/d2ci may be None)
        if neq and nineq:
            def d2L(x, lda):
                return d2f_func(x) - d2ce_func(x, lda) - d2ci_func(x, lda)
        elif neq:
            def d2L(x, lda):
                return d2f_func(x) - d2ce_func(x, lda)
        elif nineq:
            def d2L(x, lda):
                return d2f_func(x) - d2ci_func(x, lda)
        else:
            def d2L(x, lda):
                return d2f_func(x)
        self.d2L = d2L

        if neq or nineq:
            if nineq:
                def hess_upper_left(x, s, lda):
                    return np.concatenate([
                        np.concatenate([np.triu(d2L(x, lda)), np.zeros((nvar, nineq))], axis=1),
                        np.concatenate([np.zeros((nineq, nvar)), np.diag(lda[neq:] / (s + eps))], axis=1)
                    ], axis=0)
            else:
                def hess_upper_left(x, s, lda):
                    return np.triu(d2L(x, lda))

            def hess_triu(x, s, lda):
                upper = np.concatenate([hess_upper_left(x, s, lda), jaco(x)], axis=1)
                return np.concatenate([upper, np.zeros((neq + nineq, nvar + 2 * nineq + neq))], axis=0)
        else:
            def hess_triu(x, s, lda):
                return np.triu(d2L(x, lda))

        def hess(x, s, lda):
            # the reference evaluates hess_triu twice (pyipm.py:809-810); once is arithmetically identical
            ht = hess_triu(x, s, lda)
            return ht + np.triu(ht, k=1).T
        self.hess = hess

        self.cost = f_func
        # Aesara one-liners pyipm.py:906-909 and 911-914
        self.eigh = lambda M: scipy.linalg.eigvalsh(M, np.eye(M.shape[0]), lower=True)
        self.sym_solve_cmp = lambda M, b: scipy.linalg.solve(M, b, assume_a='gen')
        self.compiled = True

    # ------------------------------------------------------------------ KKT
    def KKT(self, x, s, lda):
        """pyipm.py:958-991; kkt2 is g_s * s (quirk xii); absent conditions are scalar 0.0."""
        kkts = self.grad(x, s, lda)
        nvar, neq, nineq = self.nvar, self.neq, self.nineq
        if neq and nineq:
            kkt1 = kkts[:nvar]
            kkt2 = kkts[nvar:(nvar + nineq)] * s
            kkt3 = kkts[(nvar + nineq):(nvar + nineq + neq)]
            kkt4 = kkts[(nvar + nineq + neq):]
        elif neq:
            kkt1 = kkts[:nvar]
            kkt2 = self.float_dtype(0.0)
            kkt3 = kkts[(nvar + nineq):(nvar + nineq + neq)]
            kkt4 = self.float_dtype(0.0)
        elif nineq:
            kkt1 = kkts[:nvar]
            kkt2 = kkts[nvar:(nvar + nineq)] * s
            kkt3 = self.float_dtype(0.0)
            kkt4 = kkts[(nvar + nineq + neq):]
        else:
            kkt1 = kkts[:nvar]
            kkt2 = self.float_dtype(0.0)
            kkt3 = self.float_dtype(0.0)
            kkt4 = self.float_dtype(0.0)
        return kkt1, kkt2, kkt3, kkt4


    # ------------------------------------------------------------------ L-BFGS (pyipm.py:993-1371)
    def lbfgs_init(self):
        """pyipm.py:993-1005."""
        zeta = self.float_dtype(self.lbfgs_zeta)
        S = np.array([], dtype=self.float_dtype).reshape((self.nvar, 0))
        Y = np.array([], dtype=self.float_dtype).reshape((self.nvar, 0))
        SS = np.array([], dtype=self.float_dtype).reshape((0, 0))
        L = np.array([], dtype=self.float_dtype).reshape((0, 0))
        D = np.array([], dtype=self.float_dtype).reshape((0, 0))
        return zeta, S, Y, SS, L, D, 0

    def lbfgs_dir(self, x, s, lda, g, zeta, S, Y, SS, L, D):
        """pyipm.py:1184-1246 with the graph of lbfgs_builder (pyipm.py:1007-1182) evaluated eagerly in NumPy/SciPy:
        compact representation + Woodbury for constrained problems (general, non-reduced variant pyipm.py:1099-1148),
        two-loop-free compact inverse for unconstrained ones (pyipm.py:1149-1175).  The square-Jacobian reduced variant
        (pyipm.py:1061-1097) cannot be compiled by the reference itself (duplicate `s_dev` in the inputs of
        `lbfgs_dir_func_sqr`, pyipm.py:877-878) and is not restated; `solve` here is scipy.linalg.solve(assume_a='gen')
        as in sym_solve (pyipm.py:18-20), `eigh` is numpy.linalg.eigh (aesara nlinalg.eigh)."""
        nvar, neq, nineq, eps = self.nvar, self.neq, self.nineq, self.eps
        solve = lambda A, b: scipy.linalg.solve(A, b, assume_a='gen')  # noqa: E731
        m = S.shape[1]
        if neq or nineq:
            B = self.jaco(x)
            Adiag = zeta * np.ones((nvar, 1))
            if nineq:
                Sigma = (lda[neq:] / (s + eps)).reshape((nineq, 1))
                Adiag = np.concatenate([Adiag, Sigma], axis=0)
            BT_invA = np.dot(B.T, np.diag(1.0 / Adiag.reshape((Adiag.size,))))
            BT_invA_B = np.dot(BT_invA, B)
            self.last_lbfgs = {'eq_reg': False, 'm': m}
            if neq:
                w = np.linalg.eigh(BT_invA_B[:neq, :neq])[0]
                rcond = np.min(np.abs(w)) / np.max(np.abs(w))
                self.last_lbfgs['rcond'] = float(rcond)
                if rcond <= eps:
                    BT_invA_B = np.array(BT_invA_B, copy=True)
                    BT_invA_B[:neq, :neq] += self.reg_coef * self.eta * (self.mu_dev ** self.beta) * np.eye(neq)
                    self.last_lbfgs['eq_reg'] = True
            gp = g[:nvar + nineq].reshape((nvar + nineq, 1))
            gd = g[nvar + nineq:].reshape((neq + nineq, 1))
            v00 = np.dot(BT_invA, gp)
            v01 = solve(BT_invA_B, v00)
            v02 = gp / Adiag - np.dot(BT_invA.T, v01)
            v03 = -solve(BT_invA_B, gd)
            v04 = -np.dot(BT_invA.T, v03)
            Zg = np.concatenate([v02 + v04, v01 + v03], axis=0)
            if m > 0:
                W = np.concatenate([zeta * S, Y], axis=1)
                if nineq:
                    W = np.concatenate([W, np.zeros((nineq, 2 * m))], axis=0)
                BT_gmaW = np.dot(B.T, W) / zeta
                X00 = -solve(BT_invA_B, BT_gmaW)
                X01 = W / zeta + np.dot(BT_invA.T, X00)
                X02 = np.dot(W.T, X01)
                M0 = np.concatenate([zeta * SS, L], axis=1)
                M1 = np.concatenate([L.T, -D], axis=1)
                Minv = np.concatenate([M0, M1], axis=0)
                v10 = np.dot(W.T, Zg[:nvar + nineq])
                v11 = solve(X02 - Minv, v10)
                X10 = np.concatenate([X01, -X00], axis=0)
                dz = Zg - np.dot(X10, v11)
            else:
                dz = Zg
        else:
            Hg = zeta * g.reshape((nvar, 1))
            if m > 0:
                W = np.concatenate([S, zeta * Y], axis=1)
                WT_g = np.dot(W.T, g)
                Bv = -solve(L, WT_g[:m].reshape((m, 1)))
                Av = (-solve(L.T, np.dot(D + zeta * SS, Bv)) - solve(L.T, WT_g[m:].reshape((m, 1))))
                dz = Hg + np.dot(W, np.concatenate([Av, Bv], axis=0))
            else:
                dz = Hg
        return dz.reshape((dz.size,))

    def lbfgs_update(self, x_old, x_new, g_old, g_new, zeta, S, Y, SS, L, D, lbfgs_fail):
        """pyipm.py:1282-1371 (the curvature perturbation is commented out in the reference)."""
        nvar = self.nvar
        dx = x_new - x_old
        dg = g_old[:nvar] - g_new[:nvar]
        if self.neq or self.nineq:
            zeta_new = np.dot(dg, dx) / (np.dot(dx, dx) + self.eps)
        else:
            zeta_new = np.dot(dg, dx) / (np.dot(dg, dg) + self.eps)
        if np.dot(dx, dg) > np.sqrt(self.eps) and zeta_new > np.sqrt(self.eps):
            zeta = zeta_new
            if S.shape[1] > self.lbfgs:
                S[:, :-1] = S[:, 1:]
                Y[:, :-1] = Y[:, 1:]
                SS[:-1, :-1] = SS[1:, 1:]
                L[:-1, :-1] = L[1:, 1:]
                D[:-1, :-1] = D[1:, 1:]
            else:
                lsize = S.shape[1] + 1
                S = np.concatenate([S, np.zeros((nvar, 1), dtype=self.float_dtype)], axis=1)
                Y = np.concatenate([Y, np.zeros((nvar, 1), dtype=self.float_dtype)], axis=1)
                SS = np.concatenate([SS, np.zeros((1, lsize - 1), dtype=self.float_dtype)], axis=0)
                SS = np.concatenate([SS, np.zeros((lsize, 1), dtype=self.float_dtype)], axis=1)
                L = np.concatenate([L, np.zeros((1, lsize - 1), dtype=self.float_dtype)], axis=0)
                L = np.concatenate([L, np.zeros((lsize, 1), dtype=self.float_dtype)], axis=1)
                D = np.concatenate([D, np.zeros((1, lsize - 1), dtype=self.float_dtype)], axis=0)
                D = np.concatenate([D, np.zeros((lsize, 1), dtype=self.float_dtype)], axis=1)
            S[:, -1] = dx
            Y[:, -1] = dg
            if self.neq or self.nineq:
                SS_update = np.dot(S.T, dx.reshape((nvar, 1)))
            else:
                SS_update = np.dot(Y.T, dg.reshape((nvar, 1)))      # this is YY for unconstrained problems
            SS[:, -1] = SS_update.reshape((SS_update.size,))
            SS[-1, :] = SS_update.reshape((SS_update.size,))
            lsize = SS.shape[1]
            SS = SS.reshape((lsize, lsize))
            if self.neq or self.nineq:
                L_update = np.dot(dx.reshape((1, nvar)), Y)
                L[-1, :] = L_update
                L[-1, -1] = self.float_dtype(0.0)
            else:
                L_update = np.dot(S.T, dg.reshape((nvar, 1))).reshape((S.shape[1],))   # this is R
                L[:, -1] = L_update
            L = L.reshape((lsize, lsize))
            D[-1, -1] = np.dot(dx, dg)
            D = D.reshape((lsize, lsize))
            lbfgs_fail = 0
        else:
            lbfgs_fail += 1
        if lbfgs_fail > self.lbfgs_fail_max and S.shape[1] > 0:
            zeta, S, Y, SS, L, D, lbfgs_fail = self.lbfgs_init()
        return zeta, S, Y, SS, L, D, lbfgs_fail

    # ------------------------------------------------------------------ reghess
    def reghess(self, Hc):
        """pyipm.py:1373-1406.  Records (rcond, n_neg first/last, retries) in self.last_reg."""
        w = self.eigh(Hc)
        rcond = np.min(np.abs(w)) / np.max(np.abs(w))
        n_eig = 1
        info = {'rcond': float(rcond), 'nneg0': int(np.sum(w < -self.eps)), 'eq_reg': False, 'triggered': False}
        if rcond <= self.eps or (self.neq + self.nineq) != np.sum(w < -self.eps):
            info['triggered'] = True
            if rcond <= self.eps and self.neq:
                ind1 = self.nvar + self.nineq
                ind2 = ind1 + self.neq
                Hc[ind1:ind2, ind1:ind2] -= self.reg_coef * self.eta * (self.mu_host ** self.beta) * np.eye(self.neq)
                info['eq_reg'] = True
            if self.delta == 0.0:
                self.delta = self.delta0
            else:
                self.delta = np.max([self.delta / 2, self.delta0])
            Hc[:self.nvar, :self.nvar] += self.delta * np.eye(self.nvar)
            w = self.eigh(Hc)
            n_eig += 1
            while (self.neq + self.nineq) != np.sum(w < -self.eps):
                Hc[:self.nvar, :self.nvar] -= self.delta * np.eye(self.nvar)
                self.delta *= 10.0
                Hc[:self.nvar, :self.nvar] += self.delta * np.eye(self.nvar)
                w = self.eigh(Hc)
                n_eig += 1
        info['n_eig'] = n_eig
        info['delta'] = float(self.delta)
        self.last_reg = info
        return Hc

    # ------------------------------------------------------------------ step
    def step(self, x, dx):
        """pyipm.py:1408-1436: golden-section fraction-to-the-boundary search; returns bracket `a`."""
        GOLD = (np.sqrt(5.0) + 1.0) / 2.0
        a = 0.0
        b = 1.0
        if np.all(x + b * dx >= (1.0 - self.tau) * x):
            return b
        else:
            c = b - (b - a) / GOLD
            d = a + (b - a) / GOLD
            while np.abs(b - a) > GOLD * self.Xtol:
                if np.any(x + d * dx < (1.0 - self.tau) * x):
                    b = np.copy(d)
                else:
                    a = np.copy(d)
                if c > a:
                    if np.any(x + c * dx < (1.0 - self.tau) * x):
                        b = np.copy(c)
                    else:
                        a = np.copy(c)
                c = b - (b - a) / GOLD
                d = a + (b - a) / GOLD
            return a

    # ------------------------------------------------------------------ search
    def search(self, x0, s0, lda0, dz, alpha_smax, alpha_lmax):
        """pyipm.py:1438-1565.  Records n_backtracks / SOC flags in self.last_search."""
        nvar, neq, nineq = self.nvar, self.neq, self.nineq
        dx = dz[:nvar]
        if nineq:
            ds = dz[nvar:(nvar + nineq)]
        if neq or nineq:
            dl = dz[(nvar + nineq):]
        else:
            dl = self.float_dtype(0.0)
            alpha_lmax = self.float_dtype(0.0)

        x = np.copy(x0)
        s = np.copy(s0)
        phi0 = self.phi(x0, s0)
        dphi0 = self.dphi(x0, s0, dz[:nvar + nineq])
        correction = False
        info = {'phi0': float(phi0), 'dphi0': float(dphi0), 'n_backtracks': 0, 'soc_tried': False,
                'soc_accepted': False, 'armijo_first': True}
        self.last_search = info
        if nineq:
            if self.phi(x0 + alpha_smax * dx, s0 + alpha_smax * ds) > phi0 + alpha_smax * self.eta * dphi0:
                info['armijo_first'] = False
                c_old = self.con(x0, s0)
                c_new = self.con(x0 + alpha_smax * dx, s0 + alpha_smax * ds)
                if np.sum(np.abs(c_new)) > np.sum(np.abs(c_old)):
                    info['soc_tried'] = True
                    A = self.jaco(x0).T
                    try:
                        # quirk vii: reshape only succeeds when neq == nvar, otherwise ValueError -> lstsq
                        dz_p = -self.sym_solve_cmp(
                            A, c_new.reshape((nvar + nineq, 1))
                        ).reshape((nvar + nineq,))
                    except:  # noqa: E722  (bare except is the reference's, pyipm.py:1475)
                        dz_p = -np.linalg.lstsq(A, c_new, rcond=None)[0]
                    if (self.phi(x0 + alpha_smax * dx + dz_p[:nvar], s0 + alpha_smax * ds + dz_p[nvar:]) <=
                            phi0 + alpha_smax * self.eta * dphi0):
                        alpha_corr = self.step(s0, alpha_smax * ds + dz_p[nvar:])
                        if (self.phi(x0 + alpha_corr * (alpha_smax * dx + dz_p[:nvar]),
                                     s0 + alpha_corr * (alpha_smax * ds + dz_p[nvar:])) <=
                                phi0 + alpha_smax * self.eta * dphi0):
                            correction = True
                            info['soc_accepted'] = True
                if not correction:
                    alpha_smax *= self.tau
                    alpha_lmax *= self.tau
                    info['n_backtracks'] += 1
                    while self.phi(x0 + alpha_smax * dx, s0 + alpha_smax * ds) > phi0 + alpha_smax * self.eta * dphi0:
                        # quirk viii: alpha_lmax * ds mixed with alpha_smax * dx (pyipm.py:1496)
                        if (np.sqrt(np.linalg.norm(alpha_smax * dx) ** 2 + np.linalg.norm(alpha_lmax * ds) ** 2) <
                                self.eps):
                            self.signal = -2
                            return x0, s0, lda0
                        alpha_smax *= self.tau
                        alpha_lmax *= self.tau
                        info['n_backtracks'] += 1
            if correction:
                s = s0 + alpha_corr * (alpha_smax * ds + dz_p[nvar:])
            else:
                s = s0 + alpha_smax * ds
        else:
            if self.phi(x0 + alpha_smax * dx, s0) > phi0 + alpha_smax * self.eta * dphi0:
                info['armijo_first'] = False
                if neq:
                    c_old = self.con(x0, s0)
                    c_new = self.con(x0 + alpha_smax * dx, s0)
                    if np.sum(np.abs(c_new)) > np.sum(np.abs(c_old)):
                        info['soc_tried'] = True
                        A = self.jaco(x0).T
                        try:
                            # quirk vi: this reshape always raises (nineq == 0) -> always lstsq
                            dz_p = -self.sym_solve_cmp(
                                A, c_new.reshape((nvar, nineq, 1))
                            ).reshape((nvar + nineq,))
                        except:  # noqa: E722  (pyipm.py:1527)
                            dz_p = -np.linalg.lstsq(A, c_new, rcond=None)[0]
                        if self.phi(x0 + alpha_smax * dx + dz_p, s0) <= phi0 + alpha_smax * self.eta * dphi0:
                            alpha_corr = self.float_dtype(1.0)
                            correction = True
                            info['soc_accepted'] = True
                if not correction:
                    alpha_smax *= self.tau
                    alpha_lmax *= self.tau
                    info['n_backtracks'] += 1
                    while self.phi(x0 + alpha_smax * dx, s0) > phi0 + alpha_smax * self.eta * dphi0:
                        if np.linalg.norm(alpha_smax * dx) < self.eps:
                            self.signal = -2
                            return x0, s0, lda0
                        alpha_smax *= self.tau
                        alpha_lmax *= self.tau
                        info['n_backtracks'] += 1
        if correction:
            x = x0 + alpha_corr * (alpha_smax * dx + dz_p[:nvar])
        else:
            x = x0 + alpha_smax * dx
        if neq or nineq:
            lda = lda0 + alpha_lmax * dl
        else:
            lda = np.copy(lda0)
        info['alpha_s'] = float(alpha_smax)
        info['alpha_l'] = float(alpha_lmax)
        return x, s, lda

    # ------------------------------------------------------------------ one Newton step (pyipm.py:1714-1754)
    def newton_step(self, x, s, lda):
        """Body of one inner iteration, pyipm.py:1714-1754, factored out so that tests can teacher-force it
        from an arbitrary state (x, s, lda, mu_host/mu_dev, nu_host/nu_dev, delta)."""
        import time
        nvar, neq, nineq = self.nvar, self.neq, self.nineq
        tm = self.timers
        t0 = time.perf_counter()
        if self.lbfgs:
            # pyipm.py:1702-1713
            lb = self._lb
            if lb['not_first']:
                g_old = -self.grad(lb['x_old'], s, lda)
                g_new = -self.grad(x, s, lda)
                (lb['zeta'], lb['S'], lb['Y'], lb['SS'], lb['L'], lb['D'], lb['fail']) = self.lbfgs_update(
                    lb['x_old'], x, g_old, g_new, lb['zeta'], lb['S'], lb['Y'], lb['SS'], lb['L'], lb['D'], lb['fail'])
                lb['x_old'] = np.copy(x)
                lb['g'] = np.copy(g_new)
            g = lb['g']
            t1 = t2 = t3 = time.perf_counter()
            dz = self.lbfgs_dir(x, s, lda, g, lb['zeta'], lb['S'], lb['Y'], lb['SS'], lb['L'], lb['D'])
            self.last_reg = {'n_eig': 0, 'delta': float(self.delta), 'lbfgs_m': int(lb['S'].shape[1]),
                             'lbfgs_fail': int(lb['fail']), 'zeta': float(lb['zeta'])}
            t4 = time.perf_counter()
        else:
            g = -self.grad(x, s, lda)
            t1 = time.perf_counter()
            H = self.hess(x, s, lda)
            t2 = time.perf_counter()
            Hc = self.reghess(H)
            t3 = time.perf_counter()
            dz = self.sym_solve_cmp(Hc, g.reshape((g.size, 1))).reshape((g.size,))
            t4 = time.perf_counter()
        if neq or nineq:
            dz[nvar + nineq:] = -dz[nvar + nineq:]
        if neq or nineq:
            nu_thres = np.dot(self.barrier_cost_grad(x, s), dz[:nvar + nineq]) / (1 - self.rho) / \
                np.sum(np.abs(self.con(x, s)))
            if self.nu_host < nu_thres:
                self.nu_host = self.float_dtype(nu_thres)
                self.nu_dev = self.nu_host
        rec = None
        if self.trace is not None:
            rec = {'x': np.copy(x), 's': np.copy(s), 'lda': np.copy(lda), 'mu': float(self.mu_dev),
                   'mu_host': float(self.mu_host), 'nu': float(self.nu_dev), 'g': np.copy(g),
                   'dz': np.copy(dz), 'delta': float(self.delta), 'reg': dict(self.last_reg)}
        if nineq:
            alpha_smax = self.step(s, dz[nvar:(nvar + nineq)])
            alpha_lmax = self.step(lda[neq:], dz[(nvar + nineq + neq):])
            if rec is not None:
                rec['alpha_smax'] = float(alpha_smax)
                rec['alpha_lmax'] = float(alpha_lmax)
            x, s, lda = self.search(x, s, lda, dz, self.float_dtype(alpha_smax), self.float_dtype(alpha_lmax))
        else:
            x, s, lda = self.search(x, s, lda, dz, self.float_dtype(1.0), self.float_dtype(1.0))
        t5 = time.perf_counter()
        kkt = self.KKT(x, s, lda)
        t6 = time.perf_counter()
        if rec is not None:
            rec['search'] = dict(self.last_search)
            rec['x_new'] = np.copy(x)
            rec['s_new'] = np.copy(s)
            rec['lda_new'] = np.copy(lda)
            rec['nu_after'] = float(self.nu_dev)
            rec['kkt_norms'] = np.array([np.linalg.norm(k) for k in kkt])
            self.trace.append(rec)
        if tm is not None:
            for key, dt in (('grad', t1 - t0), ('hess', t2 - t1), ('reghess', t3 - t2), ('solve', t4 - t3),
                            ('search', t5 - t4), ('kkt', t6 - t5)):
                tm[key] = tm.get(key, 0.0) + dt
            tm['steps'] = tm.get('steps', 0) + 1
        return x, s, lda, kkt

    # ------------------------------------------------------------------ solve
    def solve(self, x0=None, s0=None, lda0=None, force_recompile=False):
        """pyipm.py:1567-1863 (prints reduced to the final summary)."""
        if x0 is not None:
            self.x0 = x0
        if s0 is not None:
            self.s0 = s0
        if lda0 is not None:
            self.lda0 = lda0
        assert (self.x0 is not None) and (self.x0.size > 0)
        assert self.x0.size == self.x0.shape[0]
        self.nvar = self.x0.size
        self.x0 = self.float_dtype(self.x0)
        self.validate()
        if not self.compiled or force_recompile:
            self.compile()
        nvar, neq, nineq = self.nvar, self.neq, self.nineq

        x = self.x0
        if nineq:
            if self.s0 is None:
                s = self.init_slack(x)
            else:
                s = self.s0.astype(self.float_dtype)
            self.mu_host = self.mu          # quirk xi: mu_dev intentionally NOT reset here
        else:
            s = np.array([], dtype=self.float_dtype)
            self.mu_host = self.Ktol        # quirk x
            self.mu_dev = self.float_dtype(self.mu_host)

        if neq or nineq:
            self.nu_host = self.nu
            self.nu_dev = self.float_dtype(self.nu_host)
            if self.lda0 is None:
                lda = self.init_lambda(x)
                if nineq and neq:
                    lda_ineq = lda[neq:]
                    lda_ineq[lda_ineq < self.float_dtype(0.0)] = self.float_dtype(self.Ktol)
                    lda[neq:] = lda_ineq
                elif nineq:
                    lda[lda < self.float_dtype(0.0)] = self.float_dtype(self.Ktol)
            else:
                lda = self.lda0.astype(self.float_dtype)
        else:
            lda = np.array([], dtype=self.float_dtype)

        self.delta = self.float_dtype(0.0)
        kkt = self.KKT(x, s, lda)
        self.init_state = (np.copy(x), np.copy(s), np.copy(lda))
        if self.lbfgs:
            # pyipm.py:1633-1637
            zeta, S, Y, SS, L, D, lbfgs_fail = self.lbfgs_init()
            self._lb = {'zeta': zeta, 'S': S, 'Y': Y, 'SS': SS, 'L': L, 'D': D, 'fail': lbfgs_fail,
                        'x_old': np.copy(x), 'g': -self.grad(x, s, lda), 'not_first': False}

        iter_count = 0
        if self.Ftol is not None:
            f_past = self.cost(x)
        Ktol_converged = False
        Ftol_converged = False
        self.signal = 0
        outer = 0
        inner = 0

        for outer in range(self.niter):
            if all([np.linalg.norm(kkt[0]) <= self.Ktol, np.linalg.norm(kkt[1]) <= self.Ktol,
                    np.linalg.norm(kkt[2]) <= self.Ktol, np.linalg.norm(kkt[3]) <= self.Ktol]):
                self.signal = 1
                Ktol_converged = True
                break

            for inner in range(self.miter):
                muTol = np.max([self.Ktol, self.mu_host])
                if all([np.linalg.norm(kkt[0]) <= muTol, np.linalg.norm(kkt[1]) <= muTol,
                        np.linalg.norm(kkt[2]) <= muTol, np.linalg.norm(kkt[3]) <= muTol]):
                    if not neq and not nineq:
                        self.signal = 1
                        Ktol_converged = True
                    break

                if self.lbfgs:
                    self._lb['not_first'] = (inner > 0 or outer > 0)     # pyipm.py:1705
                x, s, lda, kkt = self.newton_step(x, s, lda)
                iter_count += 1

                if all([self.Ftol is not None, not nineq, self.signal != -2]):
                    f_new = self.cost(x)
                    if np.abs(f_past - f_new) <= np.abs(self.Ftol):
                        self.signal = 2
                        Ftol_converged = True
                        break
                    else:
                        f_past = f_new

                if self.signal == -2:
                    break

            if all([self.Ftol is not None, nineq, self.signal != -2]):
                f_new = self.cost(x)
                if np.abs(f_past - f_new) <= np.abs(self.Ftol):
                    self.signal = 2
                    Ftol_converged = True
                else:
                    f_past = f_new

            if self.Ftol is not None and Ftol_converged:
                break
            if self.signal == -2:
                break
            if outer >= self.niter - 1:
                self.signal = -1
                break

            if nineq:
                # barrier parameter update, pyipm.py:1804-1814
                xi = nineq * np.min(s * lda[neq:]) / (np.dot(s, lda[neq:]) + self.eps)
                self.mu_host = (0.1 * np.min([0.05 * (1.0 - xi) / (xi + self.eps), 2.0]) ** 3 *
                                np.dot(s, lda[neq:]) / nineq)
                if self.float_dtype(self.mu_host) < self.float_dtype(0.0):
                    self.mu_host = 0.0
                self.mu_host = self.float_dtype(self.mu_host)
                self.mu_dev = self.mu_host

        self.x = x
        self.s = s
        self.lda = lda
        self.kkt = kkt
        self.fval = self.cost(x)
        self.iter_count = iter_count
        self.outer_last = outer
        self.inner_last = inner
        self.Ktol_converged = Ktol_converged
        self.Ftol_converged = Ftol_converged

        if self.verbosity >= 0:
            print(self.summary())
        return self.x, self.s, self.lda, self.fval, self.kkt

    def summary(self):
        """Final report line, pyipm.py:1823-1851."""
        kkt, outer, inner, iter_count = self.kkt, self.outer_last, self.inner_last, self.iter_count
        msg = []
        if self.signal == -2:
            msg.append('Terminated due to bad direction in backtracking line search')
        elif all([np.linalg.norm(kkt[0]) <= self.Ktol, np.linalg.norm(kkt[1]) <= self.Ktol,
                  np.linalg.norm(kkt[2]) <= self.Ktol, np.linalg.norm(kkt[3]) <= self.Ktol]):
            msg.append('Converged to Ktol tolerance')
        elif self.Ftol is not None and self.Ftol_converged:
            msg.append('Converged to Ftol tolerance')
        else:
            msg.append('Maximum iterations reached')
            outer = self.niter
            inner = 0
        if self.nineq:
            if outer > 1:
                msg.append('after {} outer'.format(outer - 1))
                msg.append('iterations' if outer > 2 else 'iteration')
                msg.append('and')
            else:
                msg.append('after')
            msg.append('{} inner'.format(inner))
            msg.append('iterations' if inner > 1 else 'iteration')
            msg.append('({} total).'.format(iter_count))
        else:
            msg.append('after {}'.format(iter_count))
            msg.append('iterations.' if iter_count > 1 else 'iteration.')
        return ' '.join(msg)
