"""Drop-in replacement for ``pyipm.IPM`` (reference: /root/reference/pyipm.py:23-1863) whose per-iteration hot
path runs on a B200 through libb200ipm.so.

Same constructor keywords (pyipm.py:311-314), same ``solve() -> (x, s, lda, fval, kkt)`` (pyipm.py:1567, 1863),
same ``KKT()``, ``validate()``, ``compile()``, the same progress messages and ``signal`` codes.  What changed:

* Aesara is gone.  ``f`` is either a *lowerable problem description* (``problems.PolyProblem`` /
  ``problems.QuadProblem``: derivatives are generated and evaluated on the device -- this replaces symbolic
  autodiff, pyipm.py:473-509) or a set of plain Python callables ``f, df, d2f, ce, dce, d2ce, ci, dci, d2ci``
  with the reference's "precompiled function" conventions (pyipm.py:216-231).  ``x_dev`` / ``lambda_dev`` are
  accepted and ignored.
* ``lbfgs=m`` selects the reference's L-BFGS mode (pyipm.py:993-1371): compact representation + Woodbury on the device
  (``b200ipm_lbfgs_*``); second derivatives are then not needed.
* The callable slots of the reference (``cost, grad, hess, con, jaco, phi, dphi, init_lambda, init_slack``)
  exist with the same signatures and run on the device for lowered problems.

There is no CPU fallback: constructing the engine without libb200ipm.so or without a CUDA device raises.
"""
from __future__ import print_function

import numpy as np

from . import _lib
from . import problems as _problems


class IPM(object):
    def __init__(self, x0=None, x_dev=None, f=None, df=None, d2f=None, ce=None, dce=None, d2ce=None, ci=None, dci=None,
                 d2ci=None, lda0=None, lambda_dev=None, s0=None, mu=0.2, nu=10.0, rho=0.1, tau=0.995, eta=1.0E-4,
                 beta=0.4, miter=20, niter=10, Xtol=None, Ktol=1.0E-4, Ftol=None, lbfgs=False, lbfgs_zeta=None,
                 float_dtype=np.float64, verbosity=1, device=0, stream=None, nrefine=2, engine_flags=None,
                 device_callables=False):
        # pyipm.py:316-376
        self.x0 = x0
        self.x_dev = x_dev            # ignored (no symbolic graph)
        self.lda0 = lda0
        self.lambda_dev = lambda_dev  # ignored
        self.s0 = s0
        from . import symbolic as _symbolic
        if _symbolic.is_symbolic(f):
            # the reference's expression input mode (pyipm.py:83-146): SymPy expressions of the symbols in x_dev instead
            # of Aesara expressions of x_dev; polynomials are lowered to the device, anything else is differentiated
            # symbolically into callables
            assert x_dev is not None, 'symbolic f/ce/ci need x_dev = the sequence of SymPy symbols they are written in'
            assert all(d is None for d in (df, d2f, dce, d2ce, dci, d2ci)), \
                'with symbolic f/ce/ci the derivatives are generated; do not pass df/d2f/dce/d2ce/dci/d2ci'
            lowered = _symbolic.lower(x_dev, f, ce, ci)
            if isinstance(lowered, dict):
                f, df, d2f = lowered['f'], lowered['df'], lowered['d2f']
                ce, dce, d2ce = lowered.get('ce'), lowered.get('dce'), lowered.get('d2ce')
                ci, dci, d2ci = lowered.get('ci'), lowered.get('dci'), lowered.get('d2ci')
            else:
                f, ce, ci = lowered, None, None
        self.problem = f if isinstance(f, (_problems.PolyProblem, _problems.QuadProblem)) else None
        self.f, self.df, self.d2f = f, df, d2f
        self.ce, self.dce, self.d2ce = ce, dce, d2ce
        self.ci, self.dci, self.d2ci = ci, dci, d2ci
        self.nvar = self.neq = self.nineq = None
        self.eps = np.finfo(float_dtype).eps
        self.mu, self.nu, self.rho, self.tau, self.eta, self.beta = mu, nu, rho, tau, eta, beta
        self.miter, self.niter = miter, niter
        self.Xtol = Xtol if Xtol else self.eps
        self.Ktol, self.Ftol = Ktol, Ftol
        self.reg_coef = float_dtype(np.sqrt(self.eps))
        # pyipm.py:355-360
        self.lbfgs = lbfgs
        self.lbfgs_zeta = float_dtype(1.0) if (lbfgs and lbfgs_zeta is None) else lbfgs_zeta
        self.lbfgs_fail_max = lbfgs
        self.float_dtype = float_dtype
        # the reference's two shared scalars (pyipm.py:363-364)
        self.nu_dev = float(nu)
        self.mu_dev = float(mu)
        self.verbosity = verbosity
        self.delta0 = self.reg_coef
        self.compiled = False
        self.device, self.stream, self.nrefine = device, stream, nrefine
        # device_callables: f/df/d2f/ce/... take and return torch CUDA float64 tensors on `device` (the reference's
        # 'precompiled function' input mode, pyipm.py:216-231, without leaving the GPU: b200ipm_set_derivs(on_device = 1))
        self.device_callables = bool(device_callables)
        self.engine_flags = engine_flags      # b200ipm_params.flags (None: _lib.DEFAULT_FLAGS; 0: fp64 DMMA contractions)
        self.engine = None
        self.delta = 0.0
        self.mu_host = mu
        self.signal = 0
        self.step_log = []   # one dict per Newton step (timings, inertia, alphas) for benchmarking/tests

    # ------------------------------------------------------------------ validate / compile
    def validate(self):
        """pyipm.py:385-408."""
        assert self.f is not None
        if self.problem is None:
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
        assert self.float_dtype == np.float64, 'pyipm_b200 computes in float64'

    def compile(self, nvar=None, neq=None, nineq=None):
        """Size discovery (pyipm.py:414-467), device workspace allocation and problem binding."""
        if nvar is not None:
            self.nvar = nvar
        if self.problem is not None:
            self.nvar, self.neq, self.nineq = self.problem.nvar, self.problem.neq, self.problem.nineq
        else:
            self.neq = neq
            self.nineq = nineq
            if self.ce is not None and self.neq is None:
                self.neq = self._h(self.ce(self._arg(self.x0))).size
            elif neq is None:
                self.neq = 0
            if self.ci is not None and self.nineq is None:
                self.nineq = self._h(self.ci(self._arg(self.x0))).size
            elif nineq is None:
                self.nineq = 0
            need = ('df',) + (('dce',) if self.neq else ()) + (('dci',) if self.nineq else ())
            if not self.lbfgs:     # second derivatives are only used by the exact-Hessian mode (pyipm.py:478-509)
                need += ('d2f',) + (('d2ce',) if self.neq else ()) + (('d2ci',) if self.nineq else ())
            for name in need:
                assert getattr(self, name) is not None, \
                    'callable mode needs %s (no symbolic autodiff; pass a PolyProblem/QuadProblem to have ' \
                    'derivatives generated on the device)' % name
        params = _lib.default_params(mu=self.mu, nu=self.nu, rho=self.rho, tau=self.tau, eta=self.eta, beta=self.beta,
                                     Xtol=self.Xtol, Ktol=self.Ktol, nrefine=self.nrefine, flags=self.engine_flags)
        if self.engine is not None:
            self.engine.close()
        self.engine = _lib.Engine(self.nvar, self.neq, self.nineq, params, device=self.device, stream=self.stream)
        if self.problem is not None:
            self.engine.bind(self.problem)
        self.compiled = True

    # ------------------------------------------------------------------ state plumbing
    def _push(self, x, s, lda):
        self.engine.set_state(x, s if self.nineq else None, lda if (self.neq + self.nineq) else None,
                              self.mu_dev, self.nu_dev, self.delta)
        self.engine.set_mu_host(self.mu_host)
        if self.problem is None:
            self._upload_derivs(np.asarray(x, dtype=np.float64), np.asarray(lda, dtype=np.float64))

    def _upload_derivs(self, x, lda):
        D, M, N = self.nvar, self.neq, self.nineq
        if self.device_callables:
            return self._upload_derivs_device(x, lda)
        fval = float(self.f(x))
        df = np.asarray(self.df(x), dtype=np.float64).reshape(D)
        W = None if self.lbfgs else np.array(self.d2f(x), dtype=np.float64).reshape(D, D)
        ce = ci = None
        blocks = []
        if M:
            ce = np.asarray(self.ce(x), dtype=np.float64).reshape(M)
            blocks.append(np.asarray(self.dce(x), dtype=np.float64).reshape(D, M))
            if W is not None:
                W = W - np.asarray(self.d2ce(x, lda), dtype=np.float64).reshape(D, D)
        if N:
            ci = np.asarray(self.ci(x), dtype=np.float64).reshape(N)
            blocks.append(np.asarray(self.dci(x), dtype=np.float64).reshape(D, N))
            if W is not None:
                W = W - np.asarray(self.d2ci(x, lda), dtype=np.float64).reshape(D, D)
        J = np.concatenate(blocks, axis=1) if blocks else None
        self.engine.set_derivs(fval, df, ce, ci, J, W)

    # ---- torch-CUDA callables: nothing but scalars crosses PCIe
    def _dev(self, a):
        import torch
        return torch.as_tensor(np.asarray(a, dtype=np.float64), device=torch.device('cuda', self.device))

    def _upload_derivs_device(self, x, lda):
        import torch
        D, M, N = self.nvar, self.neq, self.nineq
        xt, lt = self._dev(x), self._dev(lda)
        fval = float(self.f(xt))
        df = self.df(xt).reshape(D).contiguous()
        W = None if self.lbfgs else self.d2f(xt).reshape(D, D)
        ce = ci = None
        blocks = []
        if M:
            ce = self.ce(xt).reshape(M).contiguous()
            blocks.append(self.dce(xt).reshape(D, M))
            if W is not None:
                W = W - self.d2ce(xt, lt).reshape(D, D)
        if N:
            ci = self.ci(xt).reshape(N).contiguous()
            blocks.append(self.dci(xt).reshape(D, N))
            if W is not None:
                W = W - self.d2ci(xt, lt).reshape(D, D)
        J = torch.cat(blocks, dim=1).contiguous() if blocks else None
        if W is not None:
            W = W.contiguous()
        torch.cuda.current_stream(self.device).synchronize()      # the engine copies on its own stream
        self.engine.set_derivs_device(fval, df, ce, ci, J, W)

    def _h(self, v):
        """host NumPy view of a callable's result (torch CUDA tensor in device_callables mode)"""
        return v.detach().cpu().numpy() if self.device_callables else np.asarray(v)

    def _arg(self, x):
        return self._dev(x) if self.device_callables else x

    # ------------------------------------------------------------------ operator slots (pyipm.py:855-954)
    def _at(self, x, s=None, lda=None):
        s = np.zeros(self.nineq) if s is None else s
        lda = np.zeros(self.neq + self.nineq) if lda is None else lda
        self._push(np.asarray(x, dtype=np.float64), np.asarray(s, dtype=np.float64), np.asarray(lda, dtype=np.float64))

    def cost(self, x):
        self._at(x, np.ones(self.nineq))
        return self.engine.cost()

    def grad(self, x, s, lda):
        self._at(x, s, lda)
        return self.engine.residual()[0]

    def hess(self, x, s, lda):
        self._at(x, s, lda)
        return self.engine.hess_full()

    def con(self, x, s):
        self._at(x, s)
        return self.engine.con_jac()[0]

    def jaco(self, x):
        self._at(x, np.ones(self.nineq))
        D, M, N = self.nvar, self.neq, self.nineq
        top = self.engine.con_jac()[1]
        if not N:
            return top
        bottom = np.concatenate([np.zeros((N, M)), -np.eye(N)], axis=1)
        return np.concatenate([top, bottom], axis=0)

    def KKT(self, x, s, lda):
        """pyipm.py:958-991 (absent conditions are scalar 0.0; returns a tuple)."""
        self._at(x, s, lda)
        k1, k2, k3, k4 = self.engine.kkt()
        z = self.float_dtype(0.0)
        return k1, (k2 if self.nineq else z), (k3 if self.neq else z), (k4 if self.nineq else z)

    # ------------------------------------------------------------------ solve
    def solve(self, x0=None, s0=None, lda0=None, force_recompile=False):
        """pyipm.py:1567-1863 with the inner-iteration body (pyipm.py:1714-1754) executed by
        ``b200ipm_newton_step`` on the device."""
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
        D, M, N = self.nvar, self.neq, self.nineq
        eng = self.engine
        lowered = self.problem is not None
        self.step_log = []

        # initialise weights, slacks and multipliers (pyipm.py:1596-1625)
        x = np.array(self.x0, dtype=np.float64)
        self.delta = 0.0
        if N:
            self.mu_host = self.mu               # quirk xi: mu_dev is NOT reset (pyipm.py:1603)
        else:
            self.mu_host = self.Ktol             # pyipm.py:1606-1607
            self.mu_dev = float(self.mu_host)
        if M or N:
            self.nu_host = self.nu
            self.nu_dev = float(self.nu_host)
        s = np.zeros(N)
        lda = np.zeros(M + N)
        if N and self.s0 is not None:
            s = np.asarray(self.s0, dtype=np.float64)
        if (M or N) and self.lda0 is not None:
            lda = np.asarray(self.lda0, dtype=np.float64)
        self._push(x, s if (not N or self.s0 is not None) else np.ones(N), lda)
        if N and self.s0 is None:
            if lowered:
                eng.init_slack()
            else:
                s = np.maximum(self._h(self.ci(self._arg(x))).astype(np.float64).reshape(N), self.Ktol)
                self._push(x, s, lda)
        if (M or N) and self.lda0 is None:
            eng.init_lambda()
        x, s, lda, _, _, _ = eng.get_state()
        if not lowered:
            self._push(x, s, lda)

        _, nrm = eng.residual(want_g=False)
        iter_count = 0
        if self.Ftol is not None:
            f_past = eng.cost()
        Ktol_converged = False
        Ftol_converged = False
        self.signal = 0
        if self.lbfgs:
            # pyipm.py:1633-1637
            eng.lbfgs_init(self.lbfgs, float(self.lbfgs_zeta))
            self._x_old = np.array(x)
        if self.verbosity > 0:
            if self.lbfgs:
                print('Searching for a feasible local minimizer using L-BFGS to approximate the Hessian.')
            else:
                print('Searching for a feasible local minimizer using the exact Hessian.')
        outer = inner = 0

        for outer in range(self.niter):
            if all(nrm <= self.Ktol):
                self.signal = 1
                Ktol_converged = True
                break
            if self.verbosity > 0 and N:
                print('OUTER ITERATION {}'.format(outer + 1))

            for inner in range(self.miter):
                muTol = np.max([self.Ktol, self.mu_host])
                if all(nrm <= muTol):
                    if not M and not N:
                        self.signal = 1
                        Ktol_converged = True
                    break

                if self.verbosity > 0:
                    msg = []
                    if N:
                        msg.append('* INNER ITERATION {}'.format(inner + 1))
                    else:
                        msg.append('ITERATION {}'.format(iter_count + 1))
                    if self.verbosity > 1:
                        msg.append('f(x) = {}'.format(eng.cost()))
                    if self.verbosity > 2:
                        msg.append('|dL/dx| = {}'.format(nrm[0]))
                        msg.append('|dL/ds| = {}'.format(nrm[1]))
                        msg.append('|ce| = {}'.format(nrm[2]))
                        msg.append('|ci-s| = {}'.format(nrm[3]))
                    print(', '.join(msg))

                # ---- one Newton step on the device (pyipm.py:1714-1754)
                if lowered:
                    # pyipm.py:1702-1713: in L-BFGS mode the storage is updated before every direction but the first
                    info = eng.lbfgs_step(inner > 0 or outer > 0) if self.lbfgs else eng.newton_step()
                    nrm = np.array(list(info.kkt_norm))
                    f_new_dev = info.fval
                else:
                    info, nrm, f_new_dev = self._callable_step(inner > 0 or outer > 0)
                self.nu_dev = self.nu_host = info.nu
                self.delta = info.delta
                if info.signal == -2:
                    self.signal = -2
                    if self.verbosity > 2:
                        print('Search direction is unreliable to machine precision.')
                if info.soc_accepted and self.verbosity > 2:
                    print('Second-order feasibility correction accepted')
                self.step_log.append(info.asdict())
                if self.lbfgs:
                    m_, zeta_, fail_ = eng.lbfgs_state()
                    self.step_log[-1].update(lbfgs_m=m_, lbfgs_zeta=zeta_, lbfgs_fail=fail_)
                iter_count += 1

                if all([self.Ftol is not None, not N, self.signal != -2]):
                    f_new = f_new_dev
                    if np.abs(f_past - f_new) <= np.abs(self.Ftol):
                        self.signal = 2
                        Ftol_converged = True
                        break
                    else:
                        f_past = f_new
                if self.signal == -2:
                    break
                if inner >= self.miter - 1:
                    if self.verbosity > 0 and N:
                        print('MAXIMUM INNER ITERATIONS EXCEEDED')

            if all([self.Ftol is not None, N, self.signal != -2]):
                f_new = eng.cost()
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
                if self.verbosity > 0:
                    print('MAXIMUM OUTER ITERATIONS EXCEEDED' if N else 'MAXIMUM ITERATIONS EXCEEDED')
                break

            if N:
                # barrier parameter update (pyipm.py:1804-1814)
                self.mu_host = float(eng.update_mu())
                self.mu_dev = self.mu_host
                xs, ss, ls, _, nu_, dl_ = eng.get_state()
                eng.set_state(None, None, None, self.mu_dev, nu_, dl_)
                eng.set_mu_host(self.mu_host)
                if not lowered:
                    self._push(xs, ss, ls)
                _, nrm = eng.residual(want_g=False)

        x, s, lda, _, _, _ = eng.get_state()
        self.x, self.s, self.lda = x, s, lda
        if not lowered:
            self._push(x, s, lda)
        self.kkt = self.KKT(x, s, lda)
        self.fval = eng.cost()
        kn = [np.linalg.norm(k) for k in self.kkt]

        if self.verbosity >= 0:
            msg = []
            if self.signal == -2:
                msg.append('Terminated due to bad direction in backtracking line search')
            elif all(k <= self.Ktol for k in kn):
                msg.append('Converged to Ktol tolerance')
            elif self.Ftol is not None and Ftol_converged:
                msg.append('Converged to Ftol tolerance')
            else:
                msg.append('Maximum iterations reached')
                outer = self.niter
                inner = 0
            if N:
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
            print(' '.join(msg))
            if self.verbosity > 1:
                msg = ['FINAL: f(x) = {}'.format(self.fval)]
                if self.verbosity > 2:
                    msg.append('|dL/dx| = {}'.format(kn[0]))
                    msg.append('|dL/ds| = {}'.format(kn[1]))
                    msg.append('|ce| = {}'.format(kn[2]))
                    msg.append('|ci-s| = {}'.format(kn[3]))
                print(', '.join(msg))
        self.iter_count = iter_count
        return self.x, self.s, self.lda, self.fval, self.kkt

    # ------------------------------------------------------------------ callable mode
    def _callable_step(self, not_first=True):
        """One inner iteration when f/ce/ci are opaque host callables: derivatives are evaluated by the user's
        functions and uploaded; residual, KKT formation, inertia-corrected factorisation, solve, nu rule and the
        fraction-to-the-boundary rule run on the device; the Armijo loop has to call the user's host functions
        for every trial point, so its scalar bookkeeping (pyipm.py:1457-1505) stays on the host; the second-order
        correction's least squares (pyipm.py:1468-1477, 1520-1529) is b200ipm_soc_direction on the device.  With
        device_callables=True the user's functions take / return torch CUDA tensors and only scalars reach the host."""
        eng = self.engine
        D, M, N = self.nvar, self.neq, self.nineq
        x, s, lda, _, _, _ = eng.get_state()
        if self.lbfgs:
            if not_first:
                # pyipm.py:1705-1710: dL/dx at x_old with the CURRENT multipliers, from the user's callables
                xo = self._x_old
                xa = self._arg(xo)
                go = self._h(self.df(xa)).astype(np.float64).reshape(D)
                if M:
                    go = go - np.dot(self._h(self.dce(xa)).reshape(D, M), lda[:M])
                if N:
                    go = go - np.dot(self._h(self.dci(xa)).reshape(D, N), lda[M:])
                eng.lbfgs_update(go)
                self._x_old = np.array(x)
            dz, info = eng.lbfgs_direction()
            info.delta = self.delta
        else:
            dz, info = eng.direction()
        # the shift reghess settled on persists across steps (pyipm.py:1390-1395): every later set_state of this step
        # must carry it, or the next factorisation would restart from the pre-step value
        self.delta = info.delta
        _, nrm0 = eng.residual(want_g=False)

        def con_vec(xx, ss):
            """con(x, s) = [ce ; ci - s] (pyipm.py:564-579) through the user's functions"""
            parts = []
            xa = self._arg(xx)
            if M:
                parts.append(self._h(self.ce(xa)).reshape(M))
            if N:
                parts.append(self._h(self.ci(xa)).reshape(N) - ss)
            return np.concatenate(parts) if parts else np.zeros(0)

        con_l1 = float(np.sum(np.abs(con_vec(x, s))))
        df = self._h(self.df(self._arg(x))).astype(np.float64).reshape(D)
        if M or N:
            bcg = np.concatenate([df, -self.mu_dev / (s + self.eps)]) if N else df
            nu_thres = np.dot(bcg, dz[:D + N]) / (1 - self.rho) / con_l1
            if self.nu_dev < nu_thres:
                self.nu_dev = float(nu_thres)
        eng.set_state(None, None, None, self.mu_dev, self.nu_dev, info.delta)
        a_s, a_l = eng.step_max() if N else (1.0, 1.0)
        if not (M or N):
            a_l = 0.0
        dx, ds, dl = dz[:D], dz[D:D + N], dz[D + N:]

        def phi(xx, ss):
            v = float(self.f(self._arg(xx)))
            if M or N:
                v += self.nu_dev * np.sum(np.abs(con_vec(xx, ss)))
            if N:
                v -= self.mu_dev * np.sum(np.log(ss))
            return v

        def step_host(v, dv):
            """closed form of step() (pyipm.py:1408-1436), as the device kernels compute it"""
            thr = (1.0 - self.tau) * v
            bad = ~(v + dv >= thr)
            if not bad.any():
                return 1.0
            a = np.where(dv < 0.0, (v - thr) / np.where(dv < 0.0, -dv, 1.0), 0.0)
            return float(min(np.min(a[bad]), 1.0))

        phi0 = phi(x, s)
        dphi0 = np.dot(df, dx)
        if M or N:
            dphi0 -= self.nu_dev * con_l1
        if N:
            dphi0 -= np.dot(self.mu_dev / (s + self.eps), ds)
        info.alpha_smax, info.alpha_lmax = a_s, a_l
        info.phi0, info.dphi0 = phi0, dphi0
        nb = 0
        correction = False
        alpha_corr = 0.0
        dz_p = None
        with np.errstate(all='ignore'):
            if phi(x + a_s * dx, s + a_s * ds) > phi0 + a_s * self.eta * dphi0:
                if M or N:
                    # second-order correction (pyipm.py:1464-1489 / 1516-1536): least squares on the device
                    c_new = con_vec(x + a_s * dx, s + a_s * ds)
                    if np.sum(np.abs(c_new)) > con_l1:
                        info.soc_tried = 1
                        dz_p = eng.soc_direction(c_new)
                        px, ps = dz_p[:D], dz_p[D:]
                        if phi(x + a_s * dx + px, s + a_s * ds + ps) <= phi0 + a_s * self.eta * dphi0:
                            if N:
                                alpha_corr = step_host(s, a_s * ds + ps)
                                if (phi(x + alpha_corr * (a_s * dx + px), s + alpha_corr * (a_s * ds + ps)) <=
                                        phi0 + a_s * self.eta * dphi0):
                                    correction = True
                            else:
                                alpha_corr = 1.0
                                correction = True
                if not correction:
                    a_s *= self.tau
                    a_l *= self.tau
                    nb = 1
                    while phi(x + a_s * dx, s + a_s * ds) > phi0 + a_s * self.eta * dphi0:
                        nrm_step = (np.sqrt(np.linalg.norm(a_s * dx) ** 2 + np.linalg.norm(a_l * ds) ** 2) if N
                                    else np.linalg.norm(a_s * dx))
                        if nrm_step < self.eps:
                            info.signal = -2
                            break
                        a_s *= self.tau
                        a_l *= self.tau
                        nb += 1
        info.n_backtracks = nb
        info.soc_accepted = 1 if correction else 0
        info.alpha_corr = alpha_corr
        if info.signal != -2:
            if correction:
                x = x + alpha_corr * (a_s * dx + dz_p[:D])
                s = s + alpha_corr * (a_s * ds + dz_p[D:])
            else:
                x = x + a_s * dx
                s = s + a_s * ds
            lda = lda + a_l * dl if (M or N) else lda
        info.alpha_s, info.alpha_l, info.nu = a_s, a_l, self.nu_dev
        self._push(x, s, lda)
        _, nrm = eng.residual(want_g=False)
        return info, nrm, float(self.f(self._arg(x)))
