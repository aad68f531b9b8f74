"""Batched multi-start solves of small problems on one GPU.

``solve_batch(problem, X0, ...)`` runs the reference's whole ``IPM.solve()`` loop (pyipm.py:1567-1863, exact-Hessian
mode, same keyword defaults) for every row of ``X0`` -- one warp per starting point, one kernel launch for the whole
batch (``b200ipm_batch_solve_poly``).  ``problem`` is a :class:`pyipm_b200.problems.PolyProblem` with
``K = D + 2N + M <= 32`` (all ten example problems of the reference qualify)."""
from __future__ import print_function

import ctypes as C

import numpy as np

from . import _lib
from . import problems as _problems


class BatchResult(object):
    """x (B, D), s (B, N), lda (B, M+N), fval (B,), kkt_norm (B, 4), signal (B,), iters (B,), ms (kernel time)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def solve_batch(problem, X0, mu=0.2, nu=10.0, rho=0.1, tau=0.995, eta=1.0E-4, beta=0.4, miter=20, niter=10, Ktol=1.0E-4,
                Ftol=None, device=0):
    assert isinstance(problem, _problems.PolyProblem), 'solve_batch needs a PolyProblem (the lowered polynomial form)'
    lib = _lib.load()
    X0 = np.ascontiguousarray(np.atleast_2d(X0), dtype=np.float64)
    B, D = X0.shape
    M, N = problem.neq, problem.nineq
    assert D == problem.nvar
    d = problem.descriptor()
    p = _lib.default_params(mu=mu, nu=nu, rho=rho, tau=tau, eta=eta, beta=beta, Ktol=Ktol)
    x = np.empty((B, D))
    s = np.empty((B, N))
    lda = np.empty((B, M + N))
    fval = np.empty(B)
    kkt = np.empty((B, 4))
    sig = np.empty(B, dtype=np.int32)
    its = np.empty(B, dtype=np.int32)
    ms = C.c_float()
    _lib.check(lib.b200ipm_batch_solve_poly(
        D, M, N, int(d['nterms']), d['term_row'].ctypes.data_as(_lib.c_int_p), d['term_coeff'].ctypes.data_as(_lib.c_double_p),
        d['term_ptr'].ctypes.data_as(_lib.c_int_p), d['fac_var'].ctypes.data_as(_lib.c_int_p),
        d['fac_pow'].ctypes.data_as(_lib.c_int_p), float(d['xlogx_coeff']), float(d['xlogx_shift']), C.byref(p), int(niter),
        int(miter), 0 if Ftol is None else 1, 0.0 if Ftol is None else float(Ftol), B, _lib.ptr(X0), int(device), _lib.ptr(x),
        _lib.ptr(s), _lib.ptr(lda), _lib.ptr(fval), _lib.ptr(kkt), _lib.ptr(sig), _lib.ptr(its), C.byref(ms)))
    return BatchResult(x=x, s=s, lda=lda, fval=fval, kkt_norm=kkt, signal=sig, iters=its, ms=ms.value)
