"""ctypes binding of libb200ipm.so (include/b200ipm.h).

This is the binding a maintainer of the reference would add to call the B200 engine from ``IPM.solve()``
(see INTEGRATION.md).  There is NO CPU fallback: if the shared library is missing or no CUDA device is
visible, loading / handle creation raises.
"""
from __future__ import print_function

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('B200IPM_LIB', os.path.join(_HERE, 'libb200ipm.so'))   # override: profiling builds

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class Params(C.Structure):
    """struct b200ipm_params"""
    _fields_ = [('mu', C.c_double), ('nu', C.c_double), ('rho', C.c_double), ('tau', C.c_double),
                ('eta', C.c_double), ('beta', C.c_double), ('Xtol', C.c_double), ('Ktol', C.c_double),
                ('eps', C.c_double), ('reg_coef', C.c_double), ('nrefine', C.c_int), ('ls_batch', C.c_int),
                ('max_reg_retries', C.c_int), ('flags', C.c_int)]


class StepInfo(C.Structure):
    """struct b200ipm_step_info"""
    _fields_ = [('kkt_norm', C.c_double * 4), ('fval', C.c_double), ('delta', C.c_double), ('mu', C.c_double),
                ('nu', C.c_double), ('alpha_smax', C.c_double), ('alpha_lmax', C.c_double), ('alpha_s', C.c_double),
                ('alpha_l', C.c_double), ('alpha_corr', C.c_double), ('phi0', C.c_double), ('dphi0', C.c_double),
                ('rcond', C.c_double), ('resid', C.c_double), ('con_l1', C.c_double),
                ('n_neg', C.c_int), ('n_zero', C.c_int), ('n_factor', C.c_int), ('n_backtracks', C.c_int),
                ('soc_tried', C.c_int), ('soc_accepted', C.c_int), ('signal', C.c_int), ('eq_reg', C.c_int),
                ('n_neg_first', C.c_int), ('n_zero_first', C.c_int),
                ('ms_eval', C.c_float), ('ms_assemble', C.c_float), ('ms_factor', C.c_float), ('ms_solve', C.c_float),
                ('ms_search', C.c_float), ('ms_total', C.c_float), ('ms_hess_kernel', C.c_float),
                ('ms_condense_kernel', C.c_float), ('n_spec', C.c_int), ('spec_used', C.c_int),
                ('tc_syrk', C.c_int), ('abandoned_first', C.c_int), ('cert_used', C.c_int), ('n_factor_phys', C.c_int)]

    def asdict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if name == 'kkt_norm' else v
        return d


# b200ipm_params.flags used when the caller does not choose: the two dense contractions on tcgen05 (error-free int8
# split, 128x128 tiles), speculative + abandoning reghess.  0 = fp64 DMMA contractions.
FLAG_NO_SPECULATION, FLAG_TCGEN05_SYRK, FLAG_NO_ABANDON, FLAG_DELAY_BG, FLAG_NO_CERT = 1, 2, 16, 32, 64
DEFAULT_FLAGS = FLAG_TCGEN05_SYRK | (1 << 2)

# every symbol include/b200ipm.h declares (tests/test_abi.py checks the library exports exactly these)
SYMBOLS = [
    'b200ipm_version', 'b200ipm_last_error', 'b200ipm_launch_count', 'b200ipm_struct_size', 'b200ipm_create', 'b200ipm_destroy',
    'b200ipm_set_params', 'b200ipm_sync', 'b200ipm_bind_quad', 'b200ipm_bind_poly', 'b200ipm_set_derivs',
    'b200ipm_set_state', 'b200ipm_get_state', 'b200ipm_set_mu_host', 'b200ipm_state_save', 'b200ipm_state_restore',
    'b200ipm_profile_kernel', 'b200ipm_cost', 'b200ipm_residual',
    'b200ipm_kkt', 'b200ipm_con_jac', 'b200ipm_hess_full', 'b200ipm_d2L', 'b200ipm_merit', 'b200ipm_init_slack',
    'b200ipm_init_lambda', 'b200ipm_update_mu', 'b200ipm_direction', 'b200ipm_step_max', 'b200ipm_newton_step',
    'b200ipm_ldlt_create', 'b200ipm_ldlt_destroy', 'b200ipm_ldlt_factor', 'b200ipm_ldlt_solve',
    'b200ipm_ldlt_tile_factor', 'b200ipm_ldlt_panel', 'b200ipm_ldlt_import', 'b200ipm_gemm_nt_update', 'b200ipm_gemm_nt_update_bc', 'b200ipm_test_syrk', 'b200ipm_test_syrk_i8', 'b200ipm_trace_start', 'b200ipm_trace_dump',
    'b200ipm_test_gemv', 'b200ipm_lbfgs_init', 'b200ipm_lbfgs_update', 'b200ipm_lbfgs_direction', 'b200ipm_lbfgs_step',
    'b200ipm_lbfgs_state', 'b200ipm_batch_solve_poly', 'b200ipm_soc_direction', 'b200ipm_ldlt_gemv', 'b200ipm_ldlt_block_factor', 'b200ipm_ldlt_block_panel', 'b200ipm_ldlt_colblock_factor', 'b200ipm_oz_panel_slice', 'b200ipm_oz_block_update', 'b200ipm_oz_status',
]

_lib = None


class B200Error(RuntimeError):
    pass


def load():
    """Load libb200ipm.so (once).  Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError('libb200ipm.so not found at %s: build it with `python -c "import __graft_entry__ as g; '
                          'g.build()"` (pyipm_b200 has no CPU fallback)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, d, i = C.c_void_p, C.c_double, C.c_int
    dp, ip = c_double_p, c_int_p
    sig = {
        'b200ipm_version': (i, []),
        'b200ipm_last_error': (C.c_char_p, []),
        'b200ipm_launch_count': (C.c_longlong, []),
        'b200ipm_struct_size': (i, [i]),
        'b200ipm_create': (i, [i, i, i, C.POINTER(Params), i, vp, C.POINTER(vp)]),
        'b200ipm_destroy': (i, [vp]),
        'b200ipm_set_params': (i, [vp, C.POINTER(Params)]),
        'b200ipm_sync': (i, [vp]),
        'b200ipm_bind_quad': (i, [vp, vp, vp, d, vp, vp, vp, vp, vp, vp, i]),
        'b200ipm_bind_poly': (i, [vp, i, ip, dp, ip, ip, ip, d, d]),
        'b200ipm_set_derivs': (i, [vp, d, vp, vp, vp, vp, vp, i]),
        'b200ipm_set_state': (i, [vp, vp, vp, vp, d, d, d]),
        'b200ipm_get_state': (i, [vp, vp, vp, vp, dp, dp, dp]),
        'b200ipm_set_mu_host': (i, [vp, d]),
        'b200ipm_state_save': (i, [vp]),
        'b200ipm_state_restore': (i, [vp]),
        'b200ipm_profile_kernel': (i, [vp, i, i, C.POINTER(C.c_float), dp]),
        'b200ipm_cost': (i, [vp, dp]),
        'b200ipm_residual': (i, [vp, vp, dp]),
        'b200ipm_kkt': (i, [vp, vp, vp, vp, vp]),
        'b200ipm_con_jac': (i, [vp, vp, vp]),
        'b200ipm_hess_full': (i, [vp, vp]),
        'b200ipm_d2L': (i, [vp, vp]),
        'b200ipm_merit': (i, [vp, dp, dp]),
        'b200ipm_init_slack': (i, [vp]),
        'b200ipm_init_lambda': (i, [vp]),
        'b200ipm_update_mu': (i, [vp, dp]),
        'b200ipm_direction': (i, [vp, vp, C.POINTER(StepInfo)]),
        'b200ipm_step_max': (i, [vp, dp, dp]),
        'b200ipm_newton_step': (i, [vp, C.POINTER(StepInfo)]),
        'b200ipm_ldlt_create': (i, [i, i, vp, C.POINTER(vp)]),
        'b200ipm_ldlt_destroy': (i, [vp]),
        'b200ipm_ldlt_factor': (i, [vp, vp, i, i, ip, dp]),
        'b200ipm_ldlt_solve': (i, [vp, vp, i, i, i]),
        'b200ipm_ldlt_tile_factor': (i, [vp, vp, i, i, vp, vp, vp, ip]),
        'b200ipm_ldlt_panel': (i, [vp, vp, i, i, vp, vp, vp, vp, i]),
        'b200ipm_ldlt_import': (i, [vp, vp, i, vp, vp, vp]),
        'b200ipm_gemm_nt_update': (i, [vp, vp, i, i, i, vp, i, vp, i, i, i]),
        'b200ipm_gemm_nt_update_bc': (i, [vp, vp, i, i, i, vp, i, vp, i, i, i, i, i, i, i, i, i]),
        'b200ipm_test_syrk': (i, [i, vp, d, vp, d, i, C.POINTER(vp), C.POINTER(vp), ip, dp, vp, i,
                                  C.POINTER(C.c_float)]),
        'b200ipm_test_syrk_i8': (i, [i, vp, d, vp, d, i, C.POINTER(vp), C.POINTER(vp), ip, dp, vp, C.c_uint, i, i, i,
                                     C.POINTER(C.c_float), ip]),
        'b200ipm_test_gemv': (i, [i, i, vp, vp, vp, i]),
        'b200ipm_lbfgs_init': (i, [vp, i, d]),
        'b200ipm_lbfgs_update': (i, [vp, vp]),
        'b200ipm_lbfgs_direction': (i, [vp, vp, C.POINTER(StepInfo)]),
        'b200ipm_lbfgs_step': (i, [vp, i, C.POINTER(StepInfo)]),
        'b200ipm_lbfgs_state': (i, [vp, ip, dp, ip]),
        'b200ipm_batch_solve_poly': (i, [i, i, i, i, ip, dp, ip, ip, ip, d, d, C.POINTER(Params), i, i, i, d, i, vp, i, vp, vp, vp,
                                     vp, vp, vp, vp, C.POINTER(C.c_float)]),
        'b200ipm_soc_direction': (i, [vp, vp, vp]),
        'b200ipm_ldlt_gemv': (i, [vp, vp, i, i, i, vp, vp]),
        'b200ipm_ldlt_block_factor': (i, [vp, vp, i, i, vp, vp]),
        'b200ipm_ldlt_block_panel': (i, [vp, vp, i, i, i, vp, vp]),
        'b200ipm_ldlt_colblock_factor': (i, [vp, vp, i, i, i, vp, vp, vp]),
        'b200ipm_oz_panel_slice': (i, [vp, i, vp, i, vp, i, i]),
        'b200ipm_oz_block_update': (i, [vp, vp, i, i, i, i]),
        'b200ipm_oz_status': (i, [vp, vp]),
        'b200ipm_trace_start': (i, []),
        'b200ipm_trace_dump': (i, [vp, vp, vp, vp, vp, i, ip]),
    }
    for name in SYMBOLS:
        fn = getattr(lib, name)   # AttributeError if the library does not export it
        fn.restype, fn.argtypes = sig[name]
    if lib.b200ipm_struct_size(0) != C.sizeof(Params) or lib.b200ipm_struct_size(1) != C.sizeof(StepInfo):
        raise ImportError('libb200ipm.so was built from a different include/b200ipm.h (struct sizes differ): rebuild')
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise B200Error('libb200ipm: %s (rc=%d)' % (load().b200ipm_last_error().decode('utf-8', 'replace'), rc))


def ptr(a):
    """void* of a C-contiguous float64/int32 array (or None)."""
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def torch_stream_handle(device=None):
    """cudaStream_t of torch's CURRENT stream for the C ABI.  The legacy default stream has handle 0, which the
    ABI reads as 'create a private stream', so it is passed as cudaStreamLegacy (0x1) instead: kernels of this
    library are then ordered with torch ops / NCCL collectives issued on the same stream."""
    import torch
    h = torch.cuda.current_stream(device).cuda_stream
    return h if h else 1


def default_params(mu=0.2, nu=10.0, rho=0.1, tau=0.995, eta=1.0E-4, beta=0.4, Xtol=None, Ktol=1.0E-4, nrefine=2,
                   ls_batch=32, max_reg_retries=60, flags=None):
    eps = float(np.finfo(np.float64).eps)
    return Params(mu=mu, nu=nu, rho=rho, tau=tau, eta=eta, beta=beta, Xtol=Xtol if Xtol else eps, Ktol=Ktol, eps=eps,
                  reg_coef=float(np.sqrt(eps)), nrefine=nrefine, ls_batch=ls_batch, max_reg_retries=max_reg_retries,
                  flags=DEFAULT_FLAGS if flags is None else int(flags))


class Engine(object):
    """Thin object wrapper over a b200ipm_handle (one problem size, one CUDA stream)."""

    def __init__(self, D, M, N, params=None, device=0, stream=None):
        self.lib = load()
        self.D, self.M, self.N = int(D), int(M), int(N)
        self.K = self.D + 2 * self.N + self.M
        self.params = params if params is not None else default_params()
        self.h = C.c_void_p()
        check(self.lib.b200ipm_create(self.D, self.M, self.N, C.byref(self.params), int(device),
                                      C.c_void_p(stream) if stream else None, C.byref(self.h)))
        self._keep = []

    def close(self):
        if getattr(self, 'h', None) is not None and self.h.value:
            self.lib.b200ipm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- binding
    def bind(self, prob):
        from . import problems
        if isinstance(prob, problems.QuadProblem):
            arrs = [f64(prob.Q), f64(prob.c), f64(prob.At), f64(prob.Ut), f64(prob.b), f64(prob.Gt), f64(prob.Vt),
                    f64(prob.r)]
            Q, c, At, Ut, b, Gt, Vt, r = arrs
            check(self.lib.b200ipm_bind_quad(self.h, ptr(Q), ptr(c), float(prob.q4), ptr(At), ptr(Ut), ptr(b), ptr(Gt),
                                             ptr(Vt), ptr(r), 0))
        elif isinstance(prob, problems.PolyProblem):
            d = prob.descriptor()
            check(self.lib.b200ipm_bind_poly(self.h, int(d['nterms']), d['term_row'].ctypes.data_as(c_int_p),
                                             d['term_coeff'].ctypes.data_as(c_double_p),
                                             d['term_ptr'].ctypes.data_as(c_int_p), d['fac_var'].ctypes.data_as(c_int_p),
                                             d['fac_pow'].ctypes.data_as(c_int_p), float(d['xlogx_coeff']),
                                             float(d['xlogx_shift'])))
        else:
            raise TypeError('bind() needs a QuadProblem or PolyProblem')

    def set_derivs(self, fval, df, ce, ci, J, d2L):
        df, ce, ci, J, d2L = f64(df), f64(ce), f64(ci), f64(J), f64(d2L)
        check(self.lib.b200ipm_set_derivs(self.h, float(fval), ptr(df), ptr(ce), ptr(ci), ptr(J), ptr(d2L), 0))

    def set_derivs_device(self, fval, df, ce, ci, J, d2L):
        """Same, from torch CUDA tensors (float64, contiguous) on the engine's device: device-to-device copies, nothing
        crosses PCIe (b200ipm_set_derivs with on_device = 1)."""
        def dp(t):
            return None if t is None else C.c_void_p(t.data_ptr())
        check(self.lib.b200ipm_set_derivs(self.h, float(fval), dp(df), dp(ce), dp(ci), dp(J), dp(d2L), 1))

    def soc_direction(self, cnew):
        cnew = f64(cnew)
        pz = np.empty(self.D + self.N)
        check(self.lib.b200ipm_soc_direction(self.h, ptr(cnew), ptr(pz)))
        return pz

    # ---- state
    def set_state(self, x=None, s=None, lda=None, mu=0.2, nu=10.0, delta=0.0):
        x, s, lda = f64(x), f64(s), f64(lda)
        check(self.lib.b200ipm_set_state(self.h, ptr(x), ptr(s) if self.N else None, ptr(lda) if (self.M + self.N) else None,
                                         float(mu), float(nu), float(delta)))

    def get_state(self):
        x = np.empty(self.D)
        s = np.empty(self.N)
        lda = np.empty(self.M + self.N)
        mu, nu, delta = C.c_double(), C.c_double(), C.c_double()
        check(self.lib.b200ipm_get_state(self.h, ptr(x), ptr(s), ptr(lda), C.byref(mu), C.byref(nu), C.byref(delta)))
        return x, s, lda, mu.value, nu.value, delta.value

    def set_mu_host(self, mu_host):
        check(self.lib.b200ipm_set_mu_host(self.h, float(mu_host)))

    # ---- operator slots
    def cost(self):
        v = C.c_double()
        check(self.lib.b200ipm_cost(self.h, C.byref(v)))
        return v.value

    def residual(self, want_g=True):
        g = np.empty(self.K) if want_g else None
        nrm = (C.c_double * 4)()
        check(self.lib.b200ipm_residual(self.h, ptr(g), nrm))
        return g, np.array(list(nrm))

    def kkt(self):
        k1, k2, k3, k4 = np.empty(self.D), np.empty(self.N), np.empty(self.M), np.empty(self.N)
        check(self.lib.b200ipm_kkt(self.h, ptr(k1), ptr(k2), ptr(k3), ptr(k4)))
        return k1, k2, k3, k4

    def con_jac(self):
        con = np.empty(self.M + self.N)
        J = np.empty((self.D, self.M + self.N))
        check(self.lib.b200ipm_con_jac(self.h, ptr(con), ptr(J)))
        return con, J

    def hess_full(self):
        H = np.empty((self.K, self.K))
        check(self.lib.b200ipm_hess_full(self.h, ptr(H)))
        return H

    def d2L(self):
        W = np.empty((self.D, self.D))
        check(self.lib.b200ipm_d2L(self.h, ptr(W)))
        return W

    def merit(self):
        p, dp_ = C.c_double(), C.c_double()
        check(self.lib.b200ipm_merit(self.h, C.byref(p), C.byref(dp_)))
        return p.value, dp_.value

    def init_slack(self):
        check(self.lib.b200ipm_init_slack(self.h))

    def init_lambda(self):
        check(self.lib.b200ipm_init_lambda(self.h))

    def update_mu(self):
        v = C.c_double()
        check(self.lib.b200ipm_update_mu(self.h, C.byref(v)))
        return v.value

    def direction(self, want_dz=True):
        dz = np.empty(self.K) if want_dz else None
        info = StepInfo()
        check(self.lib.b200ipm_direction(self.h, ptr(dz), C.byref(info)))
        return dz, info

    def step_max(self):
        a, b = C.c_double(), C.c_double()
        check(self.lib.b200ipm_step_max(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def newton_step(self):
        info = StepInfo()
        check(self.lib.b200ipm_newton_step(self.h, C.byref(info)))
        return info

    # ---- L-BFGS mode (pyipm.py:993-1371)
    def lbfgs_init(self, m, zeta=1.0):
        check(self.lib.b200ipm_lbfgs_init(self.h, int(m), float(zeta)))

    def lbfgs_update(self, gradx_old=None):
        gradx_old = f64(gradx_old)
        check(self.lib.b200ipm_lbfgs_update(self.h, ptr(gradx_old)))

    def lbfgs_direction(self, want_dz=True):
        dz = np.empty(self.K) if want_dz else None
        info = StepInfo()
        check(self.lib.b200ipm_lbfgs_direction(self.h, ptr(dz), C.byref(info)))
        return dz, info

    def lbfgs_step(self, do_update):
        info = StepInfo()
        check(self.lib.b200ipm_lbfgs_step(self.h, 1 if do_update else 0, C.byref(info)))
        return info

    def lbfgs_state(self):
        m, fail, zeta = C.c_int(), C.c_int(), C.c_double()
        check(self.lib.b200ipm_lbfgs_state(self.h, C.byref(m), C.byref(zeta), C.byref(fail)))
        return m.value, zeta.value, fail.value

    def sync(self):
        check(self.lib.b200ipm_sync(self.h))

    def state_save(self):
        check(self.lib.b200ipm_state_save(self.h))

    def state_restore(self):
        check(self.lib.b200ipm_state_restore(self.h))

    def profile_kernel(self, which, reps=10):
        """-> (ms per launch, algorithmic work per launch [FLOPs or bytes])"""
        ms, wk = C.c_float(), C.c_double()
        check(self.lib.b200ipm_profile_kernel(self.h, int(which), int(reps), C.byref(ms), C.byref(wk)))
        return ms.value, wk.value


class DenseLDLT(object):
    """Generic dense symmetric-indefinite factor/solve (sym_solve_cmp slot; BASELINE config 4 building block)."""

    def __init__(self, n, device=0, stream=None):
        self.lib = load()
        self.n = int(n)
        self.h = C.c_void_p()
        check(self.lib.b200ipm_ldlt_create(self.n, int(device), C.c_void_p(stream) if stream else None, C.byref(self.h)))

    def close(self):
        if getattr(self, 'h', None) is not None and self.h.value:
            self.lib.b200ipm_ldlt_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def factor(self, A):
        A = f64(A)
        assert A.shape == (self.n, self.n)
        inertia = (C.c_int * 3)()
        rc = C.c_double()
        check(self.lib.b200ipm_ldlt_factor(self.h, ptr(A), self.n, 0, inertia, C.byref(rc)))
        return tuple(inertia), rc.value

    def solve(self, B, nrefine=2):
        """B: (n,) or (n, nrhs); returns the solution with the same shape."""
        B = np.asarray(B, dtype=np.float64)
        one = (B.ndim == 1)
        Bt = np.array(B.reshape(self.n, -1).T, dtype=np.float64, order='C', copy=True)   # rhs-major private copy
        check(self.lib.b200ipm_ldlt_solve(self.h, ptr(Bt), Bt.shape[0], int(nrefine), 0))
        X = Bt.T
        return X[:, 0].copy() if one else np.ascontiguousarray(X)


def launch_count():
    return int(load().b200ipm_launch_count())


def test_syrk(n, Cin, beta, dadd, shift, terms, force_simple=False):
    """terms: list of (A (n x K), w (K) or None, alpha).  Returns (C, ms)."""
    lib = load()
    nt = len(terms)
    As = [f64(t[0]) for t in terms]
    ws = [f64(t[1]) for t in terms]
    Ap = (C.c_void_p * max(nt, 1))(*[ptr(a) for a in As])
    wp = (C.c_void_p * max(nt, 1))(*[ptr(w) if w is not None else None for w in ws])
    Ks = (C.c_int * max(nt, 1))(*[a.shape[1] for a in As])
    al = (C.c_double * max(nt, 1))(*[float(t[2]) for t in terms])
    Cin, dadd = f64(Cin), f64(dadd)
    out = np.empty((n, n))
    ms = C.c_float()
    check(lib.b200ipm_test_syrk(int(n), ptr(Cin), float(beta), ptr(dadd), float(shift), nt, Ap, wp, Ks, al, ptr(out),
                                1 if force_simple else 0, C.byref(ms)))
    return out, ms.value


def test_syrk_i8(n, Cin, beta, dadd, shift, terms, signed_mask=0, variant=0, lbo=0, sbo=0):
    """tcgen05 int8 (Ozaki) version of test_syrk.  Returns (C, (ms_slicing, ms_total), err_word)."""
    lib = load()
    nt = len(terms)
    As = [f64(t[0]) for t in terms]
    ws = [f64(t[1]) for t in terms]
    Ap = (C.c_void_p * max(nt, 1))(*[ptr(a) for a in As])
    wp = (C.c_void_p * max(nt, 1))(*[ptr(w) if w is not None else None for w in ws])
    Ks = (C.c_int * max(nt, 1))(*[a.shape[1] for a in As])
    al = (C.c_double * max(nt, 1))(*[float(t[2]) for t in terms])
    Cin, dadd = f64(Cin), f64(dadd)
    out = np.empty((n, n))
    ms = (C.c_float * 2)()
    err = C.c_int(0)
    check(lib.b200ipm_test_syrk_i8(int(n), ptr(Cin), float(beta), ptr(dadd), float(shift), nt, Ap, wp, Ks, al, ptr(out),
                                   int(signed_mask), int(variant), int(lbo), int(sbo), ms, C.byref(err)))
    return out, (ms[0], ms[1]), err.value


def trace_start():
    check(load().b200ipm_trace_start())


def trace_dump(maxrec=65536):
    """-> (id, blk, t0, t1, tag) arrays of the device-side timeline records since trace_start()"""
    ids = np.zeros(maxrec, dtype=np.int32)
    blk = np.zeros(maxrec, dtype=np.int32)
    t0 = np.zeros(maxrec, dtype=np.uint64)
    t1 = np.zeros(maxrec, dtype=np.uint64)
    tag = np.zeros(maxrec, dtype=np.uint64)
    n = C.c_int(0)
    check(load().b200ipm_trace_dump(ptr(ids), ptr(blk), ptr(t0), ptr(t1), ptr(tag), int(maxrec), C.byref(n)))
    k = n.value
    return ids[:k], blk[:k], t0[:k].astype(np.int64), t1[:k].astype(np.int64), tag[:k]


def test_gemv(A, v, transpose=False):
    lib = load()
    A, v = f64(A), f64(v)
    rows, cols = A.shape
    y = np.empty(cols if transpose else rows)
    check(lib.b200ipm_test_gemv(rows, cols, ptr(A), ptr(v), ptr(y), 1 if transpose else 0))
    return y
