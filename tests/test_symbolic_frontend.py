"""The SymPy front-end (pyipm_b200/symbolic.py; reference input mode pyipm.py:83-146, 216-231, 473-509): the ten example
problems written as expressions exactly as in pyipm.py:1920-2131 must lower to problems whose value / gradient /
Jacobian / Hessian callables agree with the hand-written monomial tables, and non-polynomial expressions must come back
as callables with correct derivatives."""
import numpy as np
import pytest
import sympy

from pyipm_b200 import problems, symbolic

EPS = float(np.finfo(np.float64).eps)


def example_expressions(k):
    """f, ce, ci of example k as SymPy expressions of x (transcribed from pyipm.py:1925-2119)."""
    n = {1: 2, 2: 2, 3: 2, 4: 2, 5: 2, 6: 6, 7: 3, 8: 3, 9: 2, 10: 3}[k]
    x = sympy.symbols('x0:%d' % n)
    ce = ci = None
    if k == 1:
        f = x[0] ** 2 - 4 * x[0] + x[1] ** 2 - x[1] - x[0] * x[1]
    elif k == 2:
        f = 100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2
    elif k == 3:
        f = -sum(x)
        ce = [x[0] ** 2 + x[1] ** 2 - 1]
    elif k == 4:
        f = -(x[0] ** 2) * x[1]
        ce = [x[0] ** 2 + x[1] ** 2 - 3]
    elif k == 5:
        f = x[0] ** 2 + 2 * x[1] ** 2 + 2 * x[0] + 8 * x[1]
        ci = [x[0] + 2 * x[1] - 10, x[0], x[1]]
    elif k == 6:
        f = sum(xi * sympy.log(xi + sympy.Float(EPS)) for xi in x)
        ce = [sum(x) - 1]
        ci = list(x)
    elif k == 7:
        f = -x[0] * x[1] * x[2]
        ce = [sum(x) - 1]
        ci = [1.0 * xi for xi in x]
    elif k == 8:
        f = 4 * x[1] - 2 * x[2]
        ce = [2 * x[0] - x[1] - x[2] - 2, x[0] ** 2 + x[1] ** 2 - 1]
    elif k == 9:
        f = (x[0] - 2) ** 2 + 2 * (x[1] - 1) ** 2
        ci = [-x[0] - 4 * x[1] + 3, x[0] - x[1]]
    else:
        f = (x[0] - 1) ** 2 + 2 * (x[1] + 2) ** 2 + 3 * (x[2] + 3) ** 2
        ce = [x[2] - x[1] - x[0] - 1]
        ci = [x[2] - x[0] ** 2]
    return x, f, ce, ci


@pytest.mark.parametrize('k', range(1, 11))
def test_examples_lower_to_the_device_form(k):
    x, f, ce, ci = example_expressions(k)
    p = symbolic.lower(x, f, ce, ci)
    assert isinstance(p, problems.PolyProblem)
    ref, _ = problems.example_problem(k)
    assert (p.nvar, p.neq, p.nineq) == (ref.nvar, ref.neq, ref.nineq)
    assert (p.xlogx is None) == (ref.xlogx is None)
    rng = np.random.default_rng(k)
    for _ in range(3):
        xv = rng.uniform(0.2, 1.5, p.nvar)
        lda = rng.standard_normal(p.neq + p.nineq)
        np.testing.assert_allclose(p.f(xv), ref.f(xv), rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(p.df(xv), ref.df(xv), rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(p.d2f(xv), ref.d2f(xv), rtol=1e-13, atol=1e-13)
        if p.neq:
            np.testing.assert_allclose(p.ce(xv), ref.ce(xv), rtol=1e-13, atol=1e-13)
            np.testing.assert_allclose(p.dce(xv), ref.dce(xv), rtol=1e-13, atol=1e-13)
            np.testing.assert_allclose(p.d2ce(xv, lda), ref.d2ce(xv, lda), rtol=1e-13, atol=1e-13)
        if p.nineq:
            np.testing.assert_allclose(p.ci(xv), ref.ci(xv), rtol=1e-13, atol=1e-13)
            np.testing.assert_allclose(p.dci(xv), ref.dci(xv), rtol=1e-13, atol=1e-13)
            np.testing.assert_allclose(p.d2ci(xv, lda), ref.d2ci(xv, lda), rtol=1e-13, atol=1e-13)
    d = p.descriptor()                      # what b200ipm_bind_poly receives
    assert d['nterms'] == len(d['term_row']) and d['term_ptr'][-1] == (len(d['fac_var']) if d['nterms'] and d['term_ptr'][-1] else 0)


def test_general_expressions_become_callables_with_reference_conventions():
    x = sympy.symbols('a b c')
    f = sympy.exp(x[0]) * x[1] + sympy.cos(x[2]) + x[0] * x[2] ** 2
    ce = [sympy.sin(x[0]) + x[1] * x[2] - 0.3]
    ci = [x[0] - sympy.exp(-x[1]), x[2] ** 3 + 1, x[0] * x[1]]
    cb = symbolic.lower(x, f, ce, ci)
    assert isinstance(cb, dict) and set(cb) == {'f', 'df', 'd2f', 'ce', 'dce', 'd2ce', 'ci', 'dci', 'd2ci'}
    rng = np.random.default_rng(0)
    xv = rng.uniform(0.3, 1.2, 3)
    lda = rng.standard_normal(4)
    assert cb['dce'](xv).shape == (3, 1) and cb['dci'](xv).shape == (3, 3)      # transposed Jacobians (pyipm.py:117,138)
    h = 1e-6

    def fd_grad(fun):
        return np.array([(fun(xv + h * e) - fun(xv - h * e)) / (2 * h) for e in np.eye(3)])
    np.testing.assert_allclose(cb['df'](xv), fd_grad(cb['f']), rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(cb['d2f'](xv), fd_grad(cb['df']), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(cb['dce'](xv), fd_grad(cb['ce']), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(cb['dci'](xv), fd_grad(cb['ci']), rtol=1e-6, atol=1e-7)
    # d2ce / d2ci take the FULL multiplier vector (pyipm.py:223-231)
    np.testing.assert_allclose(cb['d2ce'](xv, lda), fd_grad(lambda t: cb['dce'](t) @ lda[:1]), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(cb['d2ci'](xv, lda), fd_grad(lambda t: cb['dci'](t) @ lda[1:]), rtol=1e-6, atol=1e-7)


def test_foreign_symbols_are_rejected():
    x = sympy.symbols('x0:2')
    y = sympy.Symbol('y')
    with pytest.raises(AssertionError):
        symbolic.lower(x, x[0] * y)


@pytest.mark.gpu
@pytest.mark.parametrize('k', [3, 7, 10])
def test_ipm_accepts_sympy_expressions(k):
    """IPM(x0, x_dev=symbols, f=expr, ce=..., ci=...) -- the reference's call shape (pyipm.py:2049-2054) -- solves the
    example on the device and lands on the reference's known answer and on the CPU oracle's iterate."""
    from oracle.pyipm_numpy import OracleIPM
    from pyipm_b200 import IPM
    x, f, ce, ci = example_expressions(k)
    ref, gts = problems.example_problem(k)
    x0 = problems.example_x0(k)
    o = OracleIPM(x0=x0.copy(), Ftol=1.0E-8, verbosity=-1, **ref.callables())
    with np.errstate(all='ignore'):
        xo, so, lo, fo, _ = o.solve()
    p = IPM(x0=x0.copy(), x_dev=x, f=f, ce=ce, ci=ci, Ftol=1.0E-8, verbosity=-1)
    xs, ss, ls, fs, _ = p.solve()
    assert p.signal == o.signal
    assert np.linalg.norm(xs - xo) <= 1e-6 * (1 + np.linalg.norm(xo))
    assert min(np.linalg.norm(xs - g) for g in gts) <= 1e-3
