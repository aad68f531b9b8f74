"""Symbolic front-end: SymPy expressions in, derivatives out -- the "just give me f, ce, ci" input mode of the
reference (auto-differentiated Aesara expressions, pyipm.py:83-146, 216-231, 473-509) without Aesara.

``lower(x, f, ce, ci)`` returns

* a :class:`pyipm_b200.problems.PolyProblem` when f / ce / ci are polynomials in ``x`` (plus, optionally, the
  elementwise ``c * sum_i x_i log(x_i + shift)`` term of example 6, pyipm.py:2027): value, gradient, Jacobians and
  Hessians are then generated ON THE DEVICE from the monomial table (``b200ipm_bind_poly``);
* otherwise a dict of NumPy callables ``f, df, d2f, ce, dce, d2ce, ci, dci, d2ci`` with the reference's
  precompiled-function conventions (``dce`` is D x M, ``dci`` is D x N, ``d2ce(x, lda)`` / ``d2ci(x, lda)`` take the
  FULL multiplier vector, pyipm.py:216-231), derived by symbolic differentiation and ``sympy.lambdify`` -- the engine's
  callable mode (``b200ipm_set_derivs``).

SymPy is imported lazily: the rest of the package does not depend on it.
"""
from __future__ import print_function

import numpy as np

from . import problems as _problems


def _sympy():
    import sympy
    return sympy


def is_symbolic(obj):
    """True for a SymPy expression / Matrix / list of expressions."""
    try:
        sympy = _sympy()
    except ImportError:
        return False
    if isinstance(obj, (sympy.Basic, sympy.MatrixBase)):
        return True
    if isinstance(obj, (list, tuple)) and obj and all(isinstance(e, sympy.Basic) for e in obj):
        return True
    return False


def _as_list(exprs):
    sympy = _sympy()
    if exprs is None:
        return []
    if isinstance(exprs, sympy.MatrixBase):
        return [sympy.sympify(e) for e in exprs]
    if isinstance(exprs, (list, tuple)):
        return [sympy.sympify(e) for e in exprs]
    return [sympy.sympify(exprs)]


def _poly_terms(expr, xs):
    """Monomial table of a polynomial expression, or None if it is not a polynomial in xs with numeric coefficients."""
    sympy = _sympy()
    try:
        poly = sympy.Poly(sympy.expand(expr), *xs)
    except sympy.PolynomialError:
        return None
    terms = []
    for powers, coeff in poly.terms():
        if coeff.free_symbols:
            return None
        try:
            c = float(coeff)
        except TypeError:
            return None
        facs = tuple((i, int(p)) for i, p in enumerate(powers) if p)
        terms.append((c, facs))
    if not terms:
        terms = []
    return terms


def _split_xlogx(expr, xs):
    """expr = polynomial + c * sum_i x_i log(x_i + shift)  ->  (polynomial part, (c, shift)) or (expr, None)."""
    sympy = _sympy()
    expr = sympy.expand(expr)
    addends = sympy.Add.make_args(expr)
    logs, rest = {}, []
    for a in addends:
        if not a.has(sympy.log):
            rest.append(a)
            continue
        coeff, factors = a.as_coeff_mul()
        m = None
        if len(factors) == 2:
            for u, v in (factors, factors[::-1]):
                if u in xs and isinstance(v, sympy.log):
                    sh = sympy.simplify(v.args[0] - u)
                    if not sh.free_symbols:
                        m = (xs.index(u), float(coeff), float(sh))
        if m is None:
            return expr, None
        if m[0] in logs:
            return expr, None
        logs[m[0]] = (m[1], m[2])
    if not logs:
        return expr, None
    if sorted(logs) != list(range(len(xs))):
        return expr, None
    vals = set(logs.values())
    if len(vals) != 1:
        return expr, None
    return sympy.Add(*rest), vals.pop()


def lower(x, f, ce=None, ci=None, name=None):
    """See the module docstring.  ``x``: sequence of SymPy symbols (the order defines the variable vector)."""
    sympy = _sympy()
    xs = list(x)
    assert xs and all(isinstance(s, sympy.Symbol) for s in xs), 'x must be a sequence of SymPy symbols'
    f = sympy.sympify(f)
    ces, cis = _as_list(ce), _as_list(ci)
    extra = (f.free_symbols | set().union(*[e.free_symbols for e in ces + cis])) - set(xs)
    assert not extra, 'expressions contain symbols that are not variables: %s' % sorted(str(s) for s in extra)

    # ---- polynomial (+ x log x) form: lowered to the device monomial table
    f_poly, xlogx = _split_xlogx(f, xs)
    f_terms = _poly_terms(f_poly, xs)
    ce_terms = [_poly_terms(e, xs) for e in ces]
    ci_terms = [_poly_terms(e, xs) for e in cis]
    if f_terms is not None and all(t is not None for t in ce_terms + ci_terms):
        return _problems.PolyProblem(len(xs), f_terms, ce_terms=ce_terms, ci_terms=ci_terms, xlogx=xlogx, name=name)

    # ---- general twice-differentiable expressions: symbolic derivatives, NumPy callables
    D = len(xs)

    def fn(expr):
        g = sympy.lambdify(xs, expr, modules='numpy')
        return lambda xv: g(*[float(t) for t in np.asarray(xv, dtype=np.float64).reshape(D)])

    def vec(exprs):                     # R^D -> R^len(exprs)
        fs = [fn(e) for e in exprs]
        return lambda xv: np.array([np.float64(g(xv)) for g in fs], dtype=np.float64)

    def mat(rows):                      # list of lists of expressions -> 2-D array
        fs = [[fn(e) for e in row] for row in rows]
        return lambda xv: np.array([[np.float64(g(xv)) for g in row] for row in fs], dtype=np.float64)

    def hess_of(expr):
        return [[sympy.diff(expr, a, b) for b in xs] for a in xs]

    out = dict(f=lambda xv, _g=fn(f): np.float64(_g(xv)),
               df=vec([sympy.diff(f, a) for a in xs]),
               d2f=mat(hess_of(f)))

    def constraint_block(exprs, offset):
        J = mat([[sympy.diff(e, a) for e in exprs] for a in xs])          # D x len(exprs): the reference's layout
        Hs = [mat(hess_of(e)) for e in exprs]

        def d2(xv, lda):
            lda = np.asarray(lda, dtype=np.float64)
            H = np.zeros((D, D))
            for j, hj in enumerate(Hs):
                H = H + lda[offset + j] * hj(xv)
            return H
        return vec(exprs), J, d2

    if ces:
        out['ce'], out['dce'], out['d2ce'] = constraint_block(ces, 0)
    if cis:
        out['ci'], out['dci'], out['d2ci'] = constraint_block(cis, len(ces))
    return out
