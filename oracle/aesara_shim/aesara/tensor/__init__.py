import numpy as np
from .._expr import Expr, Var, Sub
from . import nlinalg, slinalg, basic  # noqa: F401


def scalar(name=None): return Var(name, 0)
def vector(name=None): return Var(name, 1)
def matrix(name=None): return Var(name, 2)
def _w(fn): return lambda *a: Expr(fn, a)


sum = _w(lambda a: np.sum(a))
log = _w(np.log)
abs_ = _w(np.abs)
dot = _w(np.dot)
eye = _w(lambda n: np.eye(int(n)))
zeros = _w(lambda shp: np.zeros(shp))
ones = _w(lambda shp: np.ones(shp))
triu = _w(np.triu)
diag = _w(np.diag)
diagonal = _w(np.diagonal)


def max(a, axis=None): return Expr(lambda v: np.max(v, axis=axis), (a,))
def min(a, axis=None): return Expr(lambda v: np.min(v, axis=axis), (a,))


le = _w(lambda a, b: a <= b)
gt = _w(lambda a, b: a > b)


def concatenate(lst, axis=0):
    return Expr(lambda *v: np.concatenate(v, axis=axis), tuple(lst))


def _upd(sub, val, inc):
    assert isinstance(sub, Sub), 'inc_subtensor / set_subtensor need an indexed expression'

    def fn(arr, idx, v):
        out = np.array(arr, copy=True)
        if inc:
            out[idx] += v
        else:
            out[idx] = v
        return out
    return Expr(fn, (sub.base, sub.idx, val))


def set_subtensor(sub, val): return _upd(sub, val, False)
def inc_subtensor(sub, val): return _upd(sub, val, True)


def grad(*a, **k):
    raise NotImplementedError('no autodiff in the aesara stand-in: pass df/d2f/dce/... explicitly')
