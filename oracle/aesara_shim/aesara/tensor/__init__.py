import numpy as np
from .._expr import Expr, Var
from . import nlinalg, slinalg, basic  # noqa: F401


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


def concatenate(lst, axis=0):
    return Expr(lambda *v: np.concatenate(v, axis=axis), tuple(lst))


def _set(a, idx, v, inc):
    def fn(arr, val):
        out = np.array(arr, copy=True)
        if inc:
            out[idx] += val
        else:
            out[idx] = val
        return out
    return fn


class _Sub(Expr):
    pass


def set_subtensor(sub, val):
    raise NotImplementedError('symbolic-expression input mode is not supported by the stand-in')


inc_subtensor = set_subtensor


def grad(*a, **k):
    raise NotImplementedError('no autodiff in the aesara stand-in: pass df/d2f/dce/... explicitly')
