import numpy as np


class Expr(object):
    """Lazy node: `fn(*evaluated_args)`; leaves are Var (bound at call time), Shared, or constants."""
    __array_priority__ = 1000.0

    def __init__(self, fn, args):
        self.fn = fn
        self.args = args

    def eval(self, env):
        vals = [_ev(a, env) for a in self.args]
        return self.fn(*vals)

    # arithmetic
    def __add__(self, o): return Expr(lambda a, b: a + b, (self, o))
    def __radd__(self, o): return Expr(lambda a, b: b + a, (self, o))
    def __sub__(self, o): return Expr(lambda a, b: a - b, (self, o))
    def __rsub__(self, o): return Expr(lambda a, b: b - a, (self, o))
    def __mul__(self, o): return Expr(lambda a, b: a * b, (self, o))
    def __rmul__(self, o): return Expr(lambda a, b: b * a, (self, o))
    def __truediv__(self, o): return Expr(lambda a, b: a / b, (self, o))
    def __rtruediv__(self, o): return Expr(lambda a, b: b / a, (self, o))
    def __pow__(self, o): return Expr(lambda a, b: a ** b, (self, o))
    def __neg__(self): return Expr(lambda a: -a, (self,))

    def __getitem__(self, idx):
        return Sub(self, idx)

    @property
    def size(self):
        return Expr(lambda a: np.asarray(a).size, (self,))

    @property
    def shape(self):
        return _Shape(self)

    @property
    def T(self):
        return Expr(lambda a: a.T, (self,))

    def reshape(self, shp):
        return Expr(lambda a, sh: np.asarray(a).reshape(sh), (self, shp))

    def ravel(self):
        return Expr(lambda a: a.ravel(), (self,))


def _ev(a, env):
    """Evaluate a leaf / node; tuples, lists and slices may contain symbolic entries (shapes such as
    (nineq, 2 * m_lbfgs), slices such as [:m_lbfgs] -- the L-BFGS graph, pyipm.py:1007-1182)."""
    if isinstance(a, Expr):
        return a.eval(env)
    if isinstance(a, tuple):
        return tuple(_ev(x, env) for x in a)
    if isinstance(a, list):
        return [_ev(x, env) for x in a]
    if isinstance(a, slice):
        return slice(_ev(a.start, env), _ev(a.stop, env), _ev(a.step, env))
    return a


class Sub(Expr):
    """a[idx]; remembers base and index so that inc_subtensor / set_subtensor can rebuild the full array."""

    def __init__(self, base, idx):
        self.base = base
        self.idx = idx
        Expr.__init__(self, lambda a, i: a[i], (base, idx))


class Lazy(Expr):
    """ifelse(cond, a, b): only the selected branch is evaluated (aesara.ifelse is lazy; the branch not taken may be
    ill-defined, e.g. a solve with a 0 x 0 matrix while no L-BFGS pair is stored)."""

    def __init__(self, cond, a, b):
        self.cond, self.a, self.b = cond, a, b

    def eval(self, env):
        return _ev(self.a, env) if bool(_ev(self.cond, env)) else _ev(self.b, env)


class _Shape(object):
    def __init__(self, e):
        self.e = e

    def __getitem__(self, i):
        return Expr(lambda a: a.shape[i], (self.e,))


class Var(Expr):
    def __init__(self, name=None, ndim=1):
        self.name = name
        self.ndim = ndim

    def eval(self, env):
        return env[id(self)]


class Shared(Expr):
    def __init__(self, value, name=None):
        self.value = value
        self.name = name

    def get_value(self):
        return self.value

    def set_value(self, v):
        self.value = v

    def eval(self, env):
        return self.value


def shared(value, name=None):
    return Shared(value, name)


class Function(object):
    """Stand-in for aesara.compile.function.types.Function.  Either compiled from (inputs, outputs) or
    wrapping a plain Python callable (how the golden generator presents NumPy derivative callables as
    'precompiled functions', pyipm.py:216-221, 378-383)."""

    def __init__(self, inputs=None, outputs=None, pyfunc=None):
        self.inputs = inputs
        self.outputs = outputs
        self.pyfunc = pyfunc

    def __call__(self, *args):
        if self.pyfunc is not None:
            return self.pyfunc(*args)
        env = {id(v): np.asarray(a) for v, a in zip(self.inputs, args)}
        out = self.outputs.eval(env) if isinstance(self.outputs, Expr) else self.outputs
        return np.asarray(out)


def function(inputs=None, outputs=None, on_unused_input=None, **kw):
    return Function(inputs=inputs, outputs=outputs)
