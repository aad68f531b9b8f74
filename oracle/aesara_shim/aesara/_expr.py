import numpy as np


class Expr(object):
    """Lazy node: `fn(*evaluated_args)`; leaves are Var (bound at call time), Shared, or constants."""
    __array_priority__ = 1000.0

    def __init__(self, fn, args):
        self.fn = fn
        self.args = args

    def eval(self, env):
        vals = [a.eval(env) if isinstance(a, Expr) else a for a in self.args]
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
        return Expr(lambda a: a[idx], (self,))

    @property
    def shape(self):
        return _Shape(self)

    @property
    def T(self):
        return Expr(lambda a: a.T, (self,))

    def reshape(self, shp):
        return Expr(lambda a: a.reshape(shp), (self,))

    def ravel(self):
        return Expr(lambda a: a.ravel(), (self,))


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
