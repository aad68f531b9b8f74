"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for the `aesara` package.

The reference `/root/reference/pyipm.py` imports `aesara` at module top (pyipm.py:3-10); Aesara is not
installed and cannot be installed in the authoring container (no network).  This stand-in implements just
enough of the API surface that the reference's *precompiled-function* code path (pyipm.py:512-562 and the
six one-line expressions it still compiles with `theano.function`) executes UNMODIFIED, so that
`tests/golden/make_golden.py` can record real reference outputs as golden fixtures.

It is a lazy expression evaluator: symbolic variables and shared scalars build `Expr` nodes; `function()`
returns a `Function` that binds inputs and evaluates the graph with NumPy/SciPy.  The linear-algebra ops
forward to the same SciPy calls Aesara's own Ops forward to (`slinalg.Solve.perform` ->
`scipy.linalg.solve(A, b, assume_a=...)`; `slinalg.Eigvalsh.perform` -> `scipy.linalg.eigvalsh(a, b,
lower=True)`; `nlinalg.pinv` -> `numpy.linalg.pinv`).  No autodiff: `grad/hessian/jacobian` raise.

Never imported by the product (`pyipm_b200/`), by `bench.py`, or by any GPU test.
"""
from ._expr import Expr, Var, Shared, Function, function, shared  # noqa: F401
from . import tensor  # noqa: F401
from . import compile  # noqa: F401
from . import gradient  # noqa: F401
from . import ifelse  # noqa: F401
