import scipy.linalg
from .._expr import Expr


def solve(a, b, assume_a='gen', lower=False, check_finite=True):
    # aesara.tensor.slinalg.Solve.perform -> scipy.linalg.solve(A, b, assume_a=..., lower=..., check_finite=...)
    return Expr(lambda A, B: scipy.linalg.solve(A, B, assume_a=assume_a, lower=lower, check_finite=check_finite),
                (a, b))


def eigvalsh(a, b, lower=True):
    # aesara.tensor.slinalg.Eigvalsh.perform -> scipy.linalg.eigvalsh(a=a, b=b, lower=lower)
    return Expr(lambda A, B: scipy.linalg.eigvalsh(a=A, b=B, lower=lower), (a, b))
