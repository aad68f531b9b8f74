import numpy as np
from .._expr import Expr


def diag(a): return Expr(np.diag, (a,))
