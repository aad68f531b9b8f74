"""pyipm_b200 -- B200-native (sm_100a) Newton-step engine behind the pyipm ``IPM(...).solve()`` API.

Only the per-iteration hot path of jkaardal/pyipm is re-implemented: Lagrangian-Hessian /
Jacobian assembly, condensed primal-dual KKT formation, inertia-corrected dense symmetric-indefinite factor/solve,
step rules and the merit-function line search, as hand-written CUDA kernels in libb200ipm.so (include/b200ipm.h).
"""
from .ipm import IPM  # noqa: F401
from . import problems  # noqa: F401

__all__ = ['IPM', 'problems']
