from ._expr import Lazy


def ifelse(cond, a, b):
    return Lazy(cond, a, b)
