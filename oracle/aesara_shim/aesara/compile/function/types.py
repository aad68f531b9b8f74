from ..._expr import Function  # noqa: F401
