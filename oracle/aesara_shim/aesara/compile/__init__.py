# deliberately NO `function_module` attribute: pyipm.py:12-15 then falls through to compile.function.types
from . import function  # noqa: F401
