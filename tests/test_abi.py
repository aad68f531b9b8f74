"""CPU-side checks of the C-ABI boundary: the library builds/loads, exports every symbol include/b200ipm.h
declares, and the product refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from pyipm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, 'include', 'b200ipm.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(b200ipm_[A-Za-z0-9_]+)\s*\(', txt)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_symbols():
        assert hasattr(raw, name), name
    assert lib.b200ipm_version() == 103


def test_struct_layouts_match_header():
    # sizes implied by include/b200ipm.h (10 doubles + 4 ints; 18 doubles + 10 ints + 8 floats + 6 ints) and the
    # sizes the compiled library reports
    assert ctypes.sizeof(_lib.Params) == 10 * 8 + 4 * 4
    assert ctypes.sizeof(_lib.StepInfo) == 18 * 8 + 10 * 4 + 8 * 4 + 6 * 4
    lib = _lib.load()
    assert lib.b200ipm_struct_size(0) == ctypes.sizeof(_lib.Params)
    assert lib.b200ipm_struct_size(1) == ctypes.sizeof(_lib.StepInfo)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    with pytest.raises(_lib.B200Error, match='no CUDA device'):
        _lib.Engine(3, 1, 3)
    with pytest.raises(_lib.B200Error, match='no CUDA device'):
        _lib.DenseLDLT(8)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'pyipm_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), fn


def test_integration_md_stub_matches_the_header():
    """The ctypes stub printed in INTEGRATION.md must stay in step with include/b200ipm.h: its struct mirrors are executed
    here and their sizes compared with the library's."""
    import ctypes as C
    import re
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    m = re.search(r"class _Params\(C\.Structure\):.*?(?=\nassert _b200)", text, re.S)
    assert m, 'stub not found'
    import numpy as np
    ns = {'C': C, 'np': np}
    exec(m.group(0), ns)
    lib = _lib.load()
    assert lib.b200ipm_struct_size(0) == C.sizeof(ns['_Params'])
    assert lib.b200ipm_struct_size(1) == C.sizeof(ns['_StepInfo'])
