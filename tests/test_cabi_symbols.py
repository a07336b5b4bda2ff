"""The C-ABI shared library loads without a GPU, exports every symbol include/pasio_b200.h declares,
and refuses to create a context (no CPU fallback) when no device is present."""
import ctypes
import os
import re

import pytest

from pasio_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'pasio_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(pasio_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = _native.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name
        assert name in _native.SIGNATURES, 'ctypes prototype missing for %s' % name
    assert sorted(_native.SIGNATURES) == names
    assert lib.pasio_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    handle = ctypes.c_void_p()
    lib = _native.load_library()
    assert lib.pasio_ctx_create(0, ctypes.byref(handle)) != 0
    assert b'no CUDA device' in lib.pasio_last_error(None)
    with pytest.raises(RuntimeError):
        _native.Engine(0)
    import numpy as np
    import pasio_b200
    with pytest.raises(RuntimeError):
        list(pasio_b200.segments_with_scores(np.array([1, 2, 3]), pasio_b200.configure_splitter()))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'pasio_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
