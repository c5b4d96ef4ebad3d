"""CPU tests of the C-ABI boundary: the shared library loads, exports every symbol
declared in include/pysteps_b200.h, the ctypes signature table covers the header,
and the product path refuses to run (loudly) without a GPU instead of falling back."""
import ctypes
import os

import numpy as np
import pytest

from pysteps_b200 import _lib


def test_library_built_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run `make -C pysteps_b200/csrc` (or __graft_entry__.build())"
    lib = _lib.load()
    assert lib.b200_version() >= 100


def test_every_header_symbol_is_exported_and_bound():
    names = _lib.header_symbols()
    assert len(names) >= 7
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
        assert name in _lib._SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    for name in _lib._SIGNATURES:
        assert name in names, f"{name} bound in _lib.py but not declared in the header"


def test_last_error_is_a_string():
    lib = _lib.load()
    msg = lib.b200_last_error()
    assert isinstance(msg, bytes)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pysteps_b200
    f = pysteps_b200.extrapolation.get_method("semilagrangian")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        f(np.zeros((8, 8)), np.ones((2, 8, 8)), 1)


def test_product_package_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "pysteps_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(d, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboracle" not in text, f


def test_get_method_contract():
    # pysteps/tests/test_interfaces.py:58-91 (names, case-insensitivity, errors)
    from pysteps_b200.extrapolation import get_method, semilagrangian, eulerian_persistence
    assert get_method("semilagrangian") is semilagrangian.extrapolate
    assert get_method("SemiLagrangian") is semilagrangian.extrapolate
    assert get_method("eulerian") is eulerian_persistence
    assert get_method(None)(None, None, None) is None
    with pytest.raises(ValueError):
        get_method("nonexistent")
    precip = np.random.rand(10, 10)
    out = eulerian_persistence(precip, None, 3)
    assert out.shape == (3, 10, 10) and np.array_equal(out[2], precip)
    out, disp = eulerian_persistence(precip, None, [1, 2], return_displacement=True)
    assert out.shape == (2, 10, 10) and not disp.any()
