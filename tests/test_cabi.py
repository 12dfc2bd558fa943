"""The C-ABI library loads and exports every symbol include/pycllp_b200.h declares.
No compute calls here (no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pycllp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pycllp_b200_[a-z_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from pycllp_b200 import _cabi
    from pycllp_b200.build import build
    build()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), "missing export " + name
    assert set(_cabi.EXPORTS) == set(names)
    lib.pycllp_b200_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.pycllp_b200_version()


def test_params_struct_layout_matches_header():
    from pycllp_b200._cabi import Params
    assert [f[0] for f in Params._fields_] == ["eps", "delta", "r", "ldl_delta", "refine_tol",
                                               "max_iter", "max_refine"]
    assert ctypes.sizeof(Params) == 5 * 8 + 2 * 4


def test_no_cpu_fallback():
    """Without a CUDA device the product fails loudly instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pycllp_b200._cabi import Engine
    from pycllp_b200.solvers import solver_registry
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        Engine(0)
    with pytest.raises(RuntimeError):
        solver_registry["cl_dense_primal_normal"]()


def test_product_does_not_import_oracle():
    """Nothing under pycllp_b200/ may reference the checker."""
    pkg = os.path.join(ROOT, "pycllp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f) + " mentions the oracle"
