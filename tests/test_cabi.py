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
    return sorted(set(re.findall(r"\b(pycllp_b200_[a-z0-9_]+)\s*\(", text)))


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
    # the field order of pycllp_b200_params in the header, parsed from the header itself
    text = open(os.path.join(ROOT, "include", "pycllp_b200.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} pycllp_b200_params;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    decl = re.findall(r"\b(double|int)\s+([a-z_, ]+);", body)
    names = [n.strip() for _, group in decl for n in group.split(",")]
    kinds = [t for t, group in decl for _ in group.split(",")]
    assert [f[0] for f in Params._fields_] == names
    assert [f[1] for f in Params._fields_] == [ctypes.c_double if k == "double" else ctypes.c_int for k in kinds]
    assert names[:7] == ["eps", "delta", "r", "ldl_delta", "refine_tol", "max_iter", "max_refine"]
    assert ctypes.sizeof(Params) == 5 * 8 + 8 * 4 + 8


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
