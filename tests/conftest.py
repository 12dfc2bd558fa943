import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle.bindings import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's kernels compiled as C; only where oracle/_ref was built."""
    from oracle.bindings import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libpycllp_ref.so not built (needs /root/reference)")
    return Reference()


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine through the C ABI. GPU tests only -- fails loudly without a GPU."""
    from pycllp_b200._cabi import Engine
    return Engine(0)


def objective(x, c):
    return np.einsum("ij,ij->i", x, c)


def assert_parity(res, ref, c, what=""):
    """The north-star parity bar: same termination status per problem, objective within
    1e-8 relative, primal and dual solutions within 1e-6."""
    status = np.asarray(ref["status"])
    np.testing.assert_array_equal(np.asarray(res["status"]), status, err_msg=what + " status")
    ok = status == 0
    o_res, o_ref = objective(res["x"], c), objective(ref["x"], c)
    np.testing.assert_allclose(o_res[ok], o_ref[ok], rtol=1e-8, atol=1e-10, err_msg=what + " objective")
    for k in ("x", "y", "z"):
        np.testing.assert_allclose(np.asarray(res[k])[ok], np.asarray(ref[k])[ok], rtol=1e-6, atol=1e-6,
                                   err_msg=what + " " + k)
