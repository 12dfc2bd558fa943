"""Oracle A -- the reference's CPU solver of the same algorithm, ``DensePrimalNormalSolver``
(``pycllp/solvers/normal_eqns.py:16-103``), test infrastructure and timed CPU baseline only.

The numerical half of that solver is Cython (``pycllp/_ldl.pyx``); ``oracle/Makefile`` (target
``ref_py``) compiles it unmodified from ``/root/reference`` into ``oracle/_ref/_ldl*.so``.  The
Python half, ``_solve`` (``normal_eqns.py:35-103``), is restated below line by line because it
cannot be imported as shipped (its package import needs pyopencl and GLPK).

Its termination status is driven by rounding noise (probe B.2: BLAS A'y in the
step against per-element np.dot in the right-hand side), so it serves as an OBJECTIVE cross-check
and as the "reference CPU solver" timing of BASELINE.json configs[0], not as a status oracle:
parity unpinned for status.
"""
import importlib.util
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EPS = 1.0e-8          # normal_eqns.py:12
MAX_ITER = 200        # :13


def available():
    return bool(glob.glob(os.path.join(HERE, "_ref", "_ldl*.so")))


def _ldl():
    path = glob.glob(os.path.join(HERE, "_ref", "_ldl*.so"))
    if not path:
        raise RuntimeError("oracle/_ref/_ldl*.so not built (make -C oracle ref_py; needs /root/reference)")
    spec = importlib.util.spec_from_file_location("_ldl", path[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def solve_one(A, b, c, ldl=None):
    """normal_eqns.py:35-103 for one problem. Returns (x, y, z, status, iterations)."""
    ldl = ldl or _ldl()
    m, n = A.shape
    x = np.ones(n)                                        # :44-46
    z = np.ones(n)
    y = np.ones(m)
    normr0 = sys.float_info.max                           # :49-50
    norms0 = sys.float_info.max
    delta = 0.1                                           # :52
    r = 0.9
    status = 5
    it = 0
    for it in range(MAX_ITER):                            # :57
        rho = b - np.dot(A, x)
        normr = np.sqrt(np.dot(rho, rho))
        sigma = c - np.dot(A.T, y) + z
        norms = np.sqrt(np.dot(sigma, sigma))
        gamma = np.dot(z, x)
        mu = delta * gamma / n                            # :65
        if normr < EPS and norms < EPS and gamma < EPS:   # :70-72
            status = 0
            break
        if normr > 10 * normr0 and normr > EPS:           # :74-76
            status = 2
            break
        if norms > 10 * norms0 and norms > EPS:           # :78-80
            status = 4
            break
        dy = ldl.solve_primal_normal(A, x, z, y, b, c, mu, delta=1e-6)   # :83
        if np.any(np.isnan(dy)):                          # :85-87
            status = 3
            break
        dx = (c - A.T.dot(y) + mu / x - A.T.dot(dy)) * x / z             # :89
        dz = (mu - x * z - z * dx) / x                                   # :90
        theta = max(np.max(-dx / x), np.max(-dz / z))                    # :92
        theta = min(r / theta, 1.0)
        x += theta * dx                                   # :95-97
        z += theta * dz
        y += theta * dy
        normr0 = normr
        norms0 = norms
    return x, y, z, status, it


def solve(A, b, c):
    """All problems, one after the other like the reference's Python loop (normal_eqns.py:30-31)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    b, c = np.atleast_2d(b), np.atleast_2d(c)
    ldl = _ldl()
    N = b.shape[0]
    m, n = A.shape
    out = dict(x=np.empty((N, n)), y=np.empty((N, m)), z=np.empty((N, n)),
               status=np.empty(N, dtype=np.int32), iters=np.empty(N, dtype=np.int32))
    for q in range(N):
        x, y, z, st, it = solve_one(A, np.array(b[q], dtype=np.float64), np.array(c[q], dtype=np.float64), ldl)
        out["x"][q], out["y"][q], out["z"][q], out["status"][q], out["iters"][q] = x, y, z, st, it
    return out
