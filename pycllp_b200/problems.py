"""Synthetic problem generators of the reference's example / benchmark configs.

``random_problem`` restates ``examples/random_problem.py:12-27``; the ``*_arrays``
helpers build the same problems directly in equality form as plain arrays (the
benchmark shapes of BASELINE.json / SURVEY.md section 8(d)) without going through the
O(m nnz) Python of ``to_equality_form``.
"""
import numpy as np


def random_problem(m, n, density, nproblems, seed=0):
    """(A, b, c, f) for ``StandardLP(*random_problem(...))`` -- examples/random_problem.py:12-27."""
    from scipy.sparse import rand
    from .lp import SparseMatrix
    np.random.seed(seed)
    A = SparseMatrix(matrix=rand(m, n, density=density))
    m, n = A.nrows, A.ncols
    b = 0.5 + np.random.rand(nproblems, m)
    c = 0.5 + np.random.rand(nproblems, n)
    return A, b, c, 0.0


def random_equality_arrays(m, n0, density, nproblems, seed=0):
    """Same random LP as ``random_problem(m, n0, ...)`` + ``to_equality_form()`` as arrays:
    dense ``A = [A0 I]`` (m, n0+m), ``b`` (N, m), ``c = [c0 0]`` (N, n0+m)."""
    from scipy.sparse import rand
    np.random.seed(seed)
    A0 = rand(m, n0, density=density).toarray()
    b = 0.5 + np.random.rand(nproblems, m)
    c0 = 0.5 + np.random.rand(nproblems, n0)
    A = np.concatenate([A0, np.eye(m)], axis=1)
    c = np.concatenate([c0, np.zeros((nproblems, m))], axis=1)
    return A, b, c


def sparse_equality_arrays(m, n0, density, nproblems, seed=0):
    """Config 4 style: CSR ``A = [A0 I]`` with ``A0 = scipy.sparse.rand(m, n0, density)``."""
    from scipy.sparse import rand, identity, hstack
    np.random.seed(seed)
    A0 = rand(m, n0, density=density, format="csr")
    b = 0.5 + np.random.rand(nproblems, m)
    c0 = 0.5 + np.random.rand(nproblems, n0)
    A = hstack([A0, identity(m, format="csr")], format="csr")
    c = np.concatenate([c0, np.zeros((nproblems, m))], axis=1)
    return A, b, c


def staircase_equality_arrays(m, n0, band, per_col, nproblems, seed=0):
    """A genuinely sparse LP whose normal-equations factor stays sparse without reordering:
    column k of ``A0`` has ``per_col`` non-zeros U(0, 1) in the rows of a window of ``band`` rows
    centred on ``k m / n0`` (a multi-period / staircase structure), so ``A0 A0'`` is banded and
    so is its LDL' factor.  Returns CSR ``A = [A0 I]``, ``b`` (N, m), ``c = [c0 0]`` (N, n0+m)."""
    from scipy.sparse import csr_matrix, identity, hstack
    rng = np.random.RandomState(seed)
    rows, cols, vals = [], [], []
    for k in range(n0):
        centre = int(k * m / float(n0))
        lo = max(0, min(m - band, centre - band // 2))
        r = lo + rng.choice(min(band, m), size=min(per_col, band, m), replace=False)
        rows.extend(r.tolist())
        cols.extend([k] * len(r))
        vals.extend(rng.rand(len(r)).tolist())
    A0 = csr_matrix((vals, (rows, cols)), shape=(m, n0))
    b = 0.5 + rng.rand(nproblems, m)
    c0 = 0.5 + rng.rand(nproblems, n0)
    A = hstack([A0, identity(m, format="csr")], format="csr")
    c = np.concatenate([c0, np.zeros((nproblems, m))], axis=1)
    return A, b, c
