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
