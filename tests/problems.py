"""Known-answer LPs used by the reference's own tests, restated as plain arrays.

* vanderbei_2_9 / vanderbei_2_10 : tests/vanderbei_problems.py:6-37 (Vanderbei, Linear
  Programming, exercises 2.9 and 2.10) with their optimal x.
* small_problem : tests/test_simple.py:17-30.
* perturb : the batched perturbation of tests/test_simple.py:33-42
  (seed 0, b and c scaled by U(0.5, 1.5) per problem).
"""
import numpy as np
from scipy.sparse import csc_matrix

from pycllp_b200.lp import SparseMatrix, StandardLP, EqualityLP


def vanderbei_2_9():
    A = csc_matrix((np.array([1, 1, 2, 1, 2, 3, 2, 3], dtype=float),
                    np.array([1, 2, 0, 1, 2, 0, 1, 2]), np.array([0, 2, 5, 8]))).tocoo()
    b = np.array([5., 4., 7.])
    c = np.array([2., 3., 4.])
    xopt = np.array([1.5, 2.5, 0.])
    return StandardLP(SparseMatrix(matrix=A), b, c, 0.0), xopt


def vanderbei_2_10():
    A = csc_matrix((np.ones(4), np.zeros(4, dtype=int), np.array([0, 1, 2, 3, 4])))
    b = np.array([1.])
    c = np.array([6., 8., 5., 9.])
    xopt = np.array([0., 0., 0., 1.])
    return EqualityLP(SparseMatrix(matrix=A), b, c, 0.0), xopt


def small_problem():
    A = csc_matrix((np.array([3, 2, 2, 5, 1, 3], dtype=float), np.array([0, 1, 0, 1, 0, 1]),
                    np.array([0, 2, 4, 6]))).tocoo()
    c = np.array([1.10685436, 3.67678309, 2.04570983])
    b = np.array([5.187898, 16.76453246])
    return SparseMatrix(matrix=A), b, c, 0.0


def perturb(b, c, N=32, seed=0):
    np.random.seed(seed)
    bb = (0.5 + np.random.rand(N, len(b))) * b
    cc = (0.5 + np.random.rand(N, len(c))) * c
    return bb, cc


def equality_arrays(lp):
    """(A dense (m, n), b (N, m), c (N, n)) of an LP, converting StandardLP to equality form."""
    if isinstance(lp, StandardLP):
        lp = lp.to_equality_form()
    return np.asarray(lp.A.todense(), dtype=float), lp.b.copy(), lp.c.copy()


def helpers_random_problem(m, n, density, nproblems):
    """tests/helpers.py:35-60 -- row-wise sparse rows with >= 3 non-zeros, b, c ~ U(0, 1)."""
    from scipy.sparse import rand
    np.random.seed(0)
    A = np.empty((m, n))
    for i in range(m):
        A[i, :] = rand(1, n, density=max(density, 3. / n)).todense()
    A = SparseMatrix(matrix=csc_matrix(A))
    b = np.random.rand(nproblems, A.nrows)
    c = np.random.rand(nproblems, A.ncols)
    return StandardLP(A, b, c, 0.0)
