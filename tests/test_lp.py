"""Problem containers (reference tests/test_lp.py restated against pycllp_b200.lp)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose
from scipy.sparse import csc_matrix

from pycllp_b200.lp import SparseMatrix, StandardLP, GeneralLP, EqualityLP
import problems


def _counts(A):
    return A.nrows, A.ncols, A.nnzeros, A.nproblems


def test_sparse_matrix_init_variants():
    assert _counts(SparseMatrix()) == (0, 0, 0, 1)
    B = np.arange(1, 7, dtype=np.float32).reshape(3, 2)
    assert _counts(SparseMatrix(matrix=csc_matrix(B))) == (3, 2, 6, 1)
    assert _counts(SparseMatrix([0, 0, 1, 1, 2, 2], [0, 1, 0, 1, 0, 1], np.arange(6))) == (3, 2, 6, 1)
    with pytest.raises(ValueError):
        SparseMatrix([0, 1], [0], [1.0, 2.0])


def test_set_and_delete_values():
    A = SparseMatrix()
    A.set_value(0, 0, 1.0)
    assert _counts(A) == (1, 1, 1, 1)
    A.set_value(0, 1, 1.0)
    assert _counts(A) == (1, 2, 2, 1)
    A.set_value(0, 1, 3.0)          # overwrite, not append
    assert A.nnzeros == 2 and A.todense()[0, 1] == 3.0
    A._del_value(0, 0)
    assert _counts(A) == (1, 2, 1, 1)   # shape is inferred from the largest index
    A._del_value(0, 1)
    assert _counts(A) == (0, 0, 0, 1)
    with pytest.raises(ValueError):
        A.set_value(-1, 0, 1.0)


def test_rows_and_cols():
    A = SparseMatrix()
    assert A.add_row([0, 2, 3], [1.0, 1.0, 1.0]) == 0
    assert _counts(A) == (1, 4, 3, 1)
    assert A.add_row([1], [1.0]) == 1
    assert_allclose(A.todense(0), [[1, 0, 1, 1], [0, 1, 0, 0]])
    A._del_row(1)
    assert _counts(A) == (1, 4, 3, 1)
    A._del_row(0)
    assert _counts(A) == (0, 0, 0, 1)
    assert A.add_col([0, 2, 3], [1.0, 1.0, 1.0]) == 0
    assert A.add_col([1], 2.0) == 1
    assert _counts(A) == (4, 2, 4, 1)
    A.update_col(1, [0, 1], [5.0, 6.0])      # the reference deletes a ROW here (lp.py:261)
    assert_allclose(np.asarray(A.todense())[:, 1], [5, 6, 0, 0])
    A.update_row(0, [0, 1], [7.0, 8.0])
    assert_allclose(np.asarray(A.todense())[0], [7, 8])
    vals, rows, ptr = A.tocsc_arrays()
    assert ptr[-1] == A.nnzeros and vals.shape == (1, A.nnzeros)
    assert_allclose(csc_matrix((vals[0], rows, ptr)).toarray(), A.todense())


def test_standard_lp_rows_cols():
    lp = StandardLP()
    assert (lp.nrows, lp.ncols, lp.nnzeros, lp.nproblems) == (0, 0, 0, 1)
    assert lp.add_row([0, 2, 3], [1.0, 1.0, 1.0], 1.0) == 0
    assert lp.c.shape[1] == 4 and lp.b.shape[1] == 1
    lp.set_bound(0, 2.0)
    assert_allclose(lp.b, [[2.0]])
    with pytest.raises(ValueError):
        lp.set_bound(1, 1.0)
    cols, value, bound = lp.get_row(0)
    assert_allclose(cols, [0, 2, 3])
    assert_allclose(value, [[1, 1, 1]])
    assert bound == 2.0
    lp2 = StandardLP()
    assert lp2.add_col([0, 2, 3], [1.0, 1.0, 1.0], 1.0) == 0
    assert (lp2.nrows, lp2.ncols) == (4, 1)
    lp2.set_objective(0, 2.0)
    assert_allclose(lp2.c, [[2.0]])
    with pytest.raises(ValueError):
        lp2.set_objective(1, 1.0)


def test_batched_b_c_f():
    A, b, c, f = problems.small_problem()
    bb, cc = problems.perturb(b, c, 5)
    lp = StandardLP(A, bb, cc, f)
    assert lp.nproblems == 5 and lp.f.shape == (5,)
    lp1 = StandardLP(A, bb, c, f)          # a single c is shared by all problems
    assert lp1.c.shape == (5, 3)
    with pytest.raises(ValueError):
        StandardLP(A, bb, cc[:3], f)
    with pytest.raises(ValueError):
        StandardLP(A, bb)


def test_to_equality_form_matches_reference_layout():
    """lp.py:551-567: slack column ncols+row per row, value 1, objective 0."""
    lp, _ = problems.vanderbei_2_9()
    elp = lp.to_equality_form()
    assert isinstance(elp, EqualityLP) and (elp.nrows, elp.ncols) == (3, 6)
    D = np.asarray(elp.A.todense())
    assert_allclose(D[:, 3:], np.eye(3))
    assert_allclose(D[:, :3], np.asarray(lp.A.todense()))
    assert_allclose(elp.c, [[2, 3, 4, 0, 0, 0]])
    cols, vals, ub = lp.get_row(0)
    assert_allclose(cols, [1, 2]); assert_allclose(vals, [[2, 3]]); assert_allclose(ub, [5])
    cols, vals, ub = elp.get_row(0)
    assert_allclose(sorted(cols), [1, 2, 3])


def test_general_lp():
    lp = GeneralLP()
    assert (lp.nrows, lp.ncols, lp.nnzeros, lp.nproblems) == (0, 0, 0, 1)
    B = np.arange(1, 7, dtype=np.float32).reshape(3, 2)
    g = GeneralLP(SparseMatrix(matrix=csc_matrix(B)), np.zeros(3), np.zeros(2), f=0.0)
    assert (g.nrows, g.ncols, g.nnzeros) == (3, 2, 6)
    cols, vals = [0, 1, 2], [[1., 1., 1.]]
    lp.add_row(cols, vals, 2.0, np.inf)
    slp = lp.to_standard_form()
    scols, svals, sbound = slp.get_row(0)
    assert_allclose(cols, scols); assert_allclose(vals, -svals); assert_allclose([2.0], -sbound)
    with pytest.raises(IndexError):
        slp.get_row(1)
    lp.set_num_problems(3)                  # the reference raises NameError here (lp.py:720-723)
    assert lp.a.shape[0] == 3 and lp.l.shape[0] == 3
    with pytest.raises(ValueError):
        lp.set_col_bounds(0, lower_bound=-np.inf)


def test_general_to_standard_with_shift_and_upper_bounds():
    lp = GeneralLP()
    lp.add_row([0, 1], [1.0, 2.0], 1.0, 4.0)       # 1 <= x0 + 2 x1 <= 4
    lp.set_col_bounds(0, 1.0, 3.0)                 # 1 <= x0 <= 3
    lp.set_objective(0, 1.0); lp.set_objective(1, 1.0)
    s = lp.to_standard_form()
    D = np.asarray(s.A.todense())
    assert_allclose(D, [[-1, -2], [1, 2], [1, 0]])
    assert_allclose(s.b, [[0.0, 3.0, 2.0]])        # shifted by A l = 1
    assert_allclose(s.f, [1.0])


def test_registry_and_dispatch():
    from pycllp_b200.solvers import solver_registry, BaseSolver

    class Probe(BaseSolver):
        name = "probe_solver_for_test"

        def init(self, lp, verbose=0):
            self.seen = ("init", lp.nproblems)

        def solve(self, lp, verbose=0):
            return "solved"

    try:
        assert solver_registry["probe_solver_for_test"] is Probe
        lp, _ = problems.vanderbei_2_9()
        s = Probe()
        lp.init(s)
        assert s.seen == ("init", 1) and lp.solve(s) == "solved"
        assert {"cl_dense_primal_normal", "cl_sparse_primal_normal"} <= set(solver_registry)
        with pytest.raises(NotImplementedError):
            BaseSolver().init(lp)
    finally:
        solver_registry.pop("probe_solver_for_test", None)


def test_install_as_pycllp():
    import pycllp_b200
    pycllp_b200.install_as_pycllp()
    from pycllp.lp import StandardLP as S2
    from pycllp.solvers import solver_registry as reg2
    assert S2 is StandardLP and "cl_dense_primal_normal" in reg2


def test_random_problem_generator_matches_example():
    """examples/random_problem.py:12-27 -> m=50, n=100 in equality form (config 1)."""
    from pycllp_b200.problems import random_problem, random_equality_arrays
    from conftest import golden
    lp = StandardLP(*random_problem(50, 50, 0.1, 64)).to_equality_form()
    assert (lp.nrows, lp.ncols, lp.nproblems) == (50, 100, 64)
    A, b, c = random_equality_arrays(50, 50, 0.1, 64)
    assert_allclose(np.asarray(lp.A.todense()), A)
    assert_allclose(lp.b, b); assert_allclose(lp.c, c)
    g = golden("cfg1_dense")   # frozen copy (scipy's RNG stream is version dependent)
    if not np.array_equal(A, g["A"]):
        pytest.skip("scipy.sparse.rand stream differs from the one the fixture was made with")
