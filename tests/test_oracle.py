"""The oracle is pinned here (CPU only): against the reference's own kernels compiled as
C (bit for bit), against the committed golden vectors, and against the reference's
known-answer tests."""
import numpy as np
import pytest

from conftest import golden, assert_parity
import problems

BITWISE_CASES = ["cfg1_dense", "cfg2_vanderbei_2_9_dense", "cfg2_vanderbei_2_10_dense", "cfg2_small_dense"]
BITWISE_SPARSE = ["cfg1_sparse", "cfg2_vanderbei_2_9_sparse", "cfg2_vanderbei_2_10_sparse",
                  "cfg2_small_sparse", "cfg4_small_sparse"]


@pytest.mark.parametrize("name", BITWISE_CASES + ["cfg3_sample"])
def test_oracle_dense_equals_golden_bitwise(oracle, name):
    g = golden(name)
    r = oracle.solve_dense(g["A"], g["b"], g["c"])
    np.testing.assert_array_equal(r.status, g["status"])
    np.testing.assert_array_equal(r.iters, g["iters"])
    for k in "xyz":
        assert np.array_equal(r[k], g[k]), name + " " + k + " differs from the reference kernels"
    assert np.array_equal(r.trace, g["trace"])


@pytest.mark.parametrize("name", BITWISE_SPARSE)
def test_oracle_sparse_equals_golden_bitwise(oracle, name):
    g = golden(name)
    r = oracle.solve_sparse(g["A"], g["b"], g["c"])
    np.testing.assert_array_equal(r.status, g["status"])
    np.testing.assert_array_equal(r.iters, g["iters"])
    for k in "xyz":
        assert np.array_equal(r[k], g[k]), name + " " + k


def test_oracle_equals_reference_kernels_live(oracle, reference):
    """Same check against the reference compiled here (not the stored vectors), on fresh
    random inputs, dense and sparse."""
    rng = np.random.RandomState(11)
    m, n0, N = 30, 40, 6
    A = np.c_[rng.rand(m, n0) * (rng.rand(m, n0) < 0.3), np.eye(m)]
    b = 0.5 + rng.rand(N, m)
    c = np.c_[0.5 + rng.rand(N, n0), np.zeros((N, m))]
    for solve_o, solve_r in ((oracle.solve_dense, reference.solve_dense),
                             (oracle.solve_sparse, reference.solve_sparse)):
        ro, rr = solve_o(A, b, c), solve_r(A, b, c)
        np.testing.assert_array_equal(ro.status, rr.status)
        np.testing.assert_array_equal(ro.iters, rr.iters)
        for k in "xyz":
            assert np.array_equal(ro[k], rr[k])


def test_golden_matches_reference_kernels_live(reference):
    """The committed vectors are what the reference produces today."""
    g = golden("cfg1_dense")
    r = reference.solve_dense(g["A"], g["b"], g["c"])
    assert np.array_equal(r.x, g["x"]) and np.array_equal(r.status, g["status"])


def test_dense_sparse_agree(oracle):
    g = golden("cfg1_dense")
    rd = oracle.solve_dense(g["A"], g["b"], g["c"])
    rs = oracle.solve_sparse(g["A"], g["b"], g["c"])
    assert_parity(rs, rd, g["c"], "sparse vs dense")


@pytest.mark.parametrize("fn", [problems.vanderbei_2_9, problems.vanderbei_2_10])
def test_known_answers_vanderbei(oracle, fn):
    """tests/test_vanderbei.py:25-42 -- status 0 and x == xopt at 1e-6."""
    lp, xopt = fn()
    A, b, c = problems.equality_arrays(lp)
    for r in (oracle.solve_dense(A, b, c), oracle.solve_sparse(A, b, c)):
        assert r.status[0] == 0
        np.testing.assert_allclose(r.x[0, :len(xopt)], xopt, rtol=1e-6, atol=1e-6)


def test_known_answer_small_problem(oracle):
    """tests/test_simple.py:57-67."""
    A, b, c, f = problems.small_problem()
    Ad, bb, cc = problems.equality_arrays(problems.StandardLP(A, b, c, f))
    r = oracle.solve_dense(Ad, bb, cc)
    assert r.status[0] == 0
    np.testing.assert_allclose(r.x[0, :3], (1.00997e-13, 1.22527e-12, 5.18790e+00), rtol=1e-1, atol=1e-1)


@pytest.mark.parametrize("size", [10, 20])
def test_random_vs_highs(oracle, size):
    """tests/test_random.py:14-17 with GLPK replaced by HiGHS (GLPK is not in the image)."""
    from scipy.optimize import linprog
    lp = problems.helpers_random_problem(size, size, 1.0, 1)
    A, b, c = problems.equality_arrays(lp)
    r = oracle.solve_dense(A, b, c)
    assert r.status[0] == 0
    h = linprog(-c[0], A_eq=A, b_eq=b[0], bounds=(0, None), method="highs")
    assert h.status == 0
    np.testing.assert_allclose(r.x[0] @ c[0], -h.fun, rtol=1e-6)
    np.testing.assert_allclose(r.x[0, :size], h.x[:size], rtol=1e-3, atol=1e-3)


def test_kernel_solve_primal_normal(oracle):
    """tests/test_ldl.py:196-273: dy == reference kernel (bitwise) == numpy solve of the
    normal equations."""
    g = golden("kernel_solve_primal_normal")
    dy, L, D = oracle.solve_primal_normal(g["A"], g["x"], g["z"], g["y"], g["b"], g["c"], float(g["mu"]),
                                          want_factor=True)
    assert np.array_equal(dy, g["dy"])
    assert np.array_equal(D, g["D"])
    A = g["A"]
    for q in range(4):
        x, z, y, b, c = (g[k][q] for k in "xzybc")
        rhs = b - A @ x - (A * x / z) @ (c - A.T @ y + 1.0 / x)
        np.testing.assert_allclose(dy[q], np.linalg.solve((A * x / z) @ A.T, -rhs), rtol=1e-7, atol=1e-9)


def test_kernel_sparse_solve_primal_normal(oracle):
    g = golden("kernel_sparse_solve_primal_normal")
    dy = oracle.sparse_solve_primal_normal(g["A"], g["x"], g["z"], g["y"], g["b"], g["c"], float(g["mu"]))
    assert np.array_equal(dy, g["dy"])


def test_kernel_ldl(oracle):
    """tests/test_ldl.py:139-193: `ldl` and `modified_ldl` kernels; L D L' == A."""
    g = golden("kernel_ldl")
    L, D = oracle.ldl(g["AA"], modified=False)
    assert np.array_equal(L, g["L_plain"]) and np.array_equal(D, g["D_plain"])
    L2, D2 = oracle.ldl(g["AA"], modified=True, beta=float(g["beta"]), delta=1e-6)
    assert np.array_equal(L2, g["L_mod"]) and np.array_equal(D2, g["D_mod"])
    m = g["AA"].shape[1]
    Lf = np.zeros((m, m))
    Lf[np.tril_indices(m)] = L[0]
    np.testing.assert_allclose(Lf @ np.diag(D[0]) @ Lf.T, g["AA"][0], rtol=1e-10, atol=1e-10)


def test_infeasible_and_limits(oracle):
    """Status codes other than 0 (primal_normal.cl:256-269): an infeasible LP must not
    report optimal, and max_iter is honoured (status 5)."""
    A = np.array([[1.0, 1.0, 1.0, 0.0], [1.0, 1.0, 0.0, 1.0]])
    b = np.array([[-1.0, 2.0]])            # x1 + x2 + s = -1 with x, s >= 0: infeasible
    c = np.array([[1.0, 1.0, 0.0, 0.0]])
    r = oracle.solve_dense(A, b, c)
    assert r.status[0] != 0
    g = golden("cfg1_dense")
    p = oracle.default_params()
    p.max_iter = 3
    r = oracle.solve_dense(g["A"], g["b"][:2], g["c"][:2], params=p)
    assert (r.status == 5).all() and (r.iters == 3).all()


def test_oracle_a_reference_cpu_solver():
    """Oracle A = the reference's DensePrimalNormalSolver (normal_eqns.py restated in
    oracle/oracle_a.py + pycllp/_ldl.pyx compiled as is into oracle/_ref): known answers of
    tests/test_vanderbei.py:25-42 and the committed config-1 fixture."""
    from oracle import oracle_a
    if not oracle_a.available():
        pytest.skip("oracle/_ref/_ldl*.so not built (needs /root/reference and cython)")
    for fn in (problems.vanderbei_2_9, problems.vanderbei_2_10):
        lp, xopt = fn()
        A, b, c = problems.equality_arrays(lp)
        r = oracle_a.solve(A, b, c)
        assert r["status"][0] == 0
        np.testing.assert_allclose(r["x"][0, :len(xopt)], xopt, rtol=1e-6, atol=1e-6)
    g = golden("cfg1_oracle_a")
    r = oracle_a.solve(g["A"], g["b"][:1], g["c"][:1])
    np.testing.assert_allclose(r["x"][0] @ g["c"][0], g["x"][0] @ g["c"][0], rtol=1e-9)
    # same optimum as the OpenCL-kernel semantics (the parity target), to the accuracy its noisy
    # termination allows
    gc = golden("cfg1_dense")
    np.testing.assert_allclose(np.einsum("ij,ij->i", g["x"], g["c"]),
                               np.einsum("ij,ij->i", gc["x"][:4], gc["c"][:4]), rtol=1e-6)
