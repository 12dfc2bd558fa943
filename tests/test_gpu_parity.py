"""Parity of the CUDA engine (through the C ABI) with the oracle / the reference's golden
vectors.  The bar (BASELINE.json north_star): same termination status per problem,
objective within 1e-8 relative, primal and dual solutions within 1e-6."""
import numpy as np
import pytest

from conftest import golden, assert_parity, objective
import problems

pytestmark = pytest.mark.gpu

DENSE_GOLDEN = ["cfg1_dense", "cfg2_vanderbei_2_9_dense", "cfg2_vanderbei_2_10_dense",
                "cfg2_small_dense", "cfg3_sample", "cfg5_sample"]
SPARSE_GOLDEN = ["cfg1_sparse", "cfg2_vanderbei_2_9_sparse", "cfg2_vanderbei_2_10_sparse",
                 "cfg2_small_sparse", "cfg4_small_sparse"]


def _dense(engine, A, b, c, **params):
    engine.setup_dense(A, b.shape[0])
    if params:
        engine.set_params(**params)
    return engine.solve_host(b, c)


def _sparse(engine, A, b, c):
    from scipy.sparse import csr_matrix
    engine.setup_sparse(csr_matrix(A), b.shape[0])
    return engine.solve_host(b, c)


@pytest.mark.parametrize("name", DENSE_GOLDEN)
def test_dense_vs_reference_golden(engine, name):
    g = golden(name)
    res = _dense(engine, g["A"], g["b"], g["c"])
    assert_parity(res, g, g["c"], name)
    assert np.abs(res["iters"] - g["iters"]).max() <= 1, "Newton step counts drifted"


@pytest.mark.parametrize("name", SPARSE_GOLDEN)
def test_sparse_vs_reference_golden(engine, name):
    g = golden(name)
    res = _sparse(engine, g["A"], g["b"], g["c"])
    assert_parity(res, g, g["c"], name)
    assert np.abs(res["iters"] - g["iters"]).max() <= 1


def test_dense_vs_oracle_random_shapes(engine, oracle):
    """Ragged shapes: m, n not multiples of the 8-wide tensor-core tiles / 64-wide macro tiles."""
    rng = np.random.RandomState(5)
    # (the low-rank shapes -- one or two dense columns next to the slacks -- sit on the boundary
    # of the theta-clamp test of the speculative factorisation and exercise its exact redo;
    # m = 203 runs the look-ahead factor with a ragged last panel and odd column count)
    for m, n0, dens, N in [(1, 3, 1.0, 4), (7, 5, 1.0, 9), (13, 29, 0.5, 17), (65, 70, 0.3, 12),
                           (130, 97, 1.0, 6), (24, 1, 1.0, 6), (40, 2, 1.0, 5), (203, 60, 1.0, 3)]:
        A0 = rng.rand(m, n0) * (rng.rand(m, n0) < dens)
        A = np.c_[A0, np.eye(m)]
        b = 0.5 + rng.rand(N, m)
        c = np.c_[0.5 + rng.rand(N, n0), np.zeros((N, m))]
        ref = oracle.solve_dense(A, b, c)
        assert_parity(_dense(engine, A, b, c), ref, c, "dense %dx%d" % (m, n0 + m))
        refs = oracle.solve_sparse(A, b, c)
        assert_parity(_sparse(engine, A, b, c), refs, c, "sparse %dx%d" % (m, n0 + m))


def test_factor_without_refinement_sweep(engine, oracle):
    """The sparse solver has no iterative refinement (ldl.cl:698-711), so any error of the
    factorisation shows up as extra Newton steps or a different status.  Sweep m over the
    panel/unit boundaries of the look-ahead factor (8-column panels, 16-row units, odd m,
    the largest m that keeps L in shared memory) and demand identical step counts."""
    rng = np.random.RandomState(11)
    for m in (9, 16, 17, 31, 40, 72, 97, 129, 177, 201, 206):
        n0, N = m // 3 + 1, 2
        A = np.c_[rng.rand(m, n0), np.eye(m)]
        b = 0.5 + rng.rand(N, m)
        c = np.c_[0.5 + rng.rand(N, n0), np.zeros((N, m))]
        ref = oracle.solve_sparse(A, b, c)
        res = _sparse(engine, A, b, c)
        assert_parity(res, ref, c, "sparse, dense A, m=%d" % m)
        np.testing.assert_array_equal(res["iters"], ref.iters, err_msg="m=%d" % m)


def test_known_answers_through_the_plugin_api(engine):
    """tests/test_vanderbei.py / test_simple.py through lp.init(solver) / lp.solve(solver)."""
    from pycllp_b200.solvers import solver_registry
    for fn in (problems.vanderbei_2_9, problems.vanderbei_2_10):
        lp, xopt = fn()
        elp = lp.to_equality_form() if isinstance(lp, problems.StandardLP) else lp
        for name in ("cl_dense_primal_normal", "cl_sparse_primal_normal"):
            solver = solver_registry[name]()
            elp.init(solver)
            assert elp.solve(solver) is None          # CL solvers return None (cl.py:85-124)
            np.testing.assert_equal(solver.status, 0)
            assert solver.x.shape == (1, elp.ncols) and solver.status.dtype == np.int32
            np.testing.assert_allclose(solver.x[0, :len(xopt)], xopt, rtol=1e-6, atol=1e-6)
    A, b, c, f = problems.small_problem()
    lp = problems.StandardLP(A, b, c, f).to_equality_form()
    solver = solver_registry["cl_dense_primal_normal"](None, None)   # (ctx, queue) positionals ignored
    lp.init(solver, verbose=0)
    lp.solve(solver, verbose=0)
    np.testing.assert_equal(solver.status, 0)
    np.testing.assert_allclose(solver.x[0, :3], (1.00997e-13, 1.22527e-12, 5.18790e+00), rtol=1e-1, atol=1e-1)


def test_batched_perturbed_small_problem_vs_highs(engine):
    """tests/test_simple.py:70-93 (32 perturbed problems), ground truth from HiGHS."""
    from scipy.optimize import linprog
    from pycllp_b200.solvers import solver_registry
    A, b, c, f = problems.small_problem()
    bb, cc = problems.perturb(b, c, 32)
    lp = problems.StandardLP(A, bb, cc, f).to_equality_form()
    for name in ("cl_dense_primal_normal", "cl_sparse_primal_normal"):
        solver = solver_registry[name]()
        lp.init(solver)
        lp.solve(solver)
        np.testing.assert_equal(solver.status, 0)
        Ad = np.asarray(lp.A.todense())
        for q in range(lp.nproblems):
            h = linprog(-lp.c[q], A_eq=Ad, b_eq=lp.b[q], bounds=(0, None), method="highs")
            np.testing.assert_allclose(solver.x[q] @ lp.c[q], -h.fun, rtol=1e-6)


@pytest.mark.parametrize("size", [10, 20])
def test_random_vs_highs(engine, size):
    from scipy.optimize import linprog
    lp = problems.helpers_random_problem(size, size, 1.0, 1)
    A, b, c = problems.equality_arrays(lp)
    res = _dense(engine, A, b, c)
    assert res["status"][0] == 0
    h = linprog(-c[0], A_eq=A, b_eq=b[0], bounds=(0, None), method="highs")
    np.testing.assert_allclose(res["x"][0] @ c[0], -h.fun, rtol=1e-6)
    np.testing.assert_allclose(res["x"][0, :size], h.x[:size], rtol=1e-3, atol=1e-3)


def test_kernel_solve_primal_normal(engine):
    """tests/test_ldl.py:219-273 at the reference's own tolerance (1e-5), plus numpy."""
    g = golden("kernel_solve_primal_normal")
    engine.setup_dense(g["A"], g["x"].shape[0])
    dy = engine.solve_primal_normal(g["x"], g["z"], g["y"], g["b"], g["c"], float(g["mu"]))
    np.testing.assert_allclose(dy, g["dy"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dy, g["dy"], rtol=1e-8, atol=1e-10)
    A = g["A"]
    x, z, y, b, c = (g[k][0] for k in "xzybc")
    rhs = b - A @ x - (A * x / z) @ (c - A.T @ y + 1.0 / x)
    np.testing.assert_allclose(dy[0], np.linalg.solve((A * x / z) @ A.T, -rhs), rtol=1e-7, atol=1e-9)


def test_kernel_sparse_solve_primal_normal(engine):
    """tests/test_ldl.py:276-361."""
    from scipy.sparse import csr_matrix
    g = golden("kernel_sparse_solve_primal_normal")
    engine.setup_sparse(csr_matrix(g["A"]), g["x"].shape[0])
    dy = engine.solve_primal_normal(g["x"], g["z"], g["y"], g["b"], g["c"], float(g["mu"]))
    np.testing.assert_allclose(dy, g["dy"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dy, g["dy"], rtol=1e-7, atol=1e-9)


def test_kernel_ldl(engine):
    """tests/test_ldl.py:139-193: `ldl`, `modified_ldl` at rtol 1e-6."""
    g = golden("kernel_ldl")
    L, D = engine.ldl(g["AA"], modified=False)
    np.testing.assert_allclose(D, g["D_plain"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(L, g["L_plain"], rtol=1e-6, atol=1e-7)
    L, D = engine.ldl(g["AA"], modified=True, beta=float(g["beta"]), delta=1e-6)
    np.testing.assert_allclose(D, g["D_mod"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(L, g["L_mod"], rtol=1e-6, atol=1e-7)


def test_modified_ldl_with_active_theta_clamp(engine, oracle):
    """A small beta makes (theta/beta)^2 exceed |D_j|: the engine's speculative panel
    elimination must detect it and fall back to the sequential rule (ldl.cl:368)."""
    g = golden("kernel_ldl")
    rng = np.random.RandomState(3)
    B = rng.rand(6, 77, 50)
    AA = np.concatenate([g["AA"], np.zeros((0, 40, 40))])          # 8 x 40 x 40, SPD
    for mats, beta in ((AA, 0.05), (np.einsum("nik,njk->nij", B, B), 0.5)):
        Lo, Do = oracle.ldl(mats, modified=True, beta=beta, delta=1e-6)
        L, D = engine.ldl(mats, modified=True, beta=beta, delta=1e-6)
        clamp_free = oracle.ldl(mats, modified=True, beta=1e9, delta=1e-6)[1]
        assert np.abs(Do - clamp_free).max() > 1e-3, "test does not exercise the clamp"
        np.testing.assert_allclose(D, Do, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(L, Lo, rtol=1e-7, atol=1e-10)


def test_status_codes_and_params(engine, oracle):
    """Non-optimal statuses agree with the oracle; max_iter gives status 5; empty batch ok."""
    A = np.array([[1.0, 1.0, 1.0, 0.0], [1.0, 1.0, 0.0, 1.0]])
    b = np.array([[-1.0, 2.0], [1.0, 2.0]])
    c = np.array([[1.0, 1.0, 0.0, 0.0], [1.0, 1.0, 0.0, 0.0]])
    ref = oracle.solve_dense(A, b, c)
    res = _dense(engine, A, b, c)
    np.testing.assert_array_equal(res["status"], ref.status)
    assert res["status"][0] != 0
    g = golden("cfg1_dense")
    res = _dense(engine, g["A"], g["b"][:3], g["c"][:3], max_iter=3)
    assert (res["status"] == 5).all() and (res["iters"] == 3).all()
    p = engine.get_params()
    assert p.max_iter == 3 and abs(p.eps - float(np.float32(1e-7))) < 1e-20
    with pytest.raises(TypeError):
        engine.set_params(nonsense=1)
    with pytest.raises(ValueError):
        engine.solve_host(g["b"][:3, :5], g["c"][:3])


def test_errors_are_loud(engine):
    from pycllp_b200._cabi import Engine
    e2 = Engine(0)
    with pytest.raises(RuntimeError, match="setup"):
        e2._check(e2._lib.pycllp_b200_solve_host(e2._h, 1, None, None, None, None, None, None, None), "solve")
    with pytest.raises(RuntimeError):
        Engine(10 ** 6)
    e2.close()


def test_device_buffers_and_stream(engine):
    """The device-pointer entry point used by the benchmark (torch only as the allocator)."""
    import torch
    g = golden("cfg1_dense")
    N, m = g["b"].shape
    n = g["c"].shape[1]
    engine.setup_dense(g["A"], N)
    dev = torch.device("cuda", 0)
    b = torch.from_numpy(g["b"]).to(dev)
    c = torch.from_numpy(g["c"]).to(dev)
    x = torch.empty(N, n, dtype=torch.float64, device=dev)
    y = torch.empty(N, m, dtype=torch.float64, device=dev)
    z = torch.empty(N, n, dtype=torch.float64, device=dev)
    st = torch.empty(N, dtype=torch.int32, device=dev)
    it = torch.empty(N, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev)
    engine.solve_device(N, b.data_ptr(), c.data_ptr(), x.data_ptr(), y.data_ptr(), z.data_ptr(),
                        st.data_ptr(), it.data_ptr(), stream.cuda_stream)
    stream.synchronize()
    res = dict(x=x.cpu().numpy(), y=y.cpu().numpy(), z=z.cpu().numpy(), status=st.cpu().numpy())
    assert_parity(res, g, g["c"], "device buffers")


def test_full_size_config3_properties(engine):
    """Config 3 at full size (m=200, n=400, N=4096): checked through size-independent
    properties -- every problem optimal, primal/dual residuals and the gap below the stop
    tolerance, x, z >= 0, weak duality gap closed -- and exactly against the golden sample
    (the first 8 problems of this very batch)."""
    from pycllp_b200.problems import random_equality_arrays
    A, b, c = random_equality_arrays(200, 200, 1.0, 4096)
    g = golden("cfg3_sample")
    if not np.array_equal(A, g["A"]):
        A, b, c = g["A"], np.tile(g["b"], (512, 1)), np.tile(g["c"], (512, 1))
    res = _dense(engine, A, b, c)
    assert (res["status"] == 0).all()
    x, y, z = res["x"], res["y"], res["z"]
    eps = float(np.float32(1e-7))
    assert np.linalg.norm(b - x @ A.T, axis=1).max() < eps
    assert np.linalg.norm(c - y @ A + z, axis=1).max() < eps
    assert np.einsum("ij,ij->i", x, z).max() < eps
    assert x.min() > 0 and z.min() > 0
    # primal objective c'x == dual objective b'y up to the gap (max c'x, A x = b; dual: A'y - z = c)
    np.testing.assert_allclose(objective(x, c), np.einsum("ij,ij->i", y, b), rtol=0, atol=1e-6)
    sub = {k: v[:8] for k, v in res.items()}
    assert_parity(sub, g, g["c"], "cfg3 head")
    assert 20 <= res["iters"].min() and res["iters"].max() <= 40


def _check_optimality(A, b, c, res, what):
    """Size-independent properties of an optimal primal-dual pair (max c'x, A x = b, x >= 0)."""
    eps = float(np.float32(1e-7))
    assert (res["status"] == 0).all(), what
    x, y, z = res["x"], res["y"], res["z"]
    At = A.T
    assert np.linalg.norm(b - (At.T @ x.T).T, axis=1).max() < eps, what
    assert np.linalg.norm(c - (At @ y.T).T + z, axis=1).max() < eps, what
    assert np.einsum("ij,ij->i", x, z).max() < eps, what
    assert x.min() > 0 and z.min() > 0, what
    np.testing.assert_allclose(objective(x, c), np.einsum("ij,ij->i", y, b), rtol=0, atol=1e-5, err_msg=what)


def test_full_size_config5_and_config4_properties(engine):
    """Configs 5 (dense m=500 n=1000) and 4 (sparse m=2000 n=5000, 1 % + slacks) at their full
    shapes, one wave of problems each (the batch only repeats the shape): every LP optimal with
    residuals and gap below the stop tolerance (exact parity at these shapes: the golden cases
    cfg5_sample and cfg4_small_sparse, test_large_problems_out_of_shared_memory)."""
    from pycllp_b200.problems import random_equality_arrays, sparse_equality_arrays
    A, b, c = random_equality_arrays(500, 500, 1.0, 148)
    res = _dense(engine, A, b, c)
    _check_optimality(A, b, c, res, "cfg5")
    assert np.array_equal(A, golden("cfg5_sample")["A"])      # (same A as the exact golden case)
    As, b, c = sparse_equality_arrays(2000, 3000, 0.01, 16, seed=0)
    res = _sparse(engine, As, b, c)
    _check_optimality(As.tocsr(), b, c, res, "cfg4")
    assert res["iters"].max() <= 80


def test_repeat_solves_are_cold_starts_and_deterministic(engine):
    """cl.py:108 re-initialises x = z = y = 1 on every solve; results are reproducible."""
    g = golden("cfg1_dense")
    engine.setup_dense(g["A"], 64)
    r1 = engine.solve_host(g["b"], g["c"])
    r2 = engine.solve_host(g["b"], g["c"])
    for k in ("x", "y", "z", "status", "iters"):
        assert np.array_equal(r1[k], r2[k])
    r3 = engine.solve_host(g["b"][:5], g["c"][:5])      # N < max_problems
    assert np.array_equal(r3["x"], r1["x"][:5])


def test_large_problems_out_of_shared_memory(engine, oracle):
    """Sizes where the factor (and then the vectors / work area) no longer fit in shared
    memory and live in the block's global scratch slot: m=520 dense against the oracle,
    m=1100 n=2600 sparse (config-4 shape, scaled) against the oracle's sparse path."""
    from pycllp_b200.problems import random_equality_arrays, sparse_equality_arrays
    A, b, c = random_equality_arrays(520, 300, 0.5, 3, seed=4)
    engine.setup_dense(A, 3)
    assert not engine.info()["factor_in_smem"]
    res = engine.solve_host(b, c)
    ref = oracle.solve_dense(A, b, c)
    assert_parity(res, ref, c, "dense m=520")

    As, b, c = sparse_equality_arrays(1100, 1500, 0.01, 4, seed=2)
    engine.setup_sparse(As, 4)
    info = engine.info()
    assert not info["factor_in_smem"] and info["smem_bytes"] < 160 * 1024
    rs = engine.solve_host(b, c)
    Ad = As.toarray()
    ref = oracle.solve_sparse(Ad, b, c)          # ~10 s per LP on a host core
    assert_parity(rs, ref, c, "sparse m=1100")
    assert np.abs(rs["iters"] - ref.iters).max() <= 1
    _check_optimality(As.tocsr(), b, c, rs, "sparse m=1100")


def test_super_panel_factor_shape_sweep(engine, oracle):
    """The factor in global memory (m > ~202: 64-column super-panels, 16x32 items, row solve of the
    16x64 units on the tensor pipe against the inverted 8x8 diagonal tiles) over its boundaries:
    ragged last row unit, ragged last super-panel, with and without a full pass of sixteen units,
    odd m.  No refinement on the sparse solver (ldl.cl:698-711), so an error of the factorisation
    shows up as a different step count."""
    rng = np.random.RandomState(23)
    for m in (209, 256, 271, 300, 337, 577):
        n0, N = m // 3 + 1, 2
        A = np.c_[rng.rand(m, n0), np.eye(m)]
        b = 0.5 + rng.rand(N, m)
        c = np.c_[0.5 + rng.rand(N, n0), np.zeros((N, m))]
        ref = oracle.solve_sparse(A, b, c)
        res = _sparse(engine, A, b, c)
        assert not engine.info()["factor_in_smem"]
        assert_parity(res, ref, c, "sparse, dense A, m=%d" % m)
        np.testing.assert_array_equal(res["iters"], ref.iters, err_msg="m=%d" % m)
        assert_parity(_dense(engine, A, b, c), oracle.solve_dense(A, b, c), c, "dense m=%d" % m)


def test_config4_at_its_named_shape_vs_reference_golden(engine):
    """Config 4 exactly as BASELINE.json names it (m=2000, n=5000, 1 % density + slacks): the
    first two LPs of the 1024-LP seed-0 workload against the outputs of the REFERENCE's own
    sparse kernels (tests/golden/make_golden_cfg4.py: 6 minutes per LP on a host core;
    oracle/ipm_oracle.c reproduced them bit for bit when the fixture was made)."""
    from scipy.sparse import csr_matrix
    g = golden("cfg4_sample")
    m, n = int(g["m"]), int(g["m"]) + int(g["n0"])
    A = csr_matrix((g["A_data"], g["A_indices"], g["A_indptr"]), shape=(m, n))
    assert bool(g["oracle_bit_identical"]) and str(g["source"]) == "reference"
    engine.setup_sparse(A, 2)
    assert not engine.info()["factor_in_smem"]
    res = engine.solve_host(g["b"], g["c"])
    assert_parity(res, g, g["c"], "cfg4 at m=2000")
    assert np.abs(res["iters"] - g["iters"]).max() <= 1, (res["iters"], g["iters"])
