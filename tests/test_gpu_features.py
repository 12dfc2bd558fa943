"""GPU tests of everything around the cold-start solve: termination status on infeasible
instances (the reference's "unreliable" statuses, gated), warm start, the per-iteration trace,
the NaN guard, the 'py' preset, packed records, launches on several streams."""
import numpy as np
import pytest

from conftest import golden, assert_parity, objective
import problems

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------
# status parity on infeasible / unbounded instances
# ---------------------------------------------------------------------------------------
def test_status_on_infeasible_and_unbounded_instances(engine):
    """224 primal-infeasible / unbounded instances (statuses 2, 4, 5; m = 3 .. 100).  The fixture
    (tests/golden/make_golden_infeasible.py) records, per instance, the status of the reference
    algorithm AND the set of statuses it reaches when A, b, c are perturbed consistently by
    1e-15 and 1e-13 relative (2 x 64 samples).  Where the reference is stable the engine must return the
    same status; where the reference's own status flips under a few ulps of input noise the
    engine's status must be one of those it reaches."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location(
        "make_golden_infeasible", os.path.join(os.path.dirname(__file__), "golden", "make_golden_infeasible.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    g = golden("infeasible_status")
    nstable = nchaotic = 0
    hist = np.zeros(6, dtype=int)
    bad = []
    for kind in ("primal", "dual"):
        for m in g["sizes"]:
            A, b, c = gen.instances(kind, int(m))
            engine.setup_dense(A, b.shape[0])
            res = engine.solve_host(b, c)
            ref = g["%s_%d_status" % (kind, m)]
            allowed = g["%s_%d_allowed" % (kind, m)]
            stable = allowed.sum(axis=1) == 1
            for q in range(b.shape[0]):
                st = int(res["status"][q])
                hist[st] += 1
                if stable[q]:
                    nstable += 1
                    if st != ref[q]:
                        bad.append((kind, int(m), q, st, int(ref[q]), "stable"))
                    elif st == 5:
                        assert res["iters"][q] == 200
                else:
                    nchaotic += 1
                    if not allowed[q, st]:
                        bad.append((kind, int(m), q, st, np.flatnonzero(allowed[q]).tolist(), "chaotic"))
    assert nstable >= 150 and nstable + nchaotic == 224
    assert hist[2] > 0 and hist[4] > 0 and hist[5] > 100, hist
    assert not bad, bad


# ---------------------------------------------------------------------------------------
# warm start (README.md:6, primal_normal.cl:213-219)
# ---------------------------------------------------------------------------------------
def test_warm_start_matches_the_oracle_from_the_same_state_and_saves_steps(engine, oracle):
    """Parity definition of a warm start: the oracle started from the same x, z, y (the raw end
    point of the previous solve, primal_normal.cl:213-219) must end with the same status and the
    same optimum.  A converged point lies on the boundary, from where the method takes long,
    badly conditioned paths (up to ~130 steps here), so the solution tolerances are those of
    the reference's own batched test (1e-3, tests/test_simple.py:70-93) tightened to 1e-5 and
    the objective 1e-6 relative -- not the cold-start bar.  With warm_floor = 1e-2 (pull the
    start back inside) the restart is cheaper than a cold start."""
    from pycllp_b200.problems import random_equality_arrays
    A, b, c = random_equality_arrays(60, 60, 1.0, 24, seed=3)
    engine.setup_dense(A, 24)
    cold = engine.solve_host(b, c)
    assert (cold["status"] == 0).all()
    rng = np.random.RandomState(1)
    b2 = b * (1.0 + 0.01 * rng.standard_normal(b.shape))
    c2 = c * (1.0 + 0.01 * rng.standard_normal(c.shape))
    warm = engine.solve_host(b2, c2, warm_start=True)
    ref = oracle.solve_dense_ex(A, b2, c2, start=(cold["x"], cold["y"], cold["z"]))
    np.testing.assert_array_equal(warm["status"], ref.status)
    assert (warm["status"] == 0).all()
    np.testing.assert_allclose(objective(warm["x"], c2), objective(ref.x, c2), rtol=1e-6)
    for k in "xyz":
        np.testing.assert_allclose(warm[k], ref[k], rtol=1e-5, atol=1e-5)
    cold2 = engine.solve_host(b2, c2)
    np.testing.assert_allclose(objective(warm["x"], c2), objective(cold2["x"], c2), rtol=1e-6)
    # pulled back inside: fewer Newton steps than the cold start, same optimum
    engine.solve_host(b, c)
    engine.set_params(warm_floor=1e-2)
    shifted = engine.solve_host(b2, c2, warm_start=True)
    engine.set_params(warm_floor=0.0)
    assert (shifted["status"] == 0).all()
    np.testing.assert_allclose(objective(shifted["x"], c2), objective(cold2["x"], c2), rtol=1e-6)
    assert shifted["iters"].mean() < cold2["iters"].mean(), (shifted["iters"].mean(), cold2["iters"].mean())
    # a warm start without a resident previous solve of the same size is refused
    engine.setup_dense(A, 24)
    with pytest.raises(RuntimeError, match="warm start"):
        engine.solve_host(b2[:5], c2[:5], warm_start=True)


def test_warm_start_through_the_plugin_api(engine):
    from pycllp_b200.lp import StandardLP
    from pycllp_b200.problems import random_problem
    from pycllp_b200.solvers import solver_registry
    lp = StandardLP(*random_problem(40, 40, 0.3, 16)).to_equality_form()
    s = solver_registry["cl_dense_primal_normal"]()
    lp.init(s)
    with pytest.raises(RuntimeError):
        s.solve(lp, warm_start=True)
    lp.solve(s)
    x0, it0 = s.x.copy(), s.iterations.copy()
    lp.b *= 1.005
    s.solve(lp, warm_start=True)
    assert (s.status == 0).all()
    assert not np.array_equal(s.x, x0)
    lp.solve(s)                                   # cold again: same optimum
    cold_x = s.x.copy()
    s.solve(lp, warm_start=True)
    np.testing.assert_allclose(objective(s.x, lp.c), objective(cold_x, lp.c), rtol=1e-6)


# ---------------------------------------------------------------------------------------
# per-iteration trace (primal_normal.cl:250-252) and verbose output
# ---------------------------------------------------------------------------------------
def test_iteration_trace_matches_the_oracle(engine, oracle, capsys):
    g = golden("cfg1_dense")
    A, b, c = g["A"], g["b"][:8], g["c"][:8]
    engine.setup_dense(A, 8)
    res = engine.solve_host(b, c, trace_iters=200)
    ref = oracle.solve_dense_ex(A, b, c, want_trace=True)
    np.testing.assert_array_equal(res["iters"], ref.iters)
    for q in range(8):
        k = int(ref.iters[q]) + 1                     # the terminating iteration is traced too
        got, want = res["trace"][q, :k], ref.itrace[q, :k]
        big = want > 1e-9                             # (below that the residuals are rounding noise)
        np.testing.assert_allclose(got[big], want[big], rtol=1e-4)
        assert np.isnan(res["trace"][q, k:]).all()
    from pycllp_b200.lp import EqualityLP
    from pycllp_b200.solvers import solver_registry
    lp, _ = problems.vanderbei_2_10()
    s = solver_registry["cl_dense_primal_normal"]()
    lp.init(s, verbose=2)
    lp.solve(s, verbose=2)
    out = capsys.readouterr().out
    assert "|rho|:" in out and "gamma:" in out and "iterations:" in out and "status: 0" in out


# ---------------------------------------------------------------------------------------
# NaN guard and the 'py' preset (normal_eqns.py)
# ---------------------------------------------------------------------------------------
def test_nan_guard_gives_status_3(engine, oracle):
    g = golden("cfg2_small_dense")
    A, b, c = g["A"], g["b"][:4].copy(), g["c"][:4].copy()
    b[1, 0] = np.nan
    engine.setup_dense(A, 4)
    res = engine.solve_host(b, c)                     # preset 'cl': no NaN test, like the kernels
    ref = oracle.solve_dense(A, b, c)
    np.testing.assert_array_equal(res["status"], ref.status)
    assert res["status"][1] == 5 and res["iters"][1] == 200
    engine.set_params(nan_guard=1)
    res = engine.solve_host(b, c)
    assert res["status"][1] == 3 and res["iters"][1] == 0     # normal_eqns.py:85-87
    np.testing.assert_array_equal(np.delete(res["status"], 1), np.delete(ref.status, 1))


def test_python_preset_solver(engine):
    """solver_registry['dense_primal_normal'] = the constants of solvers/normal_eqns.py on the
    CUDA engine: returns the status array; known answers; same optimum as the 'cl' preset."""
    from pycllp_b200.solvers import solver_registry
    for fn in (problems.vanderbei_2_9, problems.vanderbei_2_10):
        lp, xopt = fn()
        elp = lp.to_equality_form() if isinstance(lp, problems.StandardLP) else lp
        s = solver_registry["dense_primal_normal"]()
        elp.init(s)
        status = elp.solve(s)
        assert status is s.status and status[0] in (0, 2)          # (noise-driven)
        np.testing.assert_allclose(s.x[0, :len(xopt)], xopt, rtol=1e-5, atol=1e-5)
    p = s.engine.get_params()
    assert (p.eps, p.delta, p.mu_mode, p.refine_mode, p.theta_floor, p.dz_mode, p.nan_guard) == \
           (1e-8, 0.1, 1, 1, 0, 1, 1)
    g = golden("cfg1_dense")
    engine.setup_dense(g["A"], 64)
    engine.set_preset("py")
    res = engine.solve_host(g["b"], g["c"])
    assert np.isin(res["status"], (0, 2)).all()
    np.testing.assert_allclose(objective(res["x"], g["c"]), objective(g["x"], g["c"]), rtol=1e-6)
    # against the reference's CPU solver itself (Oracle A: normal_eqns.py + the compiled _ldl.pyx;
    # fixture tests/golden/cfg1_oracle_a.npz): same optimum; its statuses are noise (2 here)
    ga = golden("cfg1_oracle_a")
    np.testing.assert_allclose(objective(res["x"][:4], ga["c"]), objective(ga["x"], ga["c"]), rtol=1e-6)
    assert np.isin(ga["status"], (0, 2)).all()
    engine.set_preset("cl")
    assert_parity(engine.solve_host(g["b"], g["c"]), g, g["c"], "back to the cl preset")


# ---------------------------------------------------------------------------------------
# device-side entry points
# ---------------------------------------------------------------------------------------
def test_packed_records_and_launches_on_two_streams(engine):
    import torch
    from pycllp_b200 import sharding
    g = golden("cfg1_dense")
    A, b, c = g["A"], g["b"], g["c"]
    m, n = A.shape
    N = 64
    engine.setup_dense(A, N)
    dev = torch.device("cuda", 0)
    d_b, d_c = torch.from_numpy(b).to(dev), torch.from_numpy(c).to(dev)
    rec = torch.zeros((N, engine.record_width), dtype=torch.float64, device=dev)
    engine.solve_device_packed(N, d_b.data_ptr(), d_c.data_ptr(), rec.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = sharding.unpack_records(rec.cpu().numpy(), m, n)
    assert_parity(got, g, c, "packed records")
    np.testing.assert_array_equal(got["iters"], g["iters"])
    # two launches of ONE engine on two different streams must not share the work counter
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    o1 = [torch.zeros((N, k), dtype=torch.float64, device=dev) for k in (n, m, n)]
    o2 = [torch.zeros((N, k), dtype=torch.float64, device=dev) for k in (n, m, n)]
    st = [torch.full((N,), -1, dtype=torch.int32, device=dev) for _ in range(4)]
    torch.cuda.synchronize()
    engine.solve_device(N, d_b.data_ptr(), d_c.data_ptr(), o1[0].data_ptr(), o1[1].data_ptr(), o1[2].data_ptr(),
                        st[0].data_ptr(), st[1].data_ptr(), s1.cuda_stream)
    engine.solve_device(N, d_b.data_ptr(), d_c.data_ptr(), o2[0].data_ptr(), o2[1].data_ptr(), o2[2].data_ptr(),
                        st[2].data_ptr(), st[3].data_ptr(), s2.cuda_stream)
    torch.cuda.synchronize()
    for o, s in ((o1, st[0]), (o2, st[2])):
        assert (s.cpu().numpy() == g["status"]).all()
        np.testing.assert_allclose(o[0].cpu().numpy(), g["x"], rtol=1e-6, atol=1e-6)
    # warm start on the packed records: nothing left to do, zero further steps
    engine.solve_device_packed(N, d_b.data_ptr(), d_c.data_ptr(), rec.data_ptr(),
                               torch.cuda.current_stream().cuda_stream, warm_start=True)
    torch.cuda.synchronize()
    again = sharding.unpack_records(rec.cpu().numpy(), m, n)
    assert (again["status"] == 0).all() and again["iters"].max() <= 1


def test_setup_rejects_bad_csr(engine):
    indptr = np.array([0, 2, 1], dtype=np.int32)
    with pytest.raises(RuntimeError, match="indptr"):
        engine._check(engine._lib.pycllp_b200_setup_sparse(
            engine._h, 2, 3, indptr.ctypes.data_as(engine._lib.pycllp_b200_setup_sparse.argtypes[3]),
            np.zeros(2, dtype=np.int32).ctypes.data_as(engine._lib.pycllp_b200_setup_sparse.argtypes[4]),
            np.ones(2).ctypes.data_as(engine._lib.pycllp_b200_setup_sparse.argtypes[5]), 1), "setup_sparse")
