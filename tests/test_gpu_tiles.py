"""The tile-sparse numeric factor (csrc/ipm_tiles.cuh: L computed and stored on its symbolic
fill pattern only) against the oracle's restatement of the reference's sparse kernels
(sparse_factor_primal_normal / sparse_forward_backward_primal_normal, ldl.cl:381-574, on the
Lindptr/Lindices pattern of cl.py:185-196).  Bar: same status per problem, objective 1e-8,
x, y, z 1e-6, same Newton step counts."""
import numpy as np
import pytest
from scipy.sparse import csr_matrix

from conftest import assert_parity, objective
from pycllp_b200.problems import staircase_equality_arrays, sparse_equality_arrays

pytestmark = pytest.mark.gpu


def _tiles(engine, A, b, c):
    engine.setup_sparse(csr_matrix(A), b.shape[0], factor="tiles")
    assert engine.sparse_info()["factor"] == "tiles"
    return engine.solve_host(b, c)


def test_tiles_vs_oracle_small_shapes(engine, oracle):
    """Forced onto small problems (the auto rule keeps these on the shared-memory dense kernels):
    staircase / random sparse / dense patterns, m not a multiple of the tile size."""
    rng = np.random.RandomState(3)
    cases = [("staircase 96", staircase_equality_arrays(96, 150, 24, 3, 5, seed=1)),
             ("staircase 77 (ragged)", staircase_equality_arrays(77, 120, 16, 4, 4, seed=2)),
             ("random sparse 100x150", sparse_equality_arrays(100, 150, 0.05, 6, seed=3)),
             ("tiny 5x7", sparse_equality_arrays(5, 7, 0.6, 3, seed=4))]
    m, n0 = 41, 30
    Ad = np.c_[rng.rand(m, n0), np.eye(m)]
    cases.append(("dense pattern 41", (csr_matrix(Ad), 0.5 + rng.rand(3, m),
                                       np.c_[0.5 + rng.rand(3, n0), np.zeros((3, m))])))
    for what, (A, b, c) in cases:
        ref = oracle.solve_sparse(A.toarray(), b, c)
        res = _tiles(engine, A, b, c)
        assert_parity(res, ref, c, what)
        np.testing.assert_array_equal(res["iters"], ref.iters, err_msg=what)


def test_tiles_hook_vs_oracle_sparse_solve_primal_normal(engine, oracle):
    """One normal-equations solve on a given interior state (the kernel the reference's tests
    launch directly, tests/test_ldl.py:352) through the tile factor."""
    A, b, c = staircase_equality_arrays(120, 200, 24, 3, 4, seed=5)
    rng = np.random.RandomState(9)
    N, (m, n) = 4, A.shape
    x, z = 0.1 + rng.rand(N, n), 0.1 + rng.rand(N, n)
    y = rng.rand(N, m)
    engine.setup_sparse(A, N, factor="tiles")
    dy = engine.solve_primal_normal(x, z, y, b, c, 0.3)
    ref = oracle.sparse_solve_primal_normal(A.toarray(), x, z, y, b, c, 0.3)
    np.testing.assert_allclose(dy, ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())


def test_tiles_genuinely_sparse_lp_vs_oracle(engine, oracle):
    """A staircase LP at m = 1500, n = 3750 whose factor is < 5 % dense: the auto rule picks the
    tile factor, its storage is ~ nnz(L) (no m x m array, no dense triangle), and the first LPs
    agree with the oracle's sparse path (a few seconds per LP on a host core)."""
    A, b, c = staircase_equality_arrays(1500, 2250, 40, 4, 6, seed=0)
    engine.setup_sparse(A, 6)                               # factor="auto"
    si = engine.sparse_info()
    assert si["factor"] == "tiles"
    assert si["factor_doubles"] < 0.10 * si["dense_factor_doubles"], si
    info = engine.info()
    m = A.shape[0]
    assert info["scratch_bytes"] / info["grid"] < 0.25 * m * m * 8      # far below one m x m matrix
    res = engine.solve_host(b, c)
    ref = oracle.solve_sparse(A.toarray(), b, c)            # ~0.5 s per LP on a host core
    assert_parity(res, ref, c, "staircase m=1500")
    np.testing.assert_array_equal(res["iters"], ref.iters)
    assert (res["status"] == 0).all()


def test_tiles_large_sparse_lp_properties(engine):
    """m = 6000, n = 15000 (L < 2 % dense; the dense kernels would need 144 MB of factor per LP
    and 72 GFLOP per factorisation): every LP optimal, residuals and gap below the stop
    tolerance, and identical to the dense-factor engine on a 1500-row problem is covered above."""
    A, b, c = staircase_equality_arrays(6000, 9000, 48, 4, 8, seed=1)
    engine.setup_sparse(A, 8)
    si = engine.sparse_info()
    assert si["factor"] == "tiles" and si["factor_doubles"] < 0.03 * si["dense_factor_doubles"], si
    res = engine.solve_host(b, c)
    eps = float(np.float32(1e-7))
    assert (res["status"] == 0).all(), res["status"]
    x, y, z = res["x"], res["y"], res["z"]
    assert np.linalg.norm(b - (A @ x.T).T, axis=1).max() < eps
    assert np.linalg.norm(c - (A.T @ y.T).T + z, axis=1).max() < eps
    assert np.einsum("ij,ij->i", x, z).max() < eps
    np.testing.assert_allclose(objective(x, c), np.einsum("ij,ij->i", y, b), rtol=0, atol=1e-5)


def test_tiles_mode_has_no_refinement(engine):
    A, b, c = staircase_equality_arrays(64, 100, 16, 3, 2, seed=6)
    engine.setup_sparse(A, 2, factor="tiles")
    with pytest.raises(RuntimeError):
        engine.set_params(max_refine=3)
    with pytest.raises(RuntimeError):
        engine.set_preset("py")
    engine.setup_sparse(A, 2, factor="dense")
    assert engine.sparse_info()["factor"] == "dense"
    engine.set_params(max_refine=3)


def test_sparse_ldl_hook_vs_dense_modified_ldl(engine, oracle):
    """tests/test_ldl.py:92-108 (test_sparse_modified_ldl): the modified LDL' of a sparse SPD matrix on
    the CSR-lower pattern of its factor equals the dense modified LDL' restricted to that pattern --
    dense side: the oracle's restatement of the reference's modified_ldl kernel (ldl.cl:57-107)."""
    from scipy.sparse import rand as sparse_rand, tril
    rng = np.random.RandomState(2)
    for m, n, dens in [(50, 100, 0.025), (97, 60, 0.03), (130, 300, 0.01)]:
        N = 3
        AA = np.empty((N, m, m))
        for q in range(N):
            A = sparse_rand(m, n, density=dens, random_state=rng).toarray()
            if q:
                A = A * (A0 != 0)                       # same pattern, other values
                A[A0 != 0] = rng.rand(int((A0 != 0).sum()))
            else:
                A0 = A
            AA[q] = A @ A.T + 1e-3 * np.eye(m)
        beta = float(np.sqrt(np.abs(AA[0]).max()))
        Lp, D = oracle.ldl(AA, modified=True, beta=beta)
        Ld = np.zeros((N, m, m))
        il = np.tril_indices(m)
        Ld[:, il[0], il[1]] = Lp
        pat = tril(csr_matrix((np.abs(Ld).sum(axis=0) != 0).astype(float)), format="csr")
        pat.sort_indices()
        assert pat.nnz < 0.6 * m * (m + 1) // 2             # (the factor is actually sparse)
        Ls, Ds = engine.sparse_ldl(AA, pat.indptr, pat.indices, beta=beta)
        rows = np.repeat(np.arange(m), np.diff(pat.indptr))
        np.testing.assert_allclose(Ds, D, rtol=1e-10)
        np.testing.assert_allclose(Ls, Ld[:, rows, pat.indices], rtol=1e-9, atol=1e-12)


def test_tiles_with_reordered_constraints_vs_oracle(engine, oracle):
    """The same staircase LP with its constraints SHUFFLED: in the order given the factor is dense;
    setup_sparse reorders the constraints (RCM) for the tile factor, b is read and y written through
    the permutation, and the caller sees the solution of the LP as posed -- equal to the oracle's
    (which factorises in the given order) in status, step count and x, y, z."""
    A0, b0, c = staircase_equality_arrays(640, 960, 24, 3, 5, seed=7)
    rng = np.random.RandomState(1)
    sh = rng.permutation(A0.shape[0])
    A, b = A0.tocsr()[sh], b0[:, sh]
    engine.setup_sparse(A, b.shape[0], factor="tiles", ordering="natural")
    nat = engine.sparse_info()
    engine.setup_sparse(A, b.shape[0])                       # auto: m > 512, RCM gives fewer tiles
    si = engine.sparse_info()
    assert si["factor"] == "tiles" and si["ordering"] == "rcm"
    assert si["factor_doubles"] < 0.25 * nat["factor_doubles"]
    res = engine.solve_host(b, c)
    ref = oracle.solve_sparse(A.toarray(), b, c)
    assert_parity(res, ref, c, "shuffled staircase, RCM")
    assert np.abs(res["iters"] - ref.iters).max() <= 1
    # warm start: y0 comes back through the permutation
    r2 = engine.solve_host(b, c, warm_start=True)
    assert (r2["status"] == 0).all() and r2["iters"].max() <= res["iters"].max()
    # the hook goes through the permutation as well (y, b in; dy out)
    N, (m, n) = b.shape[0], A.shape
    x, z, y = 0.1 + rng.rand(N, n), 0.1 + rng.rand(N, n), rng.rand(N, m)
    dy = engine.solve_primal_normal(x, z, y, b, c, 0.2)
    dref = oracle.sparse_solve_primal_normal(A.toarray(), x, z, y, b, c, 0.2)
    np.testing.assert_allclose(dy, dref, rtol=1e-8, atol=1e-9 * np.abs(dref).max())
