"""The small-problem kernels (csrc/ipm_small.cuh: 128 threads per LP, everything in shared
memory, several LPs per SM; and the two-blocks-per-SM build of the 512-thread kernel) against
the oracle: same status per problem, objective 1e-8, x, y, z 1e-6, same Newton step counts."""
import numpy as np
import pytest

from conftest import golden, assert_parity
from pycllp_b200.problems import random_equality_arrays

pytestmark = pytest.mark.gpu


def _solve(engine, mode, A, b, c, **kw):
    engine.set_small_kernels(mode)
    try:
        engine.setup_dense(A, b.shape[0])
        info = engine.info()
        return engine.solve_host(b, c, **kw), info
    finally:
        engine.set_small_kernels(1)


def test_small_kernel_shape_sweep_vs_oracle(engine, oracle):
    """m = 1 .. 63 across the panel (8) and warp (32) boundaries of the kernel; with and without
    dense columns (n0 = 0: every column a slack), sparse and dense A0."""
    rng = np.random.RandomState(21)
    for m in (1, 2, 3, 7, 8, 9, 15, 16, 17, 24, 31, 32, 33, 40, 47, 48, 49, 56, 62, 63):
        for n0, dens in ((max(1, m // 2), 1.0), (m + 3, 0.3)):
            N = 5
            A = np.c_[rng.rand(m, n0) * (rng.rand(m, n0) < dens), np.eye(m)]
            b = 0.5 + rng.rand(N, m)
            c = np.c_[0.5 + rng.rand(N, n0), np.zeros((N, m))]
            ref = oracle.solve_dense(A, b, c)
            res, info = _solve(engine, 3, A, b, c)
            # (the largest shapes need > 1/3 of an SM's shared memory and stay on the 512-thread kernel)
            assert info["block"] == 128 or m > 50, (m, info)
            assert_parity(res, ref, c, "small kernel m=%d n0=%d" % (m, n0))
            assert np.abs(res["iters"] - ref.iters).max() <= 1, (m, n0, res["iters"], ref.iters)


def test_small_kernels_agree_with_the_big_one_on_a_large_batch(engine, oracle):
    """Config-1 shape, 2048 LPs (more blocks than fit at once: the work counter cycles): the kernel
    choices (mode 4 = the 128-thread kernel with every panel on its sequential fallback path) give
    the same statuses and step counts and the same solutions to 1e-9."""
    A, b, c = random_equality_arrays(50, 50, 0.1, 2048)
    out = {}
    for mode in (0, 1, 2, 4):
        out[mode], info = _solve(engine, mode, A, b, c)
        assert info["block"] == (128 if mode in (1, 4) else 512)
        assert info["grid"] > 148 if mode else info["grid"] == 148
    ref = oracle.solve_dense(A, b[:64], c[:64])
    for mode in (0, 1, 2, 4):
        assert_parity({k: v[:64] for k, v in out[mode].items()}, ref, c[:64], "mode %d" % mode)
        np.testing.assert_array_equal(out[mode]["status"], out[0]["status"])
        assert np.abs(out[mode]["iters"] - out[0]["iters"]).max() <= 1
        np.testing.assert_allclose(out[mode]["x"], out[0]["x"], rtol=1e-7, atol=1e-9)


def test_small_kernel_warm_start_and_trace(engine, oracle):
    g = golden("cfg1_dense")
    A, b, c = g["A"], g["b"], g["c"]
    engine.set_small_kernels(3)
    try:
        engine.setup_dense(A, b.shape[0])
        assert engine.info()["block"] == 128
        r1 = engine.solve_host(b, c, trace_iters=40)
        ref = oracle.solve_dense_ex(A, b, c, want_trace=True)
        k = int(r1["iters"][0])
        np.testing.assert_allclose(r1["trace"][0, :k, 2], ref.itrace[0, :k, 2], rtol=1e-6)     # gamma per iteration
        # perturbed data, warm start from the previous end point pulled back inside: fewer steps
        rng = np.random.RandomState(3)
        b2, c2 = b * (1 + 0.01 * rng.randn(*b.shape)), c * (1 + 0.01 * rng.randn(*c.shape))
        engine.set_params(warm_floor=1e-2)
        cold = engine.solve_host(b2, c2)
        engine.solve_host(b, c)
        warm = engine.solve_host(b2, c2, warm_start=True)
        assert (warm["status"] == 0).all() and (cold["status"] == 0).all()
        assert warm["iters"].mean() < cold["iters"].mean()
        obj = lambda r: np.einsum("ij,ij->i", r["x"], c2)
        np.testing.assert_allclose(obj(warm), obj(cold), rtol=1e-6)
    finally:
        engine.set_small_kernels(1)
