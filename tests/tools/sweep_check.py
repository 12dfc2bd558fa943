"""One-off robustness sweep: engine (dense and sparse paths) vs the oracle on many small shapes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import scipy.sparse as sp
from pycllp_b200 import _cabi
if os.environ.get('PB200_LIB'): _cabi.LIB_PATH = os.environ['PB200_LIB']
from pycllp_b200._cabi import Engine
from oracle.bindings import Oracle
o = Oracle(); eng = Engine(0)
bad = 0
t0 = time.time()
shapes = [] if os.environ.get('HARD_ONLY') else [(m, max(1, m // 2 + (m % 3)), 2) for m in range(1, 65)] + ([] if os.environ.get('HARD_ONLY') else [(m, m // 3 + 1, 2) for m in (70, 88, 100, 111, 128, 136, 150, 168, 185, 199, 208)])
for m, n0, N in shapes:
    rng = np.random.RandomState(100 + m)
    dens = 1.0 if m % 2 else 0.5
    A0 = rng.rand(m, n0) * (rng.rand(m, n0) < dens)
    A = np.c_[A0, np.eye(m)]
    b = 0.5 + rng.rand(N, m); c = np.c_[0.5 + rng.rand(N, n0), np.zeros((N, m))]
    for path in ("dense", "sparse"):
        if path == "dense":
            ref = o.solve_dense(A, b, c); eng.setup_dense(A, N)
        else:
            ref = o.solve_sparse(A, b, c); eng.setup_sparse(sp.csr_matrix(A), N)
        res = eng.solve_host(b, c)
        ok = np.array_equal(res["status"], ref.status)
        good = ref.status == 0
        dx = np.abs(res["x"][good] - ref.x[good]).max() if good.any() else 0.0
        di = np.abs(res["iters"] - ref.iters).max()
        if not ok or dx > 1e-6 or di > 1:
            bad += 1
            print("MISMATCH m=%d n0=%d %s status %s vs %s iters %s vs %s dx %.2e" % (m, n0, path, res["status"], ref.status, res["iters"], ref.iters, dx), flush=True)
print("sweep done: %d shapes x 2 paths, %d mismatches, %.0f s" % (len(shapes), bad, time.time() - t0))
# infeasible / badly scaled instances: the statuses must still agree with the reference
bad = 0
for m in (3, 8, 13, 24, 40, 64, 100, 200):
    rng = np.random.RandomState(700 + m)
    n0, N = m // 2 + 1, 4
    A = np.c_[rng.rand(m, n0), np.eye(m)]
    b = 0.5 + rng.rand(N, m); c = np.c_[0.5 + rng.rand(N, n0), np.zeros((N, m))]
    b[0, : max(1, m // 4)] = -1.0                      # primal infeasible
    c[1] *= 1e4                                        # badly scaled objective
    b[2] *= 1e-3                                       # tiny right-hand side
    A2 = A.copy(); A2[:, n0 - 1] = A2[:, 0]            # duplicated column
    for tag, AA in (("plain", A), ("dupcol", A2)):
        for path in ("dense", "sparse"):
            if path == "dense":
                ref = o.solve_dense(AA, b, c); eng.setup_dense(AA, N)
            else:
                ref = o.solve_sparse(AA, b, c); eng.setup_sparse(sp.csr_matrix(AA), N)
            res = eng.solve_host(b, c)
            ok = np.array_equal(res["status"], ref.status)
            di = np.abs(res["iters"] - ref.iters).max()
            if not ok or di > 1:
                bad += 1
                print("MISMATCH m=%d %s %s status %s vs %s iters %s vs %s" % (m, tag, path, res["status"], ref.status, res["iters"], ref.iters), flush=True)
print("hard cases done, %d mismatches" % bad)
