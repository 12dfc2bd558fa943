"""One-off: status parity on infeasible / unbounded-looking instances (chaotic IPM trajectories)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pycllp_b200 import _cabi
if os.environ.get('PB200_LIB'): _cabi.LIB_PATH = os.environ['PB200_LIB']
from pycllp_b200._cabi import Engine
from oracle.bindings import Oracle
o = Oracle(); eng = Engine(0)
nst = nit = tot = 0
for m in (3, 5, 8, 11, 13, 16, 21, 24, 33, 40, 50, 64, 80, 100):
    rng = np.random.RandomState(900 + m)
    n0, N = m // 2 + 1, 8
    A = np.c_[rng.rand(m, n0), np.eye(m)]
    b = 0.5 + rng.rand(N, m); c = np.c_[0.5 + rng.rand(N, n0), np.zeros((N, m))]
    for q in range(N):
        k = rng.randint(1, max(2, m // 3 + 1))
        b[q, rng.choice(m, k, replace=False)] = -rng.rand(k) - 0.1     # primal infeasible rows
    ref = o.solve_dense(A, b, c); eng.setup_dense(A, N); res = eng.solve_host(b, c)
    tot += N
    nst += int((res["status"] != ref.status).sum())
    nit += int((np.abs(res["iters"] - ref.iters) > 1).sum())
    print(m, "status", res["status"], ref.status, "iters", res["iters"], ref.iters, flush=True)
print("infeasible instances %d: status mismatches %d, iteration-count mismatches %d" % (tot, nst, nit))
