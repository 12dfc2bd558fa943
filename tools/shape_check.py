"""Ad-hoc: engine vs oracle on a list of m (dense, n0 = 60)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pycllp_b200._cabi import Engine
from oracle.bindings import Oracle
o = Oracle(); eng = Engine(0)
for m in [int(a) for a in sys.argv[1:]]:
    rng = np.random.RandomState(5 + m)
    n0, N = 60, 3
    A = np.c_[rng.rand(m, n0), np.eye(m)]
    b = 0.5 + rng.rand(N, m); c = np.c_[0.5 + rng.rand(N, n0), np.zeros((N, m))]
    ref = o.solve_dense(A, b, c)
    eng.setup_dense(A, N); res = eng.solve_host(b, c)
    print(m, "status", res["status"], ref.status, "iters", res["iters"], ref.iters, "dx %.2e" % np.abs(res["x"] - ref.x).max(), flush=True)
