"""One launch of a bench workload for ncu:  python tools/prof_run.py [cfg3|cfg5|cfg4] [N] [max_iter]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycllp_b200._cabi import Engine
from pycllp_b200.problems import random_equality_arrays, sparse_equality_arrays
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
N = int(sys.argv[2]) if len(sys.argv) > 2 else {"cfg3": 1184, "cfg5": 296, "cfg4": 148}[wl]
eng = Engine(0)
if wl == "cfg4":
    A, b, c = sparse_equality_arrays(2000, 3000, 0.01, N, seed=0)
    eng.setup_sparse(A, N)
else:
    m = 200 if wl == "cfg3" else 500
    A, b, c = random_equality_arrays(m, m, 1.0, N)
    eng.setup_dense(A, N)
if len(sys.argv) > 3:
    eng.set_params(max_iter=int(sys.argv[3]))
res = eng.solve_host(b, c)
print(wl, "N", N, "status0", int((res["status"] == 0).sum()), "steps", res["iters"].mean())
