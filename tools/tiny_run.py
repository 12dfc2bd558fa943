"""Ad-hoc: a tiny dense batch (m, n0, N from argv) through the engine; prints status/iterations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pycllp_b200 import _cabi
if os.environ.get('PB200_LIB'): _cabi.LIB_PATH = os.environ['PB200_LIB']
from pycllp_b200._cabi import Engine
m, n0, N = (int(a) for a in sys.argv[1:4])
rng = np.random.RandomState(1)
A = np.c_[rng.rand(m, n0), np.eye(m)]
b = 0.5 + rng.rand(N, m); c = np.c_[0.5 + rng.rand(N, n0), np.zeros((N, m))]
eng = Engine(0); eng.setup_dense(A, N)
if len(sys.argv) > 4: eng.set_params(max_iter=int(sys.argv[4]))
res = eng.solve_host(b, c)
print("status", res["status"], "iters", res["iters"], flush=True)
