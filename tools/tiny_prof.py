import sys, os
sys.path.insert(0, "/root/repo")
from pycllp_b200._cabi import Engine
from pycllp_b200.problems import random_equality_arrays
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
A, b, c = random_equality_arrays(50, 50, 0.1, N)
eng = Engine(0)
eng.set_small_kernels(3)
eng.setup_dense(A, N)
res = eng.solve_host(b, c)
print(N, (res["status"] == 0).sum(), res["iters"].mean())
