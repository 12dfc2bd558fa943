"""One cfg5 launch (m=500, n=1000) for ncu: python tools/prof_run5.py [N]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycllp_b200._cabi import Engine
from pycllp_b200.problems import random_equality_arrays
N = int(sys.argv[1]) if len(sys.argv) > 1 else 148
A, b, c = random_equality_arrays(500, 500, 1.0, N)
eng = Engine(0)
eng.setup_dense(A, N)
res = eng.solve_host(b, c)
print("status0", int((res["status"] == 0).sum()), "steps", res["iters"].mean())
