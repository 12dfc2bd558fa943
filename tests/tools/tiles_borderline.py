"""Which LPs of the 6000-row staircase batch do not end with status 0, and their step counts."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pycllp_b200._cabi import Engine
from pycllp_b200.problems import staircase_equality_arrays
A, b, c = staircase_equality_arrays(6000, 9000, 48, 4, 296, seed=0)
eng = Engine(0)
eng.setup_sparse(A, 296)
res = eng.solve_host(b, c, trace_iters=200)
bad = np.nonzero(res["status"] != 0)[0]
print("non-zero status:", bad, res["status"][bad], res["iters"][bad])
print("iters histogram:", np.bincount(res["iters"])[30:])
for q in bad:
    tr = res["trace"][q]
    k = int(res["iters"][q])
    print(q, "last iterations |rho| |sigma| gamma:")
    for it in range(max(0, k - 6), min(k + 1, 200)):
        print("   ", it, tr[it])
np.savez("gpurun_out/tiles_borderline.npz", bad=bad, status=res["status"], iters=res["iters"])
