"""How often the speculative factorisation is redone by the sequential rule (phase slots 7 / 11 count it)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycllp_b200._cabi import Engine
from pycllp_b200.problems import random_equality_arrays, sparse_equality_arrays
eng = Engine(0)
for wl, N in (("cfg3", 1184), ("cfg5", 296), ("cfg4", 148)):
    if wl == "cfg4":
        A, b, c = sparse_equality_arrays(2000, 3000, 0.01, N, seed=0)
        eng.setup_sparse(A, N)
    else:
        m = 200 if wl == "cfg3" else 500
        A, b, c = random_equality_arrays(m, m, 1.0, N)
        eng.setup_dense(A, N)
    eng.phase_profile(True)
    res = eng.solve_host(b, c)
    prof = eng.phase_profile(False)
    steps = int(res["iters"].sum())
    print(wl, "N", N, "steps", steps, "redo(big)", prof["f_waitEd"] if wl != "cfg3" else "-", "redo(ahead)",
          prof["f_table"] if wl == "cfg3" else "-", "factor cyc/step", prof["factor"] // steps, flush=True)
