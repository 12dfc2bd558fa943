"""Ad-hoc: the engine's iteration trace against the oracle's on chosen infeasible instances."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import make_golden_infeasible as gen
from oracle.bindings import Oracle
from pycllp_b200._cabi import Engine
o = Oracle(); eng = Engine(0)
cases = [('primal', 3, 1), ('primal', 3, 2), ('primal', 5, 2), ('primal', 8, 5)]
for kind, m, q in cases:
    A, b, c = gen.instances(kind, m)
    ref = o.solve_dense_ex(A, b[q:q+1], c[q:q+1], want_trace=True)
    print("==", kind, m, q, "oracle status", ref.status[0], "iters", ref.iters[0])
    for cv in (0, 64):
        eng.setup_dense(A, 1)
        eng.set_params(carry_v=cv)
        r = eng.solve_host(b[q:q+1], c[q:q+1], trace_iters=200)
        print("  carry_v", cv, "status", r["status"][0], "iters", r["iters"][0])
        k = int(min(ref.iters[0], r["iters"][0])) + 1
        t, w = r["trace"][0], ref.itrace[0]
        rel = np.abs(t[:k] - w[:k]) / np.maximum(np.abs(w[:k]), 1e-300)
        first = np.argmax(rel.max(axis=1) > 1e-3) if (rel.max(axis=1) > 1e-3).any() else -1
        print("   first iteration with >1e-3 relative deviation:", first)
        lo = max(0, first - 2) if first >= 0 else max(0, k - 4)
        for it in range(lo, min(k, lo + 6)):
            print("    it %3d gpu %s  oracle %s" % (it, np.array2string(t[it], precision=6), np.array2string(w[it], precision=6)))
