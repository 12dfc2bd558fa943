"""Ad-hoc: mismatch counts of the infeasible-status fixture for several carry_v settings."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import make_golden_infeasible as gen
from pycllp_b200._cabi import Engine
eng = Engine(0)
g = np.load(os.path.join(ROOT, "tests", "golden", "infeasible_status.npz"))
for cv in (0, 64):
    bad_s = bad_c = 0
    for kind in ("primal", "dual"):
        for m in g["sizes"]:
            A, b, c = gen.instances(kind, int(m))
            eng.setup_dense(A, b.shape[0]); eng.set_params(carry_v=cv)
            res = eng.solve_host(b, c)
            ref = g["%s_%d_status" % (kind, m)]; allowed = g["%s_%d_allowed" % (kind, m)]
            stable = allowed.sum(axis=1) == 1
            for q in range(b.shape[0]):
                st = int(res["status"][q])
                if stable[q] and st != ref[q]: bad_s += 1; print("  cv", cv, kind, m, q, "got", st, "want", ref[q])
                if not stable[q] and not allowed[q, st]: bad_c += 1; print("  cv", cv, kind, m, q, "got", st, "allowed", np.flatnonzero(allowed[q]))
    print("carry_v", cv, "stable mismatches", bad_s, "chaotic outside the allowed set", bad_c)
