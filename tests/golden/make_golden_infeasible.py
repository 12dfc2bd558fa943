"""Golden fixture for the termination status on INFEASIBLE / UNBOUNDED instances.

    python tests/golden/make_golden_infeasible.py

The reference calls statuses 2 and 4 "unreliable" itself (primal_normal.cl:263,268): on an
infeasible LP the iterates diverge, every rounding difference is amplified, and whether
|rho| > 10 |rho_0| (status 2), |sigma| > 10 |sigma_0| (status 4) or the 200-iteration limit
(status 5) fires first can depend on the last bit of the input.  This script makes that
precise instead of hand-waving it: every instance is solved by the oracle (bit-identical to
the reference's kernels compiled as C, tests/test_oracle.py) on the exact input AND on 3 x 128
copies whose A, b, c are perturbed CONSISTENTLY by 1e-15, 1e-13 and 1e-11 relative (SURVEY.md
probe B.3: such perturbations leave every feasible instance's status unchanged and move its
solution by < 1e-11, which is also the size of the engine's own deviation from the reference
on feasible instances).  An instance whose 385 statuses agree is STABLE: the CUDA engine must reproduce
that status.  For the others the set of statuses seen is stored and the engine's status must
lie in it -- the reference itself cannot tell which one is "right".

Instance families (m = 3 .. 100, dense A = [A0 I]):
  primal infeasible : some b_i < 0 with A >= 0            (the set of tests/tools/infeas_check.py)
  dual infeasible   : a column of A0 <= 0 with c_j > 0    (primal unbounded)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.bindings import Oracle  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SIZES = (3, 5, 8, 11, 13, 16, 21, 24, 33, 40, 50, 64, 80, 100)
NPER = 8
NPERT = 128
AMPLITUDES = (1e-15, 1e-13, 1e-11)


def instances(kind, m):
    """(A, b, c) of one shape; the generator is part of the fixture (seeds below)."""
    rng = np.random.RandomState((900 if kind == "primal" else 1900) + m)
    n0 = m // 2 + 1
    A = np.c_[rng.rand(m, n0), np.eye(m)]
    b = 0.5 + rng.rand(NPER, m)
    c = np.c_[0.5 + rng.rand(NPER, n0), np.zeros((NPER, m))]
    if kind == "primal":
        for q in range(NPER):
            k = rng.randint(1, max(2, m // 3 + 1))
            b[q, rng.choice(m, k, replace=False)] = -rng.rand(k) - 0.1     # infeasible rows
    else:
        A[:, 0] = -rng.rand(m) - 0.1          # x_0 can grow for ever, c_0 > 0: unbounded
    return A, b, c


if __name__ == "__main__":
    o = Oracle()
    out = {}
    hist = np.zeros(6, dtype=int)
    nstable = ntotal = 0
    for kind in ("primal", "dual"):
        for m in SIZES:
            A, b, c = instances(kind, m)
            base = o.solve_dense(A, b, c)
            seen = np.zeros((NPER, 6), dtype=bool)
            seen[np.arange(NPER), base.status] = True
            for amp in AMPLITUDES:
                for k in range(NPERT):
                    rng = np.random.RandomState(77 + k)
                    pert = lambda a: a * (1.0 + amp * rng.standard_normal(a.shape))
                    r = o.solve_dense(pert(A), pert(b), pert(c))
                    seen[np.arange(NPER), r.status] = True
            stable = seen.sum(axis=1) == 1
            out["%s_%d_status" % (kind, m)] = base.status
            out["%s_%d_iters" % (kind, m)] = base.iters
            out["%s_%d_allowed" % (kind, m)] = seen
            hist += np.bincount(base.status, minlength=6)
            nstable += int(stable.sum())
            ntotal += NPER
            print(kind, m, "status", base.status, "stable", stable.astype(int), flush=True)
    print("instances %d, stable under 1e-15 perturbations %d, status histogram %s" % (ntotal, nstable, hist))
    np.savez_compressed(os.path.join(OUT, "infeasible_status.npz"), sizes=np.array(SIZES), nper=NPER, **out)
