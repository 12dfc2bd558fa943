"""Generate the golden fixtures under tests/golden/ from the REFERENCE ITSELF.

Run in the build container (needs /root/reference, through oracle/_ref):

    python tests/golden/make_golden.py

Every fixture stores the inputs (A, b, c -- scipy.sparse.rand's stream is version
dependent) and the outputs of the reference's OpenCL kernels compiled
as C (oracle/_ref/libpycllp_ref.so, built by oracle/Makefile from
/root/reference/pycllp/cl/*.cl): x, y, z, status, number of Newton steps, and the last
(|rho|, |sigma|, gamma).  The kernel-level fixtures come from the reference's
`solve_primal_normal` / `sparse_solve_primal_normal` / `ldl` / `modified_ldl` kernels on
the inputs of tests/test_ldl.py (seed 123456).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.bindings import Reference, sparse_structures  # noqa: E402
from pycllp_b200.problems import random_equality_arrays, sparse_equality_arrays  # noqa: E402
import problems  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ref = Reference()


def save(name, **kw):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **kw)
    print("wrote", path, os.path.getsize(path), "bytes")


def solve_fixture(name, A, b, c, sparse=False):
    r = ref.solve_sparse(A, b, c) if sparse else ref.solve_dense(A, b, c)
    print(name, "status", np.bincount(r.status, minlength=6), "steps", r.iters.min(), r.iters.max())
    save(name, A=A, b=b, c=c, x=r.x, y=r.y, z=r.z, status=r.status, iters=r.iters, trace=r.trace)


# config 1: examples/random_problem.py, N=50, 64 problems (m=50, n=100)
A, b, c = random_equality_arrays(50, 50, 0.1, 64)
solve_fixture("cfg1_dense", A, b, c)
solve_fixture("cfg1_sparse", A, b, c, sparse=True)

# config 2: Vanderbei 2.9 / 2.10 and the small problem, unperturbed + 32 perturbations
for nm, fn in (("vanderbei_2_9", problems.vanderbei_2_9), ("vanderbei_2_10", problems.vanderbei_2_10)):
    lp, xopt = fn()
    A, b0, c0 = problems.equality_arrays(lp)
    bb, cc = problems.perturb(b0[0], c0[0], 32)
    b = np.vstack([b0, bb])
    c = np.vstack([c0, cc])
    solve_fixture("cfg2_" + nm + "_dense", A, b, c)
    solve_fixture("cfg2_" + nm + "_sparse", A, b, c, sparse=True)
Asm, bsm, csm, _ = problems.small_problem()
lp = problems.StandardLP(Asm, bsm, csm, 0.0)
A, b0, c0 = problems.equality_arrays(lp)
bb, cc = problems.perturb(bsm, csm, 32)
lp32 = problems.StandardLP(Asm, bb, cc, 0.0)
_, b32, c32 = problems.equality_arrays(lp32)
solve_fixture("cfg2_small_dense", A, np.vstack([b0, b32]), np.vstack([c0, c32]))
solve_fixture("cfg2_small_sparse", A, np.vstack([b0, b32]), np.vstack([c0, c32]), sparse=True)

# config 3 sample: m=200, n=400, density 1.0, first 8 problems of the 4096
A, b, c = random_equality_arrays(200, 200, 1.0, 4096)
solve_fixture("cfg3_sample", A, b[:8], c[:8])

# config 5 sample: m=500, n=1000, first 2 problems (batch truncated: same A, same first rows)
A, b, c = random_equality_arrays(500, 500, 1.0, 2)
solve_fixture("cfg5_sample", A, b, c)

# sparse mid-size (config 4 shape scaled down: m=300, n0=450, 3 % density, identity slacks)
A, b, c = sparse_equality_arrays(300, 450, 0.03, 8, seed=1)
solve_fixture("cfg4_small_sparse", A.toarray(), b, c, sparse=True)

# kernel-level: tests/test_ldl.py:219-273 (dense) and :276-361 (sparse), seed 123456
m, n, N = 100, 80, 32
np.random.seed(123456)
A = np.c_[np.random.rand(m, n), np.eye(m)]
x = np.random.rand(m + n, N); z = np.random.rand(m + n, N); y = np.random.rand(m, N)
b = np.random.rand(m, N); c = np.r_[np.random.rand(n, N), np.zeros((m, N))]
dy, L, D = ref.solve_primal_normal(A, x.T, z.T, y.T, b.T, c.T, 1.0)
save("kernel_solve_primal_normal", A=A, x=x.T, z=z.T, y=y.T, b=b.T, c=c.T, mu=1.0, dy=dy, D=D)

from scipy.sparse import rand as sprand  # noqa: E402
np.random.seed(123456)
A = np.c_[sprand(m, n, density=0.025).toarray(), np.eye(m)]
x = np.random.rand(m + n, N); z = np.random.rand(m + n, N); y = np.random.rand(m, N)
b = np.random.rand(m, N); c = np.r_[np.random.rand(n, N), np.zeros((m, N))]
dy = ref.sparse_solve_primal_normal(A, x.T, z.T, y.T, b.T, c.T, 1.0)
save("kernel_sparse_solve_primal_normal", A=A, x=x.T, z=z.T, y=y.T, b=b.T, c=c.T, mu=1.0, dy=dy)

# kernel-level: tests/test_ldl.py:139-193 (ldl / modified_ldl on random SPD matrices)
np.random.seed(7)
AA = np.empty((8, 40, 40))
for i in range(8):
    B = sprand(40, 32, density=0.1).toarray()
    AA[i] = B.dot(B.T) + np.eye(40) * 40
L1, D1 = ref.ldl(AA, modified=False)
beta = float(np.sqrt(np.amax(AA)))
L2, D2 = ref.ldl(AA, modified=True, beta=beta, delta=1e-6)
save("kernel_ldl", AA=AA, L_plain=L1, D_plain=D1, beta=beta, L_mod=L2, D_mod=D2)

# Oracle A (the reference's CPU solver, normal_eqns.py + _ldl.pyx): first 4 LPs of config 1.
# Its statuses are rounding-noise driven: stored for the objective cross-check of
# the 'py' preset only.
from oracle import oracle_a  # noqa: E402
if oracle_a.available():
    A, b, c = random_equality_arrays(50, 50, 0.1, 64)
    r = oracle_a.solve(A, b[:4], c[:4])
    save("cfg1_oracle_a", A=A, b=b[:4], c=c[:4], **r)
