"""Ad-hoc GPU check: engine vs oracle on small configs + a first timing. Not a test."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pycllp_b200 import _cabi
if os.environ.get("PB200_LIB"): _cabi.LIB_PATH = os.environ["PB200_LIB"]
from pycllp_b200._cabi import Engine
from pycllp_b200.problems import random_equality_arrays, sparse_equality_arrays
from oracle.bindings import Oracle

def compare(tag, res, ref):
    st_ok = np.array_equal(res["status"], ref.status)
    obj = lambda r, c: np.einsum("ij,ij->i", r["x"], c)
    print(tag, "status equal:", st_ok, "iters diff max:", np.abs(res["iters"] - ref.iters).max(),
          "| dx %.2e dy %.2e dz %.2e" % tuple(np.abs(res[k] - ref[k]).max() for k in "xyz"))
    return st_ok

o = Oracle()
eng = Engine(0)
which = sys.argv[1:] or ["cfg1", "cfg3s", "sparse", "time3"]
if "cfg1" in which:
    A, b, c = random_equality_arrays(50, 50, 0.1, 64)
    eng.setup_dense(A, 64)
    print(eng.info())
    res = eng.solve_host(b, c)
    ref = o.solve_dense(A, b, c)
    compare("cfg1", res, ref)
    print(" gpu iters", res["iters"][:16], "status", res["status"][:16])
    rel = np.abs(np.einsum("ij,ij->i", res["x"], c) - np.einsum("ij,ij->i", ref.x, c)) / np.abs(np.einsum("ij,ij->i", ref.x, c))
    print(" rel obj err max %.2e" % rel.max())
if "cfg3s" in which:
    A, b, c = random_equality_arrays(200, 200, 1.0, 16)
    eng.setup_dense(A, 16)
    print(eng.info())
    res = eng.solve_host(b, c)
    ref = o.solve_dense(A, b, c)
    compare("cfg3 sample", res, ref)
    print(" gpu iters", res["iters"], "oracle iters", ref.iters)
if "sparse" in which:
    A, b, c = sparse_equality_arrays(100, 150, 0.05, 16)
    eng.setup_sparse(A, 16)
    print(eng.info())
    res = eng.solve_host(b, c)
    ref = o.solve_sparse(A.toarray(), b, c)
    compare("sparse", res, ref)
    print(" gpu iters", res["iters"], "oracle iters", ref.iters)
if "time3" in which:
    N = 4096
    A, b, c = random_equality_arrays(200, 200, 1.0, N)
    eng.setup_dense(A, N)
    print(eng.info())
    for rep in range(3):
        t = time.time(); res = eng.solve_host(b, c); dt = time.time() - t
        print("cfg3 N=%d: %.3f s  -> %.0f solves/s; status hist %s; iters mean %.1f" % (
            N, dt, N / dt, np.bincount(res["status"], minlength=6), res["iters"].mean()))
if "time5" in which:
    N = 592
    A, b, c = random_equality_arrays(500, 500, 1.0, N)
    eng.setup_dense(A, N)
    print(eng.info())
    for rep in range(2):
        t = time.time(); res = eng.solve_host(b, c); dt = time.time() - t
        print("cfg5 N=%d: %.3f s  -> %.0f solves/s; status hist %s; iters mean %.1f" % (
            N, dt, N / dt, np.bincount(res["status"], minlength=6), res["iters"].mean()))
if "prof3" in which:
    N = 1184
    A, b, c = random_equality_arrays(200, 200, 1.0, N)
    eng.setup_dense(A, N)
    eng.solve_host(b, c)
    eng.phase_profile(True)
    t = time.time(); res = eng.solve_host(b, c); dt = time.time() - t
    prof = eng.phase_profile(False)
    tot = sum(prof.values())
    print("cfg3 N=%d %.3fs; phase cycles per Newton step:" % (N, dt))
    steps = res["iters"].sum()
    for k, v in prof.items():
        print("  %-10s %9.0f cyc/step  %5.1f%%" % (k, v / steps, 100.0 * v / tot))
    print("  total %.0f cyc/step" % (tot / steps))
if "time4" in which:
    N = int(os.environ.get("N4", "148"))
    A, b, c = sparse_equality_arrays(2000, 3000, 0.01, N, seed=0)
    t = time.time(); eng.setup_sparse(A, N); print("cfg4 setup %.2fs" % (time.time() - t), eng.info())
    t = time.time(); res = eng.solve_host(b, c); dt = time.time() - t
    print("cfg4 N=%d: %.3f s  -> %.1f solves/s; status hist %s; iters mean %.1f" % (
        N, dt, N / dt, np.bincount(res["status"], minlength=6), res["iters"].mean()))
    Ad = A.toarray()
    print(" max |b-Ax| %.2e  max |c-A'y+z| %.2e  max gap %.2e" % (
        np.linalg.norm(b - res["x"] @ Ad.T, axis=1).max(), np.linalg.norm(c - res["y"] @ Ad + res["z"], axis=1).max(),
        np.einsum("ij,ij->i", res["x"], res["z"]).max()))
if "prof5" in which:
    N = 296
    A, b, c = random_equality_arrays(500, 500, 1.0, N)
    eng.setup_dense(A, N)
    print(eng.info())
    eng.solve_host(b[:148], c[:148])
    eng.phase_profile(True)
    t = time.time(); res = eng.solve_host(b, c); dt = time.time() - t
    prof = eng.phase_profile(False)
    steps = res["iters"].sum()
    tot = sum(v for k, v in prof.items() if k in ("rhs_norms", "form_M", "factor", "tri_solve", "residual", "step"))
    print("cfg5 N=%d %.3fs (%.0f solves/s); phase cycles per Newton step:" % (N, dt, N / dt))
    for k, v in prof.items():
        print("  %-14s %10.0f cyc/step  %5.1f%%" % (k, v / steps, 100.0 * v / tot))

if "prof4" in which:
    N = int(os.environ.get("N4", "148"))
    A, b, c = sparse_equality_arrays(2000, 3000, 0.01, N, seed=0)
    eng.setup_sparse(A, N)
    eng.phase_profile(True)
    t = time.time(); res = eng.solve_host(b, c); dt = time.time() - t
    prof = eng.phase_profile(False)
    steps = res["iters"].sum()
    tot = sum(v for k, v in prof.items() if k in ("rhs_norms", "form_M", "factor", "tri_solve", "residual", "step", "f_copy"))
    print("cfg4 N=%d %.3fs (%.1f solves/s); phase cycles per Newton step:" % (N, dt, N / dt))
    for k, v in prof.items():
        print("  %-14s %10.0f cyc/step  %5.1f%%" % (k, v / steps, 100.0 * v / tot))

if "proft" in which:
    # genuinely sparse LPs (staircase structure): tile-sparse factor vs the dense packed kernels
    from pycllp_b200.problems import staircase_equality_arrays
    for (m, n0, band, N, modes) in [(1500, 2250, 40, 296, ("tiles", "dense")), (6000, 9000, 48, 296, ("tiles",))]:
        A, b, c = staircase_equality_arrays(m, n0, band, 4, N, seed=0)
        for mode in modes:
            t = time.time(); eng.setup_sparse(A, N, factor=mode); ts = time.time() - t
            si = eng.sparse_info()
            eng.solve_host(b[:148], c[:148])
            eng.phase_profile(True)
            t = time.time(); res = eng.solve_host(b, c); dt = time.time() - t
            prof = eng.phase_profile(False)
            steps = res["iters"].sum()
            tot = sum(v for k, v in prof.items() if k in ("rhs_norms", "form_M", "factor", "tri_solve", "residual", "step"))
            print("staircase m=%d n=%d %s: setup %.2fs, N=%d %.3fs (%.1f solves/s), status0 %d, steps %.1f, factor doubles %d (dense %d), pairs %d" % (
                m, m + n0, mode, ts, N, dt, N / dt, int((res["status"] == 0).sum()), res["iters"].mean(),
                si["factor_doubles"], si["dense_factor_doubles"], si["update_pairs"]))
            for k, v in prof.items():
                if v:
                    print("  %-14s %10.0f cyc/step  %5.1f%%" % (k, v / steps, 100.0 * v / tot))

if "time1" in which:
    # config-1 shape in a large batch: two blocks per SM (PB200_SMALL=0 switches that off)
    for (m, n0, dens) in [(50, 50, 0.1), (100, 100, 1.0), (24, 40, 0.5)]:
        N = 4096
        A, b, c = random_equality_arrays(m, n0, dens, N)
        eng.setup_dense(A, N)
        print(eng.info())
        ref = o.solve_dense(A, b[:32], c[:32])
        for rep in range(3):
            t = time.time(); res = eng.solve_host(b, c); dt = time.time() - t
            print("m=%d n=%d N=%d: %.4f s -> %.0f solves/s; status0 %d; iters mean %.1f" % (
                m, m + n0, N, dt, N / dt, int((res["status"] == 0).sum()), res["iters"].mean()))
        compare(" vs oracle (first 32)", {k: v[:32] for k, v in res.items()}, ref)

if "prof1" in which:
    N = 4096
    A, b, c = random_equality_arrays(50, 50, 0.1, N)
    eng.setup_dense(A, N)
    print(eng.info())
    eng.solve_host(b, c)
    eng.phase_profile(True)
    t = time.time(); res = eng.solve_host(b, c); dt = time.time() - t
    prof = eng.phase_profile(False)
    steps = res["iters"].sum()
    print("cfg1 shape N=%d %.4fs (%.0f solves/s); phase cycles per Newton step (per block):" % (N, dt, N / dt))
    for k, v in prof.items():
        if v:
            print("  %-14s %10.0f cyc/step" % (k, v / steps))

if "prof1s" in which:
    # one LP per block, alone on its SM: latency of each phase of the 128-thread kernel
    N = 64
    A, b, c = random_equality_arrays(50, 50, 0.1, N)
    eng.setup_dense(A, N)
    print(eng.info())
    eng.solve_host(b, c)
    eng.phase_profile(True)
    t = time.time(); res = eng.solve_host(b, c); dt = time.time() - t
    prof = eng.phase_profile(False)
    steps = res["iters"].sum()
    print("cfg1 N=%d %.4fs; phase cycles per Newton step (one block per SM):" % (N, dt))
    for k, v in prof.items():
        if v:
            print("  %-14s %10.0f cyc/step" % (k, v / steps))
