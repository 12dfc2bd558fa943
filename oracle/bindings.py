"""ctypes bindings of the CHECKERS under oracle/ -- test infrastructure only.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline /
``--impl reference`` legs) may import this module; nothing under ``pycllp_b200/``
does.

* ``liboracle.so``            -- oracle/ipm_oracle.c, the C restatement.
* ``_ref/libpycllp_ref.so``   -- the reference's own .cl kernels compiled as C
                                 (oracle/ref_shim.c); present when it was built in
                                 the container that has /root/reference.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """Compile the checkers (make -C oracle). Building the checker is not using it."""
    if force or not os.path.exists(os.path.join(HERE, "liboracle.so")) or \
            os.path.getmtime(os.path.join(HERE, "liboracle.so")) < os.path.getmtime(os.path.join(HERE, "ipm_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    ref_so = os.path.join(HERE, "_ref", "libpycllp_ref.so")
    have_ref_src = os.path.exists("/root/reference/pycllp/cl/primal_normal.cl")
    if have_ref_src and (force or not os.path.exists(ref_so) or
                         os.path.getmtime(ref_so) < os.path.getmtime(os.path.join(HERE, "ref_shim.c"))):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    import glob
    if os.path.exists("/root/reference/pycllp/_ldl.pyx") and (force or not glob.glob(os.path.join(HERE, "_ref", "_ldl*.so"))):
        subprocess.call(["make", "-C", HERE, "ref_py"], stdout=subprocess.DEVNULL)   # Oracle A (needs cython)


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Params(ctypes.Structure):
    _fields_ = [("eps", ctypes.c_double), ("delta", ctypes.c_double), ("r", ctypes.c_double),
                ("ldl_delta", ctypes.c_double), ("refine_tol", ctypes.c_double),
                ("max_iter", ctypes.c_int), ("max_refine", ctypes.c_int)]


def sparse_structures(A_dense):
    """The host-side symbolic analysis of the reference's sparse solver, cl.py:175-196:
    CSR of A and A', pattern of L from one dense Cholesky of A A' (natural order),
    CSR of L' and the LTmap permutation."""
    from scipy.sparse import csr_matrix, tril
    A = np.asarray(A_dense, dtype=np.float64)
    Asp = csr_matrix(A)
    ATsp = Asp.transpose().tocsr()
    L = np.linalg.cholesky(A.dot(A.T))
    L = tril(csr_matrix(L), format="csr")
    LT = L.transpose().tocsr()
    LTmap = np.argsort(L.indices, kind="mergesort").astype(np.int32)
    return dict(
        Adata=_f64(Asp.data), Aindptr=_i32(Asp.indptr), Aindices=_i32(Asp.indices),
        ATdata=_f64(ATsp.data), ATindptr=_i32(ATsp.indptr), ATindices=_i32(ATsp.indices),
        nnzL=int(L.nnz), Lindptr=_i32(L.indptr), Lindices=_i32(L.indices),
        LTindptr=_i32(LT.indptr), LTindices=_i32(LT.indices), LTmap=LTmap)


class _Result(dict):
    __getattr__ = dict.__getitem__


def _alloc(N, m, n):
    return (np.empty((N, n)), np.empty((N, m)), np.empty((N, n)),
            np.empty(N, dtype=np.int32), np.empty(N, dtype=np.int32), np.empty((N, 3)))


class Oracle(object):
    """The C restatement (oracle/ipm_oracle.c)."""

    def __init__(self):
        build()
        self.lib = ctypes.CDLL(os.path.join(HERE, "liboracle.so"))

    def default_params(self, sparse=False):
        p = Params()
        self.lib.oracle_default_params(ctypes.byref(p), int(sparse))
        return p

    def solve_dense(self, A, b, c, params=None, nthreads=None):
        A, b, c = _f64(A), _f64(np.atleast_2d(b)), _f64(np.atleast_2d(c))
        m, n = A.shape
        N = b.shape[0]
        x, y, z, status, iters, trace = _alloc(N, m, n)
        nrefs = np.empty(N, dtype=np.int32)
        self.lib.oracle_solve_dense(N, m, n, _d(A), _d(b), _d(c), _d(x), _d(y), _d(z), _i(status),
                                    _i(iters), _i(nrefs), _d(trace),
                                    ctypes.byref(params) if params is not None else None,
                                    int(nthreads or os.cpu_count() or 1))
        return _Result(x=x, y=y, z=z, status=status, iters=iters, nrefs=nrefs, trace=trace)

    def solve_dense_ex(self, A, b, c, start=None, params=None, nthreads=None, want_trace=False):
        """Dense solve from a given start (x0, y0, z0) -- warm start -- and/or with the
        (|rho|, |sigma|, gamma) of every iteration, shape (N, max_iter, 3), NaN where not reached."""
        A, b, c = _f64(A), _f64(np.atleast_2d(b)), _f64(np.atleast_2d(c))
        m, n = A.shape
        N = b.shape[0]
        p = params if params is not None else self.default_params()
        x, y, z, status, iters, _ = _alloc(N, m, n)
        if start is not None:
            x[:], y[:], z[:] = start
        it = np.full((N, p.max_iter, 3), np.nan) if want_trace else None
        self.lib.oracle_solve_dense_ex(N, m, n, _d(A), _d(b), _d(c), _d(x), _d(y), _d(z), _i(status),
                                       _i(iters), ctypes.byref(p), int(nthreads or os.cpu_count() or 1),
                                       int(start is not None), _d(it))
        return _Result(x=x, y=y, z=z, status=status, iters=iters, itrace=it)

    def solve_sparse(self, A, b, c, params=None, nthreads=None, structures=None):
        A, b, c = _f64(A), _f64(np.atleast_2d(b)), _f64(np.atleast_2d(c))
        m, n = A.shape
        N = b.shape[0]
        s = structures or sparse_structures(A)
        x, y, z, status, iters, trace = _alloc(N, m, n)
        self.lib.oracle_solve_sparse(
            N, m, n, _d(s["Adata"]), _i(s["Aindptr"]), _i(s["Aindices"]), _d(s["ATdata"]),
            _i(s["ATindptr"]), _i(s["ATindices"]), s["nnzL"], _i(s["Lindptr"]), _i(s["Lindices"]),
            _i(s["LTindptr"]), _i(s["LTindices"]), _i(s["LTmap"]), _d(b), _d(c), _d(x), _d(y), _d(z),
            _i(status), _i(iters), _d(trace), ctypes.byref(params) if params is not None else None,
            int(nthreads or os.cpu_count() or 1))
        return _Result(x=x, y=y, z=z, status=status, iters=iters, trace=trace)

    def solve_primal_normal(self, A, x, z, y, b, c, mu, delta=1e-6, max_refine=5, want_factor=False):
        """Per problem: dy of ldl.cl:602-653. x,z,c: (N,n); y,b: (N,m)."""
        A = _f64(A)
        m, n = A.shape
        x, z, y, b, c = (_f64(np.atleast_2d(v)) for v in (x, z, y, b, c))
        N = x.shape[0]
        dy = np.empty((N, m))
        L = np.empty((N, m * (m + 1) // 2)) if want_factor else None
        D = np.empty((N, m)) if want_factor else None
        self.lib.oracle_solve_primal_normal(N, m, n, _d(A), _d(x), _d(z), _d(y), _d(b), _d(c),
                                            ctypes.c_double(mu), ctypes.c_double(delta),
                                            int(max_refine), _d(dy), _d(L), _d(D))
        return (dy, L, D) if want_factor else dy

    def sparse_solve_primal_normal(self, A, x, z, y, b, c, mu, delta=1e-6, structures=None):
        A = _f64(A)
        m, n = A.shape
        s = structures or sparse_structures(A)
        x, z, y, b, c = (_f64(np.atleast_2d(v)) for v in (x, z, y, b, c))
        N = x.shape[0]
        dy = np.empty((N, m))
        self.lib.oracle_sparse_solve_primal_normal(
            N, m, n, _d(s["Adata"]), _i(s["Aindptr"]), _i(s["Aindices"]), _d(s["ATdata"]),
            _i(s["ATindptr"]), _i(s["ATindices"]), s["nnzL"], _i(s["Lindptr"]), _i(s["Lindices"]),
            _i(s["LTindptr"]), _i(s["LTindices"]), _i(s["LTmap"]), _d(x), _d(z), _d(y), _d(b), _d(c),
            ctypes.c_double(mu), ctypes.c_double(delta), _d(dy))
        return dy

    def ldl(self, AA, modified=False, beta=1.0, delta=1e-6):
        """AA: (N, m, m). Returns packed L (N, m(m+1)/2) row-major-packed and D (N, m)."""
        AA = _f64(AA)
        N, m, _ = AA.shape
        L = np.empty((N, m * (m + 1) // 2))
        D = np.empty((N, m))
        self.lib.oracle_ldl(N, m, _d(AA), _d(L), _d(D), int(modified), ctypes.c_double(beta),
                            ctypes.c_double(delta))
        return L, D


class Reference(object):
    """The reference's own kernels compiled as C (oracle/_ref/libpycllp_ref.so)."""

    @staticmethod
    def available():
        return os.path.exists(os.path.join(HERE, "_ref", "libpycllp_ref.so"))

    def __init__(self):
        build()
        self.lib = ctypes.CDLL(os.path.join(HERE, "_ref", "libpycllp_ref.so"))

    @staticmethod
    def _steps(status, last_iter):
        # the kernel prints at the top of every iteration; the last printed index is
        # the iteration whose stop test fired, or MAX_ITER-1 when none did (status 5)
        return np.where(status == 5, last_iter + 1, last_iter).astype(np.int32)

    def solve_dense(self, A, b, c, nthreads=None):
        A, b, c = _f64(A), _f64(np.atleast_2d(b)), _f64(np.atleast_2d(c))
        m, n = A.shape
        N = b.shape[0]
        x, y, z, status, iters, trace = _alloc(N, m, n)
        self.lib.ref_run_dense(N, m, n, _d(A), _d(b), _d(c), _d(x), _d(y), _d(z), _i(status),
                               _i(iters), _d(trace), int(nthreads or os.cpu_count() or 1))
        return _Result(x=x, y=y, z=z, status=status, iters=self._steps(status, iters), trace=trace)

    def solve_sparse(self, A, b, c, nthreads=None, structures=None):
        A, b, c = _f64(A), _f64(np.atleast_2d(b)), _f64(np.atleast_2d(c))
        m, n = A.shape
        N = b.shape[0]
        s = structures or sparse_structures(A)
        x, y, z, status, iters, trace = _alloc(N, m, n)
        self.lib.ref_run_sparse(
            N, m, n, _d(s["Adata"]), _i(s["Aindptr"]), _i(s["Aindices"]), _d(s["ATdata"]),
            _i(s["ATindptr"]), _i(s["ATindices"]), s["nnzL"], _i(s["Lindptr"]), _i(s["Lindices"]),
            _i(s["LTindptr"]), _i(s["LTindices"]), _i(s["LTmap"]), _d(b), _d(c), _d(x), _d(y), _d(z),
            _i(status), _i(iters), _d(trace), int(nthreads or os.cpu_count() or 1))
        return _Result(x=x, y=y, z=z, status=status, iters=self._steps(status, iters), trace=trace)

    # -- raw kernel hooks: interleaved (problem-minor) arrays as in tests/test_ldl.py ----
    def solve_primal_normal(self, A, x, z, y, b, c, mu, delta=1e-6):
        """x,z,c: (N,n); y,b: (N,m) per problem; interleaves like the reference tests."""
        A = _f64(A)
        m, n = A.shape
        xi, zi, yi, bi, ci = (_f64(np.atleast_2d(v).T) for v in (x, z, y, b, c))
        N = xi.shape[1]
        L = np.empty(N * m * (m + 1) // 2)
        D = np.empty(N * m)
        S = np.empty(N * m)
        dy = np.empty(N * m)
        self.lib.ref_solve_primal_normal(N, m, n, _d(A), _d(xi), _d(zi), _d(yi), _d(bi), _d(ci),
                                         ctypes.c_double(mu), _d(L), _d(D), _d(S), _d(dy),
                                         ctypes.c_double(delta))
        return (dy.reshape(m, N).T.copy(), L.reshape(-1, N).T.copy(), D.reshape(m, N).T.copy())

    def sparse_solve_primal_normal(self, A, x, z, y, b, c, mu, delta=1e-6, structures=None):
        A = _f64(A)
        m, n = A.shape
        s = structures or sparse_structures(A)
        xi, zi, yi, bi, ci = (_f64(np.atleast_2d(v).T) for v in (x, z, y, b, c))
        N = xi.shape[1]
        Ld = np.empty(N * s["nnzL"])
        D = np.empty(N * m)
        S = np.empty(N * m)
        dy = np.empty(N * m)
        self.lib.ref_sparse_solve_primal_normal(
            N, m, n, _d(s["Adata"]), _i(s["Aindptr"]), _i(s["Aindices"]), _d(s["ATdata"]),
            _i(s["ATindptr"]), _i(s["ATindices"]), _d(xi), _d(zi), _d(yi), _d(bi), _d(ci),
            ctypes.c_double(mu), _d(Ld), _i(s["Lindptr"]), _i(s["Lindices"]), _i(s["LTindptr"]),
            _i(s["LTindices"]), _i(s["LTmap"]), _d(D), _d(S), _d(dy), ctypes.c_double(delta))
        return dy.reshape(m, N).T.copy()

    def ldl(self, AA, modified=False, beta=1.0, delta=1e-6):
        """AA: (N, m, m) per problem -> interleaved (m, m, N) like tests/test_ldl.py:139."""
        AA = _f64(AA)
        N, m, _ = AA.shape
        Ai = _f64(np.transpose(AA, (1, 2, 0)))
        L = np.empty(N * m * (m + 1) // 2)
        D = np.empty(N * m)
        if modified:
            self.lib.ref_modified_ldl(N, m, m, _d(Ai), _d(L), _d(D), ctypes.c_double(beta),
                                      ctypes.c_double(delta))
        else:
            self.lib.ref_ldl(N, m, m, _d(Ai), _d(L), _d(D))
        return L.reshape(-1, N).T.copy(), D.reshape(m, N).T.copy()
