"""Thin ctypes binding of the C ABI in ``include/pycllp_b200.h``.

This is what replaces the reference's ``pycllp/cl_tools.py`` (reading ``.cl`` text into
a ``pyopencl.Program``, ``cl_tools.py:11-29``) and the pyopencl calls in
``pycllp/solvers/cl.py``: it loads the in-tree CUDA shared library and exposes an
``Engine`` object.  There is no CPU fallback -- if the library is missing or no CUDA
device is usable, construction raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpycllp_b200.so")

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_vp = ctypes.c_void_p

EXPORTS = (
    "pycllp_b200_create", "pycllp_b200_destroy", "pycllp_b200_last_error", "pycllp_b200_version",
    "pycllp_b200_setup_dense", "pycllp_b200_setup_sparse", "pycllp_b200_set_params",
    "pycllp_b200_get_params", "pycllp_b200_solve_host", "pycllp_b200_solve_device",
    "pycllp_b200_solve_primal_normal", "pycllp_b200_ldl", "pycllp_b200_launch_count",
    "pycllp_b200_info", "pycllp_b200_phase_profile",
)


class Params(ctypes.Structure):
    """Mirror of ``pycllp_b200_params``."""
    _fields_ = [("eps", ctypes.c_double), ("delta", ctypes.c_double), ("r", ctypes.c_double),
                ("ldl_delta", ctypes.c_double), ("refine_tol", ctypes.c_double),
                ("max_iter", ctypes.c_int), ("max_refine", ctypes.c_int)]


_lib = None


def load_library():
    """Load ``libpycllp_b200.so`` (built in-tree by ``pycllp_b200.build``). Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "pycllp_b200: CUDA engine %s is missing; build it with `python -m pycllp_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.pycllp_b200_last_error.restype = ctypes.c_char_p
    lib.pycllp_b200_last_error.argtypes = [_vp]
    lib.pycllp_b200_version.restype = ctypes.c_char_p
    lib.pycllp_b200_create.argtypes = [ctypes.c_int, ctypes.POINTER(_vp)]
    lib.pycllp_b200_destroy.argtypes = [_vp]
    lib.pycllp_b200_setup_dense.argtypes = [_vp, ctypes.c_int, ctypes.c_int, _dp, ctypes.c_int]
    lib.pycllp_b200_setup_sparse.argtypes = [_vp, ctypes.c_int, ctypes.c_int, _ip, _ip, _dp, ctypes.c_int]
    lib.pycllp_b200_set_params.argtypes = [_vp, ctypes.POINTER(Params)]
    lib.pycllp_b200_get_params.argtypes = [_vp, ctypes.POINTER(Params)]
    lib.pycllp_b200_solve_host.argtypes = [_vp, ctypes.c_int, _dp, _dp, _dp, _dp, _dp, _ip, _ip]
    lib.pycllp_b200_solve_device.argtypes = [_vp, ctypes.c_int] + [_vp] * 8
    lib.pycllp_b200_solve_primal_normal.argtypes = [_vp, ctypes.c_int, _dp, _dp, _dp, _dp, _dp,
                                                    ctypes.c_double, _dp]
    lib.pycllp_b200_ldl.argtypes = [_vp, ctypes.c_int, ctypes.c_int, _dp, _dp, _dp, ctypes.c_int,
                                    ctypes.c_double, ctypes.c_double]
    lib.pycllp_b200_launch_count.restype = ctypes.c_longlong
    lib.pycllp_b200_launch_count.argtypes = [_vp]
    lib.pycllp_b200_info.argtypes = [_vp, _ip, _ip, _ip, ctypes.POINTER(ctypes.c_size_t),
                                     ctypes.POINTER(ctypes.c_size_t), _ip]
    lib.pycllp_b200_phase_profile.argtypes = [_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong)]
    _lib = lib
    return lib


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Engine(object):
    """One CUDA device's batched IPM engine (opaque handle + checked calls)."""

    def __init__(self, device=0):
        self._lib = load_library()
        self._h = _vp()
        rc = self._lib.pycllp_b200_create(int(device), ctypes.byref(self._h))
        if rc != 0:
            msg = self._lib.pycllp_b200_last_error(None)
            raise RuntimeError("pycllp_b200_create failed (%d): %s" % (rc, (msg or b"").decode()))
        self.device = int(device)
        self.m = self.n = self.max_problems = 0
        self.sparse = False

    # -- plumbing ----------------------------------------------------------------
    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.pycllp_b200_last_error(self._h)
            raise RuntimeError("%s failed (%d): %s" % (what, rc, (msg or b"").decode()))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pycllp_b200_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setup --------------------------------------------------------------------
    def setup_dense(self, A, max_problems):
        A = _f64(A)
        m, n = A.shape
        self._check(self._lib.pycllp_b200_setup_dense(self._h, m, n, _d(A), int(max_problems)),
                    "pycllp_b200_setup_dense")
        self.m, self.n, self.max_problems, self.sparse = m, n, int(max_problems), False

    def setup_sparse(self, csr, max_problems):
        """``csr``: scipy.sparse CSR matrix (sorted indices)."""
        csr = csr.tocsr()
        csr.sort_indices()
        m, n = csr.shape
        indptr = np.ascontiguousarray(csr.indptr, dtype=np.int32)
        indices = np.ascontiguousarray(csr.indices, dtype=np.int32)
        data = _f64(csr.data)
        self._check(self._lib.pycllp_b200_setup_sparse(self._h, m, n, _i(indptr), _i(indices),
                                                       _d(data), int(max_problems)),
                    "pycllp_b200_setup_sparse")
        self.m, self.n, self.max_problems, self.sparse = m, n, int(max_problems), True

    def get_params(self):
        p = Params()
        self._check(self._lib.pycllp_b200_get_params(self._h, ctypes.byref(p)), "get_params")
        return p

    def set_params(self, **kw):
        p = self.get_params()
        for k, v in kw.items():
            if not hasattr(p, k):
                raise TypeError("unknown parameter %r" % k)
            setattr(p, k, v)
        self._check(self._lib.pycllp_b200_set_params(self._h, ctypes.byref(p)), "set_params")

    # -- solve ----------------------------------------------------------------------
    def solve_host(self, b, c, want_yz=True):
        """b (N, m), c (N, n) numpy -> dict(x, y, z, status, iters) numpy."""
        b, c = _f64(np.atleast_2d(b)), _f64(np.atleast_2d(c))
        N = b.shape[0]
        if b.shape != (N, self.m) or c.shape != (N, self.n):
            raise ValueError("b must be (N, %d) and c (N, %d)" % (self.m, self.n))
        x = np.empty((N, self.n))
        y = np.empty((N, self.m)) if want_yz else None
        z = np.empty((N, self.n)) if want_yz else None
        status = np.empty(N, dtype=np.int32)
        iters = np.empty(N, dtype=np.int32)
        self._check(self._lib.pycllp_b200_solve_host(self._h, N, _d(b), _d(c), _d(x), _d(y), _d(z),
                                                     _i(status), _i(iters)),
                    "pycllp_b200_solve_host")
        return dict(x=x, y=y, z=z, status=status, iters=iters)

    def solve_host_into(self, N, b, c, x, y, z, status, iters):
        """Raw-pointer variant for pinned buffers: every argument is an int address or None."""
        cast = lambda p, t: ctypes.cast(ctypes.c_void_p(p), t) if p else None
        self._check(self._lib.pycllp_b200_solve_host(
            self._h, int(N), cast(b, _dp), cast(c, _dp), cast(x, _dp), cast(y, _dp), cast(z, _dp),
            cast(status, _ip), cast(iters, _ip)), "pycllp_b200_solve_host")

    def solve_device(self, N, d_b, d_c, d_x=0, d_y=0, d_z=0, d_status=0, d_iters=0, stream=0):
        """Device pointers (ints, e.g. torch ``tensor.data_ptr()``); enqueues on ``stream``."""
        self._check(self._lib.pycllp_b200_solve_device(
            self._h, int(N), _vp(d_b), _vp(d_c), _vp(d_x or None), _vp(d_y or None),
            _vp(d_z or None), _vp(d_status or None), _vp(d_iters or None), _vp(stream or None)),
            "pycllp_b200_solve_device")

    # -- kernel-level hooks -------------------------------------------------------------
    def solve_primal_normal(self, x, z, y, b, c, mu):
        x, z, y, b, c = (_f64(np.atleast_2d(v)) for v in (x, z, y, b, c))
        N = x.shape[0]
        dy = np.empty((N, self.m))
        self._check(self._lib.pycllp_b200_solve_primal_normal(self._h, N, _d(x), _d(z), _d(y), _d(b),
                                                              _d(c), float(mu), _d(dy)),
                    "pycllp_b200_solve_primal_normal")
        return dy

    def ldl(self, AA, modified=False, beta=1.0, delta=1e-6):
        AA = _f64(AA)
        N, m, _ = AA.shape
        L = np.empty((N, m * (m + 1) // 2))
        D = np.empty((N, m))
        self._check(self._lib.pycllp_b200_ldl(self._h, N, m, _d(AA), _d(L), _d(D), int(modified),
                                              float(beta), float(delta)), "pycllp_b200_ldl")
        return L, D

    def phase_profile(self, enable=True):
        """Per-phase SM cycles since the last call (dict), then (re)arm or disarm the counters."""
        out = (ctypes.c_ulonglong * 16)()
        self._check(self._lib.pycllp_b200_phase_profile(self._h, int(enable), out), "phase_profile")
        names = ("rhs_norms", "form_M", "factor", "tri_solve", "residual", "step",
                 "f_copy", "f_waitEd", "f_blockrow", "f_diag", "f_solve", "f_table", "f_waitE3", "f_old", "f_waitEb", "f_update")
        return {k: int(out[i]) for i, k in enumerate(names)}

    # -- introspection ---------------------------------------------------------------------
    @property
    def launch_count(self):
        return int(self._lib.pycllp_b200_launch_count(self._h))

    def info(self):
        sms, grid, block, fac = (ctypes.c_int() for _ in range(4))
        smem, scratch = ctypes.c_size_t(), ctypes.c_size_t()
        self._check(self._lib.pycllp_b200_info(self._h, ctypes.byref(sms), ctypes.byref(grid),
                                               ctypes.byref(block), ctypes.byref(smem),
                                               ctypes.byref(scratch), ctypes.byref(fac)), "info")
        return dict(num_sms=sms.value, grid=grid.value, block=block.value, smem_bytes=smem.value,
                    scratch_bytes=scratch.value, factor_in_smem=bool(fac.value))
