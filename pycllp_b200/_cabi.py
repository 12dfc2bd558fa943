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
    "pycllp_b200_info", "pycllp_b200_phase_profile", "pycllp_b200_set_preset",
    "pycllp_b200_solve_host_ex", "pycllp_b200_solve_device_ex", "pycllp_b200_host_alloc",
    "pycllp_b200_host_free", "pycllp_b200_host_register", "pycllp_b200_host_unregister",
    "pycllp_b200_solve_device_packed", "pycllp_b200_fp64_probe", "pycllp_b200_set_sparse_factor",
    "pycllp_b200_sparse_info", "pycllp_b200_tile_analysis", "pycllp_b200_sparse_ldl",
    "pycllp_b200_set_small_kernels", "pycllp_b200_set_sparse_ordering", "pycllp_b200_sparse_reordered",
    "pycllp_b200_rcm_ordering",
)


class Params(ctypes.Structure):
    """Mirror of ``pycllp_b200_params``."""
    _fields_ = [("eps", ctypes.c_double), ("delta", ctypes.c_double), ("r", ctypes.c_double),
                ("ldl_delta", ctypes.c_double), ("refine_tol", ctypes.c_double),
                ("max_iter", ctypes.c_int), ("max_refine", ctypes.c_int),
                ("nan_guard", ctypes.c_int), ("carry_v", ctypes.c_int), ("mu_mode", ctypes.c_int),
                ("refine_mode", ctypes.c_int), ("theta_floor", ctypes.c_int), ("dz_mode", ctypes.c_int),
                ("warm_floor", ctypes.c_double)]


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
    lib.pycllp_b200_set_preset.argtypes = [_vp, ctypes.c_char_p]
    lib.pycllp_b200_solve_host_ex.argtypes = [_vp, ctypes.c_int, _dp, _dp, ctypes.c_int, _dp, _dp, _dp,
                                              _ip, _ip, _dp, ctypes.c_int]
    lib.pycllp_b200_solve_device_ex.argtypes = [_vp, ctypes.c_int] + [_vp] * 11 + [ctypes.c_int, _vp]
    lib.pycllp_b200_solve_device_packed.argtypes = [_vp, ctypes.c_int, _vp, _vp, _vp, ctypes.c_int, _vp]
    lib.pycllp_b200_fp64_probe.argtypes = [_vp, _dp]
    lib.pycllp_b200_host_alloc.argtypes = [_vp, ctypes.c_size_t, ctypes.POINTER(_vp)]
    lib.pycllp_b200_host_free.argtypes = [_vp, _vp]
    lib.pycllp_b200_host_register.argtypes = [_vp, _vp, ctypes.c_size_t]
    lib.pycllp_b200_host_unregister.argtypes = [_vp, _vp]
    lib.pycllp_b200_solve_primal_normal.argtypes = [_vp, ctypes.c_int, _dp, _dp, _dp, _dp, _dp,
                                                    ctypes.c_double, _dp]
    lib.pycllp_b200_ldl.argtypes = [_vp, ctypes.c_int, ctypes.c_int, _dp, _dp, _dp, ctypes.c_int,
                                    ctypes.c_double, ctypes.c_double]
    lib.pycllp_b200_launch_count.restype = ctypes.c_longlong
    lib.pycllp_b200_launch_count.argtypes = [_vp]
    lib.pycllp_b200_info.argtypes = [_vp, _ip, _ip, _ip, ctypes.POINTER(ctypes.c_size_t),
                                     ctypes.POINTER(ctypes.c_size_t), _ip]
    lib.pycllp_b200_phase_profile.argtypes = [_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong)]
    lib.pycllp_b200_tile_analysis.argtypes = [ctypes.c_int, ctypes.c_int, _ip, _ip, _ip, _ip,
                                              ctypes.POINTER(ctypes.c_longlong)] + [_ip] * 6
    lib.pycllp_b200_sparse_ldl.argtypes = [_vp, ctypes.c_int, ctypes.c_int, _ip, _ip, _dp, _dp, _dp,
                                           ctypes.c_double, ctypes.c_double]
    lib.pycllp_b200_set_sparse_factor.argtypes = [_vp, ctypes.c_int]
    lib.pycllp_b200_set_small_kernels.argtypes = [_vp, ctypes.c_int]
    lib.pycllp_b200_set_sparse_ordering.argtypes = [_vp, ctypes.c_int]
    lib.pycllp_b200_sparse_reordered.argtypes = [_vp]
    lib.pycllp_b200_rcm_ordering.argtypes = [ctypes.c_int, ctypes.c_int, _ip, _ip, _ip]
    lib.pycllp_b200_sparse_info.argtypes = [_vp, _ip, ctypes.POINTER(ctypes.c_longlong),
                                            ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong),
                                            _dp]
    _lib = lib
    return lib


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def rcm_ordering(csr):
    """Host-only: the band-reducing ordering of the constraints ``setup_sparse`` considers for the
    tile-sparse factor; ``perm[i]`` = the row of ``csr`` placed at position i."""
    lib = load_library()
    csr = csr.tocsr()
    m, n = csr.shape
    indptr = np.ascontiguousarray(csr.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(csr.indices, dtype=np.int32)
    perm = np.zeros(m, np.int32)
    rc = lib.pycllp_b200_rcm_ordering(m, n, _i(indptr), _i(indices), _i(perm))
    if rc != 0:
        raise RuntimeError("pycllp_b200_rcm_ordering failed (%d)" % rc)
    return perm


def tile_analysis(csr):
    """Host-only symbolic analysis of the tile-sparse factor for the pattern of ``csr`` (what
    ``setup_sparse`` computes for ``factor='tiles'``); returns the tile structure as numpy arrays."""
    lib = load_library()
    csr = csr.tocsr()
    csr.sort_indices()
    m, n = csr.shape
    indptr = np.ascontiguousarray(csr.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(csr.indices, dtype=np.int32)
    nbk, nt, pairs = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_longlong(0)
    null = ctypes.POINTER(ctypes.c_int)()
    rc = lib.pycllp_b200_tile_analysis(m, n, _i(indptr), _i(indices), ctypes.byref(nbk), ctypes.byref(nt),
                                       ctypes.byref(pairs), null, null, null, null, null, null)
    if rc != 0:
        raise RuntimeError("pycllp_b200_tile_analysis failed (%d)" % rc)
    out = dict(nbk=nbk.value, ntiles=nt.value, pairs=pairs.value,
               colptr=np.zeros(nbk.value + 1, np.int32), row=np.zeros(nt.value, np.int32),
               col=np.zeros(nt.value, np.int32), updptr=np.zeros(nt.value + 1, np.int32),
               upda=np.zeros(max(pairs.value, 1), np.int32), updb=np.zeros(max(pairs.value, 1), np.int32))
    rc = lib.pycllp_b200_tile_analysis(m, n, _i(indptr), _i(indices), ctypes.byref(nbk), ctypes.byref(nt),
                                       ctypes.byref(pairs), _i(out["colptr"]), _i(out["row"]), _i(out["col"]),
                                       _i(out["updptr"]), _i(out["upda"]), _i(out["updb"]))
    if rc != 0:
        raise RuntimeError("pycllp_b200_tile_analysis failed (%d)" % rc)
    out["upda"] = out["upda"][:pairs.value]
    out["updb"] = out["updb"][:pairs.value]
    return out


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
        self._pinned = []        # addresses from host_alloc
        self._registered = {}    # address -> nbytes of host_register'ed caller buffers

    # -- plumbing ----------------------------------------------------------------
    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.pycllp_b200_last_error(self._h)
            raise RuntimeError("%s failed (%d): %s" % (what, rc, (msg or b"").decode()))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            for addr in list(self._registered):
                self._lib.pycllp_b200_host_unregister(self._h, _vp(addr))
            self._registered.clear()
            for addr in self._pinned:
                self._lib.pycllp_b200_host_free(self._h, _vp(addr))
            self._pinned = []
            self._lib.pycllp_b200_destroy(self._h)
            self._h = _vp()

    # -- page-locked host memory -----------------------------------------------------
    def pinned_empty(self, shape, dtype=np.float64):
        """A numpy array in page-locked host memory (owned by this engine, freed in close())."""
        dtype = np.dtype(dtype)
        shape = tuple(int(v) for v in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        ptr = _vp()
        self._check(self._lib.pycllp_b200_host_alloc(self._h, max(nbytes, 1), ctypes.byref(ptr)),
                    "pycllp_b200_host_alloc")
        self._pinned.append(ptr.value)
        buf = (ctypes.c_char * max(nbytes, 1)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)

    def pin_in_place(self, arr):
        """Page-lock a caller's C-contiguous array where it lies (cached; returns False when the
        driver refuses, in which case the copies simply stay pageable)."""
        addr, nbytes = arr.ctypes.data, arr.nbytes
        if nbytes == 0:
            return False
        if self._registered.get(addr, 0) >= nbytes:
            return True
        if addr in self._registered:
            self._lib.pycllp_b200_host_unregister(self._h, _vp(addr))
            del self._registered[addr]
        if self._lib.pycllp_b200_host_register(self._h, _vp(addr), nbytes) != 0:
            return False
        self._registered[addr] = nbytes
        return True

    def unpin_all(self):
        for addr in list(self._registered):
            self._lib.pycllp_b200_host_unregister(self._h, _vp(addr))
        self._registered.clear()

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

    def set_small_kernels(self, mode):
        """0: one 512-thread block per SM always; 1 (default): small-problem kernels where they fit
        (128-thread kernel for dense m <= 63 and batches of more than 1.5 LPs per SM, else two
        blocks per SM); 2: only the latter; 3: the 128-thread kernel for every batch size.
        Takes effect at the next setup."""
        self._check(self._lib.pycllp_b200_set_small_kernels(self._h, int(mode)), "pycllp_b200_set_small_kernels")

    SPARSE_FACTOR = {"auto": 0, "tiles": 1, "dense": 2}

    SPARSE_ORDERING = {"auto": 0, "natural": 1, "rcm": 2}

    def setup_sparse(self, csr, max_problems, factor="auto", ordering="auto"):
        """``csr``: scipy.sparse CSR matrix (sorted indices).  ``factor``: numeric factor of the
        sparse path -- 'tiles' (L on its symbolic fill pattern, 8x8 tiles, memory ~ nnz(L)),
        'dense' (packed dense kernels) or 'auto'.  ``ordering`` (tiles only): 'natural' (the order
        the constraints come in, as the reference), 'rcm' (band-reducing reordering of the
        constraints) or 'auto' (RCM when it gives fewer tiles)."""
        self._check(self._lib.pycllp_b200_set_sparse_factor(self._h, self.SPARSE_FACTOR[factor]),
                    "pycllp_b200_set_sparse_factor")
        self._check(self._lib.pycllp_b200_set_sparse_ordering(self._h, self.SPARSE_ORDERING[ordering]),
                    "pycllp_b200_set_sparse_ordering")
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

    def sparse_info(self):
        """Which numeric factor the sparse path uses and what it stores per LP."""
        mode = ctypes.c_int(0)
        fd, dd, up = ctypes.c_longlong(0), ctypes.c_longlong(0), ctypes.c_longlong(0)
        fill = ctypes.c_double(0.0)
        self._check(self._lib.pycllp_b200_sparse_info(self._h, ctypes.byref(mode), ctypes.byref(fd),
                                                      ctypes.byref(dd), ctypes.byref(up), ctypes.byref(fill)),
                    "pycllp_b200_sparse_info")
        return {"factor": "tiles" if mode.value else "dense",
                "ordering": "rcm" if self._lib.pycllp_b200_sparse_reordered(self._h) else "natural",
                "factor_doubles": fd.value,
                "dense_factor_doubles": dd.value, "update_pairs": up.value, "tile_fill": fill.value}

    def get_params(self):
        p = Params()
        self._check(self._lib.pycllp_b200_get_params(self._h, ctypes.byref(p)), "get_params")
        return p

    def set_preset(self, name):
        """'cl' (the OpenCL kernels' constants, default) or 'py' (solvers/normal_eqns.py)."""
        self._check(self._lib.pycllp_b200_set_preset(self._h, name.encode()), "pycllp_b200_set_preset")

    def set_params(self, **kw):
        p = self.get_params()
        for k, v in kw.items():
            if not hasattr(p, k):
                raise TypeError("unknown parameter %r" % k)
            setattr(p, k, v)
        self._check(self._lib.pycllp_b200_set_params(self._h, ctypes.byref(p)), "set_params")

    # -- solve ----------------------------------------------------------------------
    def solve_host(self, b, c, want_yz=True, warm_start=False, trace_iters=0, out=None):
        """b (N, m), c (N, n) numpy -> dict(x, y, z, status, iters[, trace]) numpy.

        warm_start: begin from the x, y, z the previous solve left on the device (same N).
        trace_iters > 0: also ``trace`` (N, trace_iters, 3) = |rho|, |sigma|, gamma per iteration.
        out: dict of preallocated result arrays (e.g. pinned) to fill instead of new ones."""
        b, c = _f64(np.atleast_2d(b)), _f64(np.atleast_2d(c))
        N = b.shape[0]
        if b.shape != (N, self.m) or c.shape != (N, self.n):
            raise ValueError("b must be (N, %d) and c (N, %d)" % (self.m, self.n))
        out = out or {}
        x = out.get("x") if out.get("x") is not None else np.empty((N, self.n))
        y = out.get("y") if out.get("y") is not None else (np.empty((N, self.m)) if want_yz else None)
        z = out.get("z") if out.get("z") is not None else (np.empty((N, self.n)) if want_yz else None)
        status = out.get("status") if out.get("status") is not None else np.empty(N, dtype=np.int32)
        iters = out.get("iters") if out.get("iters") is not None else np.empty(N, dtype=np.int32)
        trace = np.empty((N, int(trace_iters), 3)) if trace_iters else None
        self._check(self._lib.pycllp_b200_solve_host_ex(
            self._h, N, _d(b), _d(c), int(bool(warm_start)), _d(x), _d(y), _d(z), _i(status), _i(iters),
            _d(trace), int(trace_iters)), "pycllp_b200_solve_host_ex")
        res = dict(x=x, y=y, z=z, status=status, iters=iters)
        if trace is not None:
            res["trace"] = trace
        return res

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

    def solve_device_ex(self, N, d_b, d_c, d_x0=0, d_z0=0, d_y0=0, d_x=0, d_y=0, d_z=0, d_status=0,
                        d_iters=0, d_trace=0, trace_iters=0, stream=0):
        """Device pointers; x0/z0/y0 (all or none) give a warm start and may alias the outputs."""
        v = lambda p: _vp(p or None)
        self._check(self._lib.pycllp_b200_solve_device_ex(
            self._h, int(N), v(d_b), v(d_c), v(d_x0), v(d_z0), v(d_y0), v(d_x), v(d_y), v(d_z),
            v(d_status), v(d_iters), v(d_trace), int(trace_iters), v(stream)),
            "pycllp_b200_solve_device_ex")

    def solve_device_packed(self, N, d_b, d_c, d_rec, stream=0, warm_start=False):
        """One record per problem, (N, 2n+m+1) float64: x | y | z | (status, iterations as int32)."""
        self._check(self._lib.pycllp_b200_solve_device_packed(
            self._h, int(N), _vp(d_b), _vp(d_c), _vp(d_rec), int(bool(warm_start)), _vp(stream or None)),
            "pycllp_b200_solve_device_packed")

    @property
    def record_width(self):
        return 2 * self.n + self.m + 1

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

    def sparse_ldl(self, AA, indptr, indices, beta=0.0, delta=1e-6):
        """Modified LDL' on a given CSR-lower pattern (diagonal last per row): (Ldata (N, nnz), D (N, m))."""
        AA = _f64(AA)
        if AA.ndim == 2:
            AA = AA[None]
        N, m, _ = AA.shape
        indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        L = np.empty((N, len(indices)))
        D = np.empty((N, m))
        self._check(self._lib.pycllp_b200_sparse_ldl(self._h, N, m, _i(indptr), _i(indices), _d(AA), _d(L), _d(D),
                                                     float(beta), float(delta)), "pycllp_b200_sparse_ldl")
        return L, D

    def phase_profile(self, enable=True):
        """Per-phase SM cycles since the last call (dict), then (re)arm or disarm the counters."""
        out = (ctypes.c_ulonglong * 16)()
        self._check(self._lib.pycllp_b200_phase_profile(self._h, int(enable), out), "phase_profile")
        names = ("rhs_norms", "form_M", "factor", "tri_solve", "residual", "step",
                 "f_copy", "f_waitEd", "f_blockrow", "f_diag", "f_solve", "f_table", "f_waitE3", "f_old", "f_waitEb", "f_update")
        return {k: int(out[i]) for i, k in enumerate(names)}

    def fp64_probe(self):
        """FP64 tensor-core (DMMA) TFLOP/s of this device, measured now."""
        v = ctypes.c_double()
        self._check(self._lib.pycllp_b200_fp64_probe(self._h, ctypes.byref(v)), "pycllp_b200_fp64_probe")
        return float(v.value)

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
