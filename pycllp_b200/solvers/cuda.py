"""CUDA-backed replacements of the reference's solvers of the primal normal-equations path.

Reference: ``pycllp/solvers/cl.py`` -- ``ClDensePrimalNormalSolver`` (``cl.py:12-124``) and
``ClSparsePrimalNormalSolver`` (``cl.py:127-278``) -- and the CPU twin of the same algorithm,
``DensePrimalNormalSolver`` (``solvers/normal_eqns.py:16-103`` with ``_ldl.pyx``).  Same registry
names, same ``init(lp)`` / ``solve(lp)`` protocol, same result attributes (``solver.x`` of shape
(nproblems, ncols) float64, ``solver.status`` int32 of shape (nproblems,)); ``solve`` returns
``None`` for the ``cl_*`` solvers and the status array for ``dense_primal_normal``, as the
reference's do.  Additionally ``solver.y``, ``solver.z`` (the dual solution, which the reference
leaves on the device) and ``solver.iterations`` are filled.

Every solve is a cold start from x = z = y = 1 exactly as in the reference (``cl.py:108,263``)
unless ``warm_start=True`` is passed to ``solve``: then the iteration starts from the x, z, y of
the previous solve, which stay resident on the GPU -- the repeat-solve use the library was
written for (``README.md:6``, ``primal_normal.cl:213-219``).  ``lp.f`` is ignored, as there.

Host buffers: the result arrays live in page-locked memory owned by the solver and ``lp.b`` /
``lp.c`` are page-locked in place on first use, so the per-solve copies are plain DMA
(the reference pays a host transpose plus pageable copies, ``cl.py:99-121``).

Multi-GPU: pass ``group=`` (a ``torch.distributed`` process group, or ``True`` for the
default group) with one process per GPU; each rank solves a contiguous slice of the
problems on its own device with its own replica of A and the slices are collected
with a single all-gather of packed records (see ``pycllp_b200/sharding.py``).
"""
from __future__ import print_function

import time

import numpy as np

from . import BaseSolver
from .._cabi import Engine
from .. import sharding

DTYPE = np.float64
IDTYPE = np.int32


class _CudaPrimalNormalBase(BaseSolver):
    name = None
    _sparse = False
    _preset = "cl"
    _returns_status = False

    def __init__(self, device=None, stream=None, group=None, **params):
        """``device``: CUDA device index (default: LOCAL_RANK or 0).

        The reference's constructor takes ``(ctx, queue)`` (``cl.py:18``); positional
        arguments that are not integers (e.g. pyopencl objects) are ignored so that
        existing call sites keep working.  ``params`` overrides algorithm constants
        (eps, delta, r, ldl_delta, refine_tol, max_iter, max_refine, nan_guard, carry_v, ...;
        ``preset='cl'|'py'`` selects one of the reference's two sets first).
        """
        super(_CudaPrimalNormalBase, self).__init__()
        if not isinstance(device, (int, np.integer)) or isinstance(device, bool):
            device = None
        self.group = group
        if device is None:
            device = sharding.default_device()
        self.device = int(device)
        self.engine = Engine(self.device)
        params = dict(params)
        self._preset = params.pop("preset", self._preset)
        # sparse path only: 'auto' | 'tiles' (L on its symbolic fill pattern) | 'dense'
        self._factor = params.pop("factor", "auto")
        self._ordering = params.pop("ordering", "auto")
        self._params = params
        self._span = None
        self._out = None
        self._torch = None

    # -- BaseSolver protocol ----------------------------------------------------------
    def init(self, lp, verbose=0):
        if verbose > 0:
            print("Initializing {} solver...".format(type(self).__name__))
        world, rank = sharding.world_and_rank(self.group)
        lo, hi = sharding.shard_bounds(lp.nproblems, world, rank)
        self._span = (lo, hi, world, rank)
        nlocal = max(hi - lo, 1)
        if self._sparse:
            self.engine.setup_sparse(lp.A.tocsr(), nlocal, factor=self._factor, ordering=self._ordering)
        else:
            self.engine.setup_dense(np.asarray(lp.A.todense(), dtype=DTYPE), nlocal)
        if self._preset != "cl":
            self.engine.set_preset(self._preset)
        if self._params:
            self.engine.set_params(**self._params)
        m, n, N = lp.nrows, lp.ncols, lp.nproblems
        if world == 1:
            # results in page-locked memory: device -> host copies without a staging pass
            eng = self.engine
            self._out = dict(x=eng.pinned_empty((N, n)), y=eng.pinned_empty((N, m)),
                             z=eng.pinned_empty((N, n)), status=eng.pinned_empty(N, IDTYPE),
                             iters=eng.pinned_empty(N, IDTYPE))
        self.status = np.empty(N, dtype=IDTYPE)
        self._solved = False
        if verbose > 0:
            print("Solver initialized.")

    def _pinned_input(self, arr):
        """lp.b / lp.c as they are (float64, C-contiguous: a view, no copy), page-locked in place."""
        a = np.ascontiguousarray(arr, dtype=DTYPE)
        self.engine.pin_in_place(a)
        return a

    def solve(self, lp, verbose=0, warm_start=False):
        if self._span is None or self.engine.m != lp.nrows or self.engine.n != lp.ncols:
            raise RuntimeError("solve() called before init() (or with a different LP)")
        if warm_start and not self._solved:
            raise RuntimeError("warm_start=True needs a previous solve() of this solver")
        lo, hi, world, rank = self._span
        if verbose > 0:
            print("Solving LP using {}...".format(type(self).__name__))
            t0 = time.time()
        trace_iters = int(self.engine.get_params().max_iter) if verbose > 1 else 0
        if world == 1:
            b, c = self._pinned_input(lp.b), self._pinned_input(lp.c)
            res = self.engine.solve_host(b, c, warm_start=warm_start, trace_iters=trace_iters,
                                         out=self._out)
        else:
            res = self._solve_sharded(lp, lo, hi, warm_start)
        if verbose > 0:
            print("Kernel complete in {} seconds.".format(time.time() - t0))
        self.x, self.y, self.z = res["x"], res["y"], res["z"]
        self.status = res["status"]
        self.iterations = res["iters"]
        self._solved = True
        if verbose > 1:
            # what the kernels print from every work-item at verbose > 1 (primal_normal.cl:250-252)
            tr = res.get("trace")
            for q in range(lp.nproblems):
                if tr is not None:
                    for it in range(min(int(self.iterations[q]) + 1, tr.shape[1])):
                        if not np.isnan(tr[q, it, 0]):
                            print("{:d}/{:d} {:d} |rho|: {:8.1e}  |sigma| {:8.1e}  gamma: {:8.1e}".format(
                                q, lp.nproblems, it, tr[q, it, 0], tr[q, it, 1], tr[q, it, 2]))
                print("{}/{} iterations: {:3d} status: {}".format(
                    q, lp.nproblems, int(self.iterations[q]), int(self.status[q])))
        if verbose > 0:
            print("Solve complete.")
        return self.status if self._returns_status else None

    # -- one process per GPU ------------------------------------------------------------
    def _solve_sharded(self, lp, lo, hi, warm_start):
        """This rank's slice on its GPU into a packed record block, ONE all-gather of the blocks
        straight from the engine's output buffer (NCCL over NVLink), one copy to the host."""
        import torch
        m, n, N = lp.nrows, lp.ncols, lp.nproblems
        dev = torch.device("cuda", self.device)
        nloc = hi - lo
        st = self._torch
        if st is None or st["rec"].shape[0] != max(nloc, 1):
            st = self._torch = dict(rec=torch.zeros((max(nloc, 1), sharding.record_width(m, n)),
                                                    dtype=torch.float64, device=dev))
        rec = st["rec"]
        if nloc > 0:
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream(dev)
                b = torch.from_numpy(self._pinned_input(lp.b)[lo:hi]).to(dev, non_blocking=True)
                c = torch.from_numpy(self._pinned_input(lp.c)[lo:hi]).to(dev, non_blocking=True)
                # (warm start: x, y, z of the previous solve are still in this rank's records)
                self.engine.solve_device_packed(nloc, b.data_ptr(), c.data_ptr(), rec.data_ptr(),
                                                stream.cuda_stream, warm_start=warm_start)
        full = sharding.allgather_records(rec[:nloc], N, self.group)
        # one DMA into page-locked memory; x, y, z are handed out as views of that block
        if st.get("host") is None or st["host"].shape != tuple(full.shape):
            st["host"] = self.engine.pinned_empty(tuple(full.shape))
            st["host_t"] = torch.from_numpy(st["host"])
        st["host_t"].copy_(full, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return sharding.unpack_records(st["host"], m, n, copy=False)


class CudaDensePrimalNormalSolver(_CudaPrimalNormalBase):
    """Dense-A path (reference ``ClDensePrimalNormalSolver``, ``cl.py:12-124``)."""
    name = 'cl_dense_primal_normal'
    _sparse = False


class CudaSparsePrimalNormalSolver(_CudaPrimalNormalBase):
    """Sparse-A path (reference ``ClSparsePrimalNormalSolver``, ``cl.py:127-278``):
    CSR mat-vecs, M formed on the shared pattern of A A', no iterative refinement
    (``ldl.cl:698-711``).  ``factor='tiles'`` (or 'auto' with a sparse enough fill) computes and
    stores L on its symbolic pattern only -- the batched counterpart of
    ``sparse_factor_primal_normal`` (``ldl.cl:381-502``); ``engine.sparse_info()`` reports it."""
    name = 'cl_sparse_primal_normal'
    _sparse = True


class CudaDensePrimalNormalPySolver(_CudaPrimalNormalBase):
    """The constants and conventions of the reference's CPU twin of the algorithm,
    ``DensePrimalNormalSolver`` (``solvers/normal_eqns.py:16-103`` + ``_ldl.pyx:35-152``), on the
    same CUDA engine: EPS 1e-8, delta 0.1, mu = delta gamma / n, refinement tolerance 1e-6 with
    that code's sign convention, no floor on theta, NaN in dy => status 3.  ``solve`` returns the
    status array like the reference's (``normal_eqns.py:33``).  The statuses of the Python
    implementation are rounding-noise driven, so this preset is compared with it
    on objective values, not on status."""
    name = 'dense_primal_normal'
    _sparse = False
    _preset = "py"
    _returns_status = True
