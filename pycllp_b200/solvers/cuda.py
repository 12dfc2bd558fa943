"""CUDA-backed replacements of the reference's two OpenCL solvers.

Reference: ``pycllp/solvers/cl.py`` -- ``ClDensePrimalNormalSolver`` (``cl.py:12-124``)
and ``ClSparsePrimalNormalSolver`` (``cl.py:127-278``).  Same registry names, same
``init(lp)`` / ``solve(lp)`` protocol, same result attributes (``solver.x`` of shape
(nproblems, ncols) float64, ``solver.status`` int32 of shape (nproblems,)); ``solve``
returns ``None`` like the OpenCL solvers do.  Additionally ``solver.y``, ``solver.z``
(the dual solution, which the reference leaves on the device) and
``solver.iterations`` are filled.

Every solve is a cold start from x = z = y = 1 exactly as in the reference
(``cl.py:108,263``).  ``lp.f`` is ignored, as there.

Multi-GPU: pass ``group=`` (a ``torch.distributed`` process group, or ``True`` for the
default group) with one process per GPU; each rank solves a contiguous slice of the
problems on its own device with its own replica of A and the slices are collected
with a single all-gather (see ``pycllp_b200/sharding.py``).
"""
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

    def __init__(self, device=None, stream=None, group=None, **params):
        """``device``: CUDA device index (default: LOCAL_RANK or 0).

        The reference's constructor takes ``(ctx, queue)`` (``cl.py:18``); positional
        arguments that are not integers (e.g. pyopencl objects) are ignored so that
        existing call sites keep working.  ``params`` overrides algorithm constants
        (eps, delta, r, ldl_delta, refine_tol, max_iter, max_refine).
        """
        super(_CudaPrimalNormalBase, self).__init__()
        if not isinstance(device, (int, np.integer)) or isinstance(device, bool):
            device = None
        self.group = group
        if device is None:
            device = sharding.default_device()
        self.device = int(device)
        self.engine = Engine(self.device)
        self._params = dict(params)
        self._span = None

    # -- BaseSolver protocol ----------------------------------------------------------
    def init(self, lp, verbose=0):
        if verbose > 0:
            print("Initializing {} solver...".format(type(self).__name__))
        world, rank = sharding.world_and_rank(self.group)
        lo, hi = sharding.shard_bounds(lp.nproblems, world, rank)
        self._span = (lo, hi, world, rank)
        nlocal = max(hi - lo, 1)
        if self._sparse:
            self.engine.setup_sparse(lp.A.tocsr(), nlocal)
        else:
            self.engine.setup_dense(np.asarray(lp.A.todense(), dtype=DTYPE), nlocal)
        if self._params:
            self.engine.set_params(**self._params)
        self.status = np.empty(lp.nproblems, dtype=IDTYPE)
        if verbose > 0:
            print("Solver initialized.")

    def solve(self, lp, verbose=0):
        if self._span is None or self.engine.m != lp.nrows or self.engine.n != lp.ncols:
            raise RuntimeError("solve() called before init() (or with a different LP)")
        lo, hi, world, rank = self._span
        if verbose > 0:
            print("Solving LP using {}...".format(type(self).__name__))
            t0 = time.time()
        if hi > lo:
            res = self.engine.solve_host(lp.b[lo:hi], lp.c[lo:hi])
        else:
            res = dict(x=np.empty((0, lp.ncols)), y=np.empty((0, lp.nrows)),
                       z=np.empty((0, lp.ncols)), status=np.empty(0, dtype=IDTYPE),
                       iters=np.empty(0, dtype=IDTYPE))
        if world > 1:
            res = sharding.allgather_results(res, lp.nproblems, self.group, device=self.device)
        if verbose > 0:
            print("Kernel complete in {} seconds.".format(time.time() - t0))
        self.x, self.y, self.z = res["x"], res["y"], res["z"]
        self.status = res["status"]
        self.iterations = res["iters"]
        if verbose > 1:
            for q in range(lp.nproblems):
                print("{}/{} iterations: {:3d} status: {}".format(
                    q, lp.nproblems, int(self.iterations[q]), int(self.status[q])))
        if verbose > 0:
            print("Solve complete.")


class CudaDensePrimalNormalSolver(_CudaPrimalNormalBase):
    """Dense-A path (reference ``ClDensePrimalNormalSolver``, ``cl.py:12-124``)."""
    name = 'cl_dense_primal_normal'
    _sparse = False


class CudaSparsePrimalNormalSolver(_CudaPrimalNormalBase):
    """Sparse-A path (reference ``ClSparsePrimalNormalSolver``, ``cl.py:127-278``):
    CSR mat-vecs, M formed on the shared pattern of A A', no iterative refinement
    (``ldl.cl:698-711``)."""
    name = 'cl_sparse_primal_normal'
    _sparse = True
