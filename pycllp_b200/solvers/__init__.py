"""Solver plugin registry -- the drop-in boundary of the hot path.

Restates the contract of the reference's ``pycllp/solvers/__init__.py:3-21``: every
``BaseSolver`` subclass with a ``name`` registers itself in ``solver_registry`` and
implements ``init(lp, verbose=0)`` / ``solve(lp, verbose=0)``; ``EqualityLP.init`` /
``.solve`` dispatch to them (``lp.py:531-535``).

The two OpenCL solvers of the reference (``solvers/cl.py``) are provided here under
their original registry names, backed by the CUDA engine:

    solver_registry['cl_dense_primal_normal']   -> CudaDensePrimalNormalSolver
    solver_registry['cl_sparse_primal_normal']  -> CudaSparsePrimalNormalSolver
"""

solver_registry = {}


class MetaSolver(type):
    """Registers each named solver class (reference ``solvers/__init__.py:6-11``)."""

    def __new__(mcs, clsname, bases, attrs):
        cls = super(MetaSolver, mcs).__new__(mcs, clsname, bases, attrs)
        if cls.name is not None:
            solver_registry[cls.name] = cls
        return cls


class BaseSolver(metaclass=MetaSolver):
    name = None

    def init(self, lp, verbose=0):
        raise NotImplementedError()

    def solve(self, lp, verbose=0):
        raise NotImplementedError()


from .cuda import CudaDensePrimalNormalSolver, CudaSparsePrimalNormalSolver  # noqa: E402,F401

# the reference's class names, for code that imports them from pycllp.solvers.cl
ClDensePrimalNormalSolver = CudaDensePrimalNormalSolver
ClSparsePrimalNormalSolver = CudaSparsePrimalNormalSolver
