"""pycllp_b200 -- B200-native batched interior-point LP engine behind pycllp's API.

Drop-in for the one hot path of jetuk/pycllp: many LPs sharing a constraint matrix,
solved by the primal normal-equations path-following method (reference:
``pycllp/solvers/cl.py`` + ``pycllp/cl/*.cl``).  ``pycllp_b200.lp`` restates the problem
containers, ``pycllp_b200.solvers`` the plugin registry with the two solver names
``cl_dense_primal_normal`` / ``cl_sparse_primal_normal`` now backed by hand-written
sm_100a CUDA kernels through the C ABI in ``include/pycllp_b200.h``.

``install_as_pycllp()`` aliases this package as ``pycllp`` so that scripts written
against the reference (``from pycllp.lp import StandardLP``;
``from pycllp.solvers import solver_registry``) run unchanged.
"""
import sys

from . import lp            # noqa: F401
from . import solvers       # noqa: F401
from .solvers import solver_registry  # noqa: F401

__version__ = "0.1.0"

# MPS section / bound-type constants of the reference package (pycllp/__init__.py:1-21)
HEADER, NAME, ROWS, COLS, RHS, RNGS, BNDS, QUADS, END = range(9)
UNSET, PRIMAL, DUAL = 0, 1, 2
FINITE, INFINITE, UNCONST = 0x1, 0x2, 0x4
FREEVAR, BDD_BELOW, BDD_ABOVE, BOUNDED = 0x1, 0x2, 0x4, 0x8


def install_as_pycllp():
    """Register this package under the name ``pycllp`` (and its lp/solvers submodules)."""
    me = sys.modules[__name__]
    sys.modules.setdefault("pycllp", me)
    sys.modules.setdefault("pycllp.lp", lp)
    sys.modules.setdefault("pycllp.solvers", solvers)
    sys.modules.setdefault("pycllp.solvers.cl", solvers.cuda)
    return me
