"""Problem containers for batches of LPs that share one constraint matrix.

This module restates the API surface of the reference's ``pycllp/lp.py`` (row a1
of SURVEY.md section 8): ``SparseMatrix`` (``lp.py:16-302``), ``EqualityLP``
(``lp.py:306-535``), ``StandardLP`` (``lp.py:538-567``) and ``GeneralLP``
(``lp.py:570-792``), so that code written against ``pycllp.lp`` runs unchanged:
one sparse ``A`` with the shape inferred from the largest index (``lp.py:57-68``),
``b`` of shape (nproblems, nrows), ``c`` of shape (nproblems, ncols), ``f`` of
shape (nproblems,), and ``lp.init(solver)`` / ``lp.solve(solver)`` dispatching to
the solver plugin (``lp.py:531-535``).

It is a clean re-statement, not a port: coordinates are kept as growable numpy
arrays with vectorised bulk operations (``to_equality_form`` appends all slack
columns in one step instead of one ``set_value`` scan per row, ``lp.py:96-110``),
and the latent defects listed in SURVEY.md section 2 row 1 are fixed rather than
reproduced (``update_col`` deleting a row, ``lp.py:261``; ``GeneralLP.set_num_problems``
using undefined names, ``lp.py:720-723``; ``np.neginf``, ``lp.py:684``).
"""
import numpy as np
from scipy.sparse import coo_matrix

__all__ = ["SparseMatrix", "EqualityLP", "StandardLP", "GeneralLP"]


def _per_entry_values(value, count):
    """Normalise the ``value`` argument of add_row/add_col/update_*.

    Returns an array of shape (k, count) with k == 1 (same for all problems) or
    k == nproblems; accepts a scalar, a 1-D array of length ``count`` or a 2-D
    array (problems, count) -- the three forms the reference accepts
    (``lp.py:137-147``).
    """
    v = np.asarray(value, dtype=np.float64)
    if v.ndim == 0:
        return np.full((1, count), float(v))
    if v.ndim == 1:
        if v.shape[0] != count:
            raise ValueError("Inconsistent data array provided.")
        return v.reshape(1, count)
    if v.ndim == 2:
        if v.shape[1] != count:
            raise ValueError("Inconsistent data array provided.")
        return v
    raise ValueError("Inconsistent data array provided.")


class SparseMatrix(object):
    """One sparsity structure (COO) carrying ``nproblems`` sets of values.

    Reference: ``lp.py:16-302``.  ``data`` has shape (nproblems, nnzeros).
    """

    def __init__(self, rows=None, cols=None, data=None, matrix=None):
        if matrix is not None:
            coo = matrix.tocoo()
            self._rows = np.asarray(coo.row).copy()
            self._cols = np.asarray(coo.col).copy()
            self.data = np.asarray(coo.data).reshape(1, -1).copy()
        elif data is not None:
            if not (len(rows) == len(cols) == np.shape(data)[-1]):
                raise ValueError("Arrays rows, cols and data must be the same length.")
            self._rows = np.array(rows)
            self._cols = np.array(cols)
            self.data = np.atleast_2d(np.array(data))
        else:
            self._rows = np.zeros(0, dtype=np.int64)
            self._cols = np.zeros(0, dtype=np.int64)
            self.data = np.zeros((1, 0))

    # -- shape ---------------------------------------------------------------
    @property
    def nrows(self):
        return int(self._rows.max()) + 1 if self._rows.size else 0

    @property
    def ncols(self):
        return int(self._cols.max()) + 1 if self._cols.size else 0

    @property
    def nnzeros(self):
        return int(self._rows.size)

    @property
    def nproblems(self):
        return self.data.shape[0]

    # -- single entries --------------------------------------------------------
    def _find(self, row, col):
        return np.flatnonzero((self._rows == row) & (self._cols == col))

    def _append(self, rows, cols, values):
        """Bulk-append entries; ``values`` is (k, len(rows)), k in {1, nproblems}."""
        values = np.asarray(values, dtype=self.data.dtype if self.data.size else np.float64)
        if values.shape[0] not in (1, self.nproblems):
            raise ValueError("The number of coordinate values must match the number of problems.")
        block = np.broadcast_to(values, (self.nproblems, values.shape[1]))
        self._rows = np.concatenate([self._rows, np.asarray(rows, dtype=self._rows.dtype)])
        self._cols = np.concatenate([self._cols, np.asarray(cols, dtype=self._cols.dtype)])
        self.data = np.concatenate([self.data.astype(np.result_type(self.data, block)), block], axis=1)

    def set_value(self, row, col, value):
        """Set entry (row, col) for all problems (scalar) or per problem (array)."""
        if row < 0 or col < 0:
            raise ValueError("Coordinates (i,j) must be >= 0")
        v = np.asarray(value, dtype=np.float64)
        if v.ndim >= 1 and v.size not in (1, self.nproblems):
            raise ValueError("The number of coordinate values must match the number of problems.")
        hit = self._find(row, col)
        if hit.size == 1:
            self.data[:, hit[0]] = v.reshape(-1) if v.ndim else v
        elif hit.size == 0:
            self._append([row], [col], v.reshape(-1, 1) if v.ndim else v.reshape(1, 1))
        else:
            raise ValueError("Multiple entries with the same coordinate pair. Bad things have happened!")

    def _keep(self, mask):
        self._rows = self._rows[mask]
        self._cols = self._cols[mask]
        self.data = self.data[:, mask]

    def _del_value(self, row, col):
        hit = self._find(row, col)
        if hit.size != 1:
            raise ValueError("Multiple entries with the same coordinate pair. Bad things have happened!")
        mask = np.ones(self.nnzeros, dtype=bool)
        mask[hit[0]] = False
        self._keep(mask)

    # -- rows --------------------------------------------------------------------
    def _put_many(self, rows, cols, value):
        vals = _per_entry_values(value, len(rows))
        for k, (r, c) in enumerate(zip(rows, cols)):
            self.set_value(int(r), int(c), vals[:, k] if vals.shape[0] > 1 else vals[0, k])

    def add_row(self, cols, value):
        """Append a row; returns its index (the current number of rows)."""
        row = self.nrows
        cols = list(cols)
        self._put_many([row] * len(cols), cols, value)
        return row

    def get_row(self, row):
        sel = self._rows == row
        return self._cols[sel], self.data[:, sel]

    @property
    def rows(self):
        for row in range(self.nrows):
            cols, value = self.get_row(row)
            yield row, cols, value

    def _del_row(self, row):
        self._keep(self._rows != row)

    def update_row(self, row, cols, value):
        self._del_row(row)
        cols = list(cols)
        self._put_many([row] * len(cols), cols, value)

    # -- columns -----------------------------------------------------------------
    def add_col(self, rows, value):
        """Append a column; returns its index (the current number of columns)."""
        col = self.ncols
        rows = list(rows)
        self._put_many(rows, [col] * len(rows), value)
        return col

    def get_col(self, col):
        sel = self._cols == col
        return self._rows[sel], self.data[:, sel]

    @property
    def cols(self):
        for col in range(self.ncols):
            rows, value = self.get_col(col)
            yield col, rows, value

    def _del_col(self, col):
        self._keep(self._cols != col)

    def update_col(self, col, rows, value):
        self._del_col(col)
        rows = list(rows)
        self._put_many(rows, [col] * len(rows), value)

    def set_num_problems(self, nproblems):
        """Grow the number of value sets; new problems are zero filled."""
        extra = nproblems - self.data.shape[0]
        if extra > 0:
            self.data = np.pad(self.data, ((0, extra), (0, 0)), mode="constant")

    # -- conversions ---------------------------------------------------------------
    def tocoo(self, problem=0):
        return coo_matrix((self.data[problem, :], (self._rows, self._cols)),
                          shape=(self.nrows, self.ncols))

    def tocsc(self, problem=0):
        return self.tocoo(problem).tocsc()

    def tocsr(self, problem=0):
        return self.tocoo(problem).tocsr()

    def tocsc_arrays(self):
        """(values (nproblems, nnz), row indices, column pointers) in CSC order.

        Reference: ``lp.py:289-299`` (entries keep their insertion order within a
        column, as there).
        """
        order = np.argsort(self._cols, kind="stable")
        counts = np.bincount(self._cols, minlength=self.ncols) if self.nnzeros else np.zeros(0, int)
        kA = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        return (np.ascontiguousarray(self.data[:, order]), self._rows[order].astype(np.int32), kA)

    def todense(self, problem=0):
        return self.tocoo(problem).todense()


class EqualityLP(object):
    """maximize c'x  subject to  A x = b, x >= 0, for ``nproblems`` (b, c, f) sets.

    Reference: ``lp.py:306-535``.  A single ``A`` is shared by all problems
    (``lp.py:335-336``).
    """

    def __init__(self, A=None, b=None, c=None, f=None):
        if A is None:
            self.A = SparseMatrix()
            self.b = np.zeros((1, 0))
            self.c = np.zeros((1, 0))
            self.f = np.zeros(1)
            return
        if b is None or c is None or f is None:
            raise ValueError("If A matrix is provided then b, c and f must also be provided.")
        if not isinstance(A, SparseMatrix):
            A = SparseMatrix(matrix=A)
        if A.nproblems > 1:
            raise ValueError("A matrix can only have a single problem in the current implementation.")
        self.A = A
        self.b = np.atleast_2d(np.array(b, dtype=np.float64))
        nprb = self.b.shape[0]
        c = np.array(c, dtype=np.float64)
        self.c = np.tile(c, (nprb, 1)) if c.ndim == 1 else c
        if self.c.shape[0] != nprb:
            raise ValueError("A matrix and c array do not have the same number of problems.")
        self.f = np.full(nprb, float(f)) if np.isscalar(f) else np.array(f, dtype=np.float64)

    # -- shape ---------------------------------------------------------------
    nrows = property(lambda self: self.A.nrows)
    ncols = property(lambda self: self.A.ncols)
    nnzeros = property(lambda self: self.A.nnzeros)
    nproblems = property(lambda self: self.b.shape[0])
    m = property(lambda self: self.A.nrows, doc="Number of rows (constraints)")
    n = property(lambda self: self.A.ncols, doc="Number of columns (variables)")

    # -- bounds ----------------------------------------------------------------
    def set_bound(self, row, bound):
        if row >= self.b.shape[1]:
            raise ValueError("Can not set bounds for row that does not exist.")
        self._set_bound(row, bound)

    def _set_bound(self, row, bound):
        bnd = np.asarray(bound, dtype=np.float64)
        if self.b.shape[1] == 0 and bnd.ndim > 0 and bnd.shape[0] != self.nproblems:
            self.set_num_problems(bnd.shape[0])
        grow = row + 1 - self.b.shape[1]
        if grow > 0:
            self.b = np.pad(self.b, ((0, 0), (0, grow)), mode="constant")
        self.b[:, row] = bnd

    # -- objective -----------------------------------------------------------------
    def set_objective(self, col, obj):
        if col >= self.c.shape[1]:
            raise ValueError("Can not set objective coefficient for column that does not exist.")
        self._set_objective(col, obj)

    def _set_objective(self, col, obj):
        grow = col + 1 - self.c.shape[1]
        if grow > 0:
            self.c = np.pad(self.c, ((0, 0), (0, grow)), mode="constant")
        self.c[:, col] = obj

    def _sync_objective_width(self):
        if self.c.shape[1] < self.A.ncols:
            self.c = np.pad(self.c, ((0, 0), (0, self.A.ncols - self.c.shape[1])), mode="constant")

    # -- rows / columns ----------------------------------------------------------
    def add_row(self, cols, value, bound):
        row = self.A.add_row(cols, value)
        self._set_bound(row, bound)
        self._sync_objective_width()
        return row

    def get_row(self, row):
        cols, value = self.A.get_row(row)
        return cols, value, self.b[:, row]

    @property
    def rows(self):
        for row in range(self.nrows):
            cols, value, bound = self.get_row(row)
            yield row, cols, value, bound

    def add_col(self, rows, value, obj):
        col = self.A.add_col(rows, value)
        self._set_objective(col, obj)
        return col

    def get_col(self, col):
        rows, value = self.A.get_col(col)
        return rows, value, self.c[:, col]

    @property
    def cols(self):
        for col in range(self.ncols):
            rows, value, obj = self.get_col(col)
            yield col, rows, value, obj

    def set_num_problems(self, nproblems):
        """Grow b, c, f to ``nproblems`` (zero filled). A stays shared."""
        extra = nproblems - self.b.shape[0]
        if extra > 0:
            self.b = np.pad(self.b, ((0, extra), (0, 0)), mode="constant")
            self.c = np.pad(self.c, ((0, extra), (0, 0)), mode="constant")
            self.f = np.pad(self.f, (0, extra), mode="constant")

    def remove_unbounded(self):
        """Copy as a StandardLP without the rows whose bound is infinite."""
        inf = np.isinf(self.b)
        all_inf, any_inf = inf.all(axis=0), inf.any(axis=0)
        mixed = np.flatnonzero(any_inf & ~all_inf)
        if mixed.size:
            raise ValueError("Can not remove unbounded rows. Row {} has some unbounded rounds.".format(mixed[0]))
        keep = np.flatnonzero(~all_inf[: self.nrows])
        remap = -np.ones(max(self.nrows, 1), dtype=np.int64)
        remap[keep] = np.arange(keep.size)
        sel = remap[self.A._rows] >= 0 if self.A.nnzeros else np.zeros(0, dtype=bool)
        lp = StandardLP()
        lp.A = SparseMatrix(remap[self.A._rows[sel]], self.A._cols[sel], self.A.data[:, sel]) \
            if sel.any() else SparseMatrix()
        lp.b = self.b[:, keep].copy()
        lp.c = self.c.copy()
        lp.f = self.f.copy()
        return lp

    # -- solver dispatch (lp.py:531-535) -------------------------------------------
    def init(self, solver, verbose=0):
        solver.init(self, verbose=verbose)

    def solve(self, solver, verbose=0):
        return solver.solve(self, verbose=verbose)

    # -- bulk construction helper (avoid O(m nnz) Python) ----------
    @classmethod
    def from_arrays(cls, A, b, c, f=0.0):
        """Build directly from a scipy sparse / dense ``A`` (already in equality form)."""
        if not hasattr(A, "tocoo"):
            A = coo_matrix(np.asarray(A))
        return cls(SparseMatrix(matrix=A), b, c, f)


class StandardLP(EqualityLP):
    """maximize c'x  subject to  A x <= b, x >= 0  (reference ``lp.py:538-567``)."""

    def to_equality_form(self):
        """Append one slack column per row: [A I] x = b, slack objective 0.

        Same result as the reference (``lp.py:551-567``: slack ``ncols + row`` for
        each row, value 1.0, objective 0.0) built in one vectorised append.
        """
        m, n0 = self.nrows, self.ncols
        A = SparseMatrix(self.A._rows.copy(), self.A._cols.copy(), self.A.data.copy()) \
            if self.A.nnzeros else SparseMatrix()
        if m:
            A._append(np.arange(m), n0 + np.arange(m), np.ones((1, m)))
        c = np.concatenate([self.c, np.zeros((self.c.shape[0], m))], axis=1)
        return EqualityLP(A, self.b.copy(), c, self.f.copy())


class GeneralLP(StandardLP):
    """optimize c'x + f  subject to  a <= A x <= b,  l <= x <= u  (``lp.py:570-792``)."""

    def __init__(self, A=None, b=None, c=None, a=None, l=None, u=None, f=None):
        super(GeneralLP, self).__init__(A=A, b=b, c=c, f=f)
        if A is None:
            self.a = np.zeros((1, 0))
            self.l = np.zeros((1, 0))
            self.u = np.zeros((1, 0))
            return
        nprb = self.nproblems

        def spread(v, default, shape):
            if v is None:
                return np.full(shape, default, dtype=np.float64)
            v = np.array(v, dtype=np.float64)
            return np.tile(v, (nprb, 1)) if v.ndim == 1 else v

        # The reference defaults the row LOWER bounds to +inf (lp.py:607); that is a
        # defect -- "no lower bound" is -inf.
        self.a = spread(a, -np.inf, self.b.shape)
        self.l = spread(l, 0.0, self.c.shape)
        self.u = spread(u, np.inf, self.c.shape)

    def set_bound(self, row, lower_bound, upper_bound):
        if row >= self.b.shape[1]:
            raise ValueError("Can not set bounds for row that does not exist.")
        self._set_bound(row, lower_bound, upper_bound)

    def _set_bound(self, row, lower_bound, upper_bound):
        super(GeneralLP, self)._set_bound(row, upper_bound)
        if self.a.shape[0] != self.b.shape[0]:
            self.a = np.pad(self.a, ((0, self.b.shape[0] - self.a.shape[0]), (0, 0)), mode="constant")
        grow = row + 1 - self.a.shape[1]
        if grow > 0:
            self.a = np.pad(self.a, ((0, 0), (0, grow)), mode="constant")
        self.a[:, row] = lower_bound

    def add_row(self, cols, value, lower_bound, upper_bound):
        known = set(self.A._cols.tolist())
        new_cols = [col for col in cols if col not in known]
        row = self.A.add_row(cols, value)
        self._set_bound(row, lower_bound, upper_bound)
        for col in new_cols:
            self._set_objective(col, 0.0)
            self._set_col_bounds(col)
        return row

    def get_row(self, row):
        cols, value, ub = StandardLP.get_row(self, row)
        return cols, value, self.a[:, row], ub

    def set_col_bounds(self, col, lower_bound=0.0, upper_bound=np.inf):
        if col >= self.l.shape[1]:
            raise ValueError("Can not set bounds for column that does not exist.")
        if np.any(np.isneginf(lower_bound)):
            raise ValueError("Column lower bounds can not be -inf.")
        self._set_col_bounds(col, lower_bound=lower_bound, upper_bound=upper_bound)

    def _set_col_bounds(self, col, lower_bound=0.0, upper_bound=np.inf):
        for name in ("l", "u"):
            arr = getattr(self, name)
            if arr.shape[0] != self.c.shape[0]:
                arr = np.pad(arr, ((0, self.c.shape[0] - arr.shape[0]), (0, 0)), mode="constant")
            grow = col + 1 - arr.shape[1]
            if grow > 0:
                arr = np.pad(arr, ((0, 0), (0, grow)), mode="constant")
            setattr(self, name, arr)
        self.l[:, col] = lower_bound
        self.u[:, col] = upper_bound

    def add_col(self, rows, value, obj, lower_bound=0.0, upper_bound=np.inf):
        col = self.A.add_col(rows, value)
        self._set_objective(col, obj)
        self._set_col_bounds(col, lower_bound, upper_bound)
        return col

    def set_num_problems(self, nproblems):
        extra = nproblems - self.b.shape[0]
        super(GeneralLP, self).set_num_problems(nproblems)
        if extra > 0:
            self.a = np.pad(self.a, ((0, extra), (0, 0)), mode="constant")
            self.l = np.pad(self.l, ((0, extra), (0, 0)), mode="constant")
            self.u = np.pad(self.u, ((0, extra), (0, 0)), mode="constant")

    def to_standard_form(self):
        """Return the equivalent StandardLP (``lp.py:725-792``).

        Variables are shifted by their lower bound (x <- x - l); every finite row
        lower bound becomes ``-A x <= -(a - A l)``, every finite row upper bound
        ``A x <= b - A l``, every finite variable upper bound ``x_j <= u_j - l_j``;
        rows unbounded for all problems are dropped.
        """
        if np.isneginf(self.l).any():
            raise ValueError("Lower bounds (l) contains -inf.")
        m, n = self.nrows, self.ncols
        l = self.l[:, :n]
        Acsr = self.A.tocsr() if self.A.nnzeros else None
        Al = (Acsr.dot(l.T)).T if Acsr is not None else np.zeros((self.nproblems, m))
        lo, up = self.a[:, :m] - Al, self.b[:, :m] - Al
        f = self.f + np.einsum("pj,pj->p", self.c[:, :n], l)

        rows, cols, vals, bounds = [], [], [], []
        nxt = 0
        for sign, bnd in ((-1.0, -lo), (1.0, up)):
            for r in range(m):
                sel = self.A._rows == r
                rows.append(np.full(int(sel.sum()), nxt))
                cols.append(self.A._cols[sel])
                vals.append(sign * self.A.data[:, sel])
                bounds.append(bnd[:, r])
                nxt += 1
        ub = self.u[:, :n] - l
        for j in np.flatnonzero(np.isfinite(ub).any(axis=0)):
            rows.append(np.array([nxt]))
            cols.append(np.array([j]))
            vals.append(np.ones((self.A.nproblems, 1)))
            bounds.append(ub[:, j])
            nxt += 1
        if rows:
            A = SparseMatrix(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals, axis=1))
            b = np.stack(bounds, axis=1)
        else:
            A, b = SparseMatrix(), np.zeros((self.nproblems, 0))
        lp = StandardLP()
        lp.A, lp.b, lp.c, lp.f = A, b, self.c.copy(), f
        return lp.remove_unbounded()
