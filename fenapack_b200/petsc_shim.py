"""Duck-typed stand-in for the slice of petsc4py that FENaPack's hot path touches
(SURVEY.md appendix A).  Used ONLY when petsc4py is not importable (it is not in
this image), so that the python-PC protocol, the options-prefix handling and the
value-refresh logic of the drop-in classes can be exercised.  It performs no
linear algebra of the hot path: Mat.mult / Vec.axpy here are host conveniences
for tests, the PCD apply itself always goes through libfenapack_cuda.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


class Options:
    """PETSc options database: one global table, prefix-aware accessors."""
    _db: dict = {}

    def __init__(self, prefix=None):
        self.prefix = prefix or ""

    def _key(self, name):
        return self.prefix + name.lstrip("-")

    def setValue(self, name, value):
        Options._db[self._key(name)] = "" if value is None else str(value)

    def __setitem__(self, name, value):
        self.setValue(name, value)

    def hasName(self, name):
        return self._key(name) in Options._db

    def getString(self, name, default=None):
        return Options._db.get(self._key(name), default)

    def getInt(self, name, default=None):
        v = Options._db.get(self._key(name))
        return default if v is None else int(v)

    def getReal(self, name, default=None):
        v = Options._db.get(self._key(name))
        return default if v is None else float(v)

    def delValue(self, name):
        Options._db.pop(self._key(name), None)

    def getAll(self):
        return dict(Options._db)

    @classmethod
    def clear(cls):
        cls._db.clear()


class Comm:
    size, rank = 1, 0

    def tompi4py(self):
        return self


COMM_WORLD = Comm()


class Vec:
    def __init__(self, array=None, comm=None):
        self.array = None if array is None else np.ascontiguousarray(array, dtype=np.float64)
        self.comm = comm or COMM_WORLD

    @classmethod
    def createWithArray(cls, array, comm=None):
        return cls(array, comm)

    def getArray(self, readonly=False):
        return self.array

    def getSize(self):
        return self.array.size

    def getLocalSize(self):
        return self.array.size

    def duplicate(self):
        return Vec(np.zeros_like(self.array), self.comm)

    def copy(self, result=None):
        if result is None:
            return Vec(self.array.copy(), self.comm)
        result.array[:] = self.array
        return result

    def axpy(self, alpha, x):
        self.array += alpha * x.array

    def scale(self, alpha):
        self.array *= alpha

    def set(self, value):
        self.array[:] = value

    def norm(self):
        return float(np.linalg.norm(self.array))


class IS:
    def __init__(self, indices, comm=None):
        self.indices = np.ascontiguousarray(indices, dtype=np.int64)
        self.comm = comm or COMM_WORLD

    def getIndices(self):
        return self.indices

    def getSize(self):
        return self.indices.size


class Mat:
    class Option:
        SPD = "spd"

    def __init__(self, csr=None, comm=None):
        self.comm = comm or COMM_WORLD
        self._prefix = None
        self._opts = {}
        self.state = 0
        self.csr = None
        if csr is not None:
            self.set_csr(csr)

    @property
    def type(self):
        return None if self.csr is None else "seqaij"

    def set_csr(self, csr):
        csr = sp.csr_matrix(csr)
        csr.sort_indices()
        self.csr = csr
        self.state += 1

    def isAssembled(self):
        return self.csr is not None

    def getSize(self):
        return self.csr.shape

    def getValuesCSR(self):
        return self.csr.indptr, self.csr.indices, self.csr.data

    def mult(self, x, y):
        y.array[:] = self.csr @ x.array

    def setOptionsPrefix(self, p):
        self._prefix = p

    def getOptionsPrefix(self):
        return self._prefix

    def setOption(self, opt, flag):
        self._opts[opt] = flag

    def getDiagonal(self, result=None):
        d = self.csr.diagonal()
        if result is None:
            return Vec(d, self.comm)
        result.array[:] = d
        return result

    def getVecLeft(self):
        return Vec(np.zeros(self.csr.shape[0]), self.comm)

    def createSubMatrix(self, isrow, iscol=None, submat=None):
        """Deep sub-matrix; with ``submat`` given the existing object is refilled
        (MAT_REUSE_MATRIX: same pattern expected, new values)."""
        iscol = isrow if iscol is None else iscol
        sub = _submatrix(self.csr, isrow.getIndices(), iscol.getIndices())
        if submat is None or submat.csr is None:
            out = Mat(sub, self.comm) if submat is None else submat
            if submat is not None:
                out.set_csr(sub)
            return out
        if sub.nnz != submat.csr.nnz or not np.array_equal(sub.indices, submat.csr.indices):
            raise RuntimeError("createSubMatrix(submat=...): non-zero pattern changed")
        submat.csr.data[:] = sub.data
        submat.state += 1
        return submat


def _submatrix(A, rows, cols):
    """A[rows, cols] keeping every *stored* entry (explicit zeros included), as
    MatCreateSubMatrix does."""
    coo = A.tocoo()
    rmap = np.full(A.shape[0], -1, dtype=np.int64)
    cmap = np.full(A.shape[1], -1, dtype=np.int64)
    rmap[rows] = np.arange(len(rows))
    cmap[cols] = np.arange(len(cols))
    r, c = rmap[coo.row], cmap[coo.col]
    keep = (r >= 0) & (c >= 0)
    out = sp.coo_matrix((coo.data[keep], (r[keep], c[keep])), shape=(len(rows), len(cols))).tocsr()
    out.sort_indices()
    return out


class PC:
    """Just enough of PETSc.PC to drive a python-type context."""

    class Type:
        PYTHON = "python"

    def __init__(self, comm=None, prefix=""):
        self.comm = comm or COMM_WORLD
        self._prefix = prefix
        self._ctx = None

    def getOptionsPrefix(self):
        return self._prefix

    def setOptionsPrefix(self, p):
        self._prefix = p

    def setPythonContext(self, ctx):
        self._ctx = ctx
        ctx.create(self)

    def getPythonContext(self):
        return self._ctx

    def setFromOptions(self):
        if self._ctx is None:
            name = Options(self._prefix).getString("pc_python_type", "")
            if name:
                import importlib
                mod, _, cls = name.rpartition(".")
                self.setPythonContext(getattr(importlib.import_module(mod), cls)())
        if self._ctx is not None:
            self._ctx.setFromOptions(self)

    def setUp(self):
        self._ctx.setUp(self)

    def apply(self, x, y):
        self._ctx.apply(self, x, y)


class _PETSc:
    Options = Options
    Vec = Vec
    Mat = Mat
    IS = IS
    PC = PC
    Comm = Comm
    COMM_WORLD = COMM_WORLD


PETSc = _PETSc
