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
    """Serial communicator stand-in (what PETSc.COMM_WORLD is in a one-process run)."""
    size, rank = 1, 0

    def tompi4py(self):
        return self

    # the three collectives the drop-in layer needs (mpi4py spelling)
    def bcast(self, obj, root=0):
        return obj

    def allgather(self, obj):
        return [obj]

    def exscan(self, value):
        return 0

    def allreduce(self, value):
        return value


class TorchDistComm(Comm):
    """Communicator stand-in on ``torch.distributed`` (gloo on the CPU, nccl on GPUs): lets the
    multi-rank host logic of the drop-in classes run without MPI.  With real petsc4py the
    communicator is an ``mpi4py`` one and offers the same four calls."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist, self._group = dist, group
        self.size, self.rank = dist.get_world_size(group), dist.get_rank(group)

    def bcast(self, obj, root=0):
        box = [obj]
        self._dist.broadcast_object_list(box, src=root, group=self._group)
        return box[0]

    def allgather(self, obj):
        out = [None] * self.size
        self._dist.all_gather_object(out, obj, group=self._group)
        return out

    def exscan(self, value):
        return sum(self.allgather(value)[:self.rank])

    def allreduce(self, value):
        return sum(self.allgather(value))


COMM_WORLD = Comm()


class Vec:
    """Local part of a (row-partitioned) vector."""

    def __init__(self, array=None, comm=None):
        self.array = None if array is None else np.ascontiguousarray(array, dtype=np.float64)
        self.comm = comm or COMM_WORLD

    @classmethod
    def createWithArray(cls, array, comm=None):
        return cls(array, comm)

    def getArray(self, readonly=False):
        return self.array

    def getSize(self):
        return int(self.comm.allreduce(self.array.size))

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
        return float(np.sqrt(self.comm.allreduce(float(self.array @ self.array))))


class IS:
    """Index set: the locally owned part (global ids)."""

    def __init__(self, indices, comm=None):
        self.indices = np.ascontiguousarray(indices, dtype=np.int64)
        self.comm = comm or COMM_WORLD

    def getIndices(self):
        return self.indices

    def getLocalSize(self):
        return self.indices.size

    def getSize(self):
        return int(self.comm.allreduce(self.indices.size))


class Mat:
    """Row-partitioned sparse matrix: this rank's rows (scipy CSR, global column ids) and the
    ownership range of the row space -- the slice of MPIAIJ the drop-in layer reads."""

    class Option:
        SPD = "spd"

    def __init__(self, csr=None, comm=None, row_range=None, col_range=None):
        self.comm = comm or COMM_WORLD
        self._prefix = None
        self._opts = {}
        self.state = 0
        self.csr = None
        self.row_range = row_range
        self.col_range = col_range
        if csr is not None:
            self.set_csr(csr, row_range, col_range)

    @property
    def type(self):
        return None if self.csr is None else ("seqaij" if self.comm.size == 1 else "mpiaij")

    def set_csr(self, csr, row_range=None, col_range=None):
        """(Re)fill: in a multi-rank run ``csr`` holds the local rows and ``row_range`` /
        ``col_range`` the ownership ranges of the row / column spaces (default: what was set
        before, or the whole matrix on one rank)."""
        csr = sp.csr_matrix(csr)
        csr.sort_indices()
        self.csr = csr
        if row_range is not None:
            self.row_range = tuple(row_range)
        elif self.row_range is None or self.comm.size == 1:
            self.row_range = (0, csr.shape[0])
        if col_range is not None:
            self.col_range = tuple(col_range)
        elif self.col_range is None or self.comm.size == 1:
            self.col_range = self.row_range if csr.shape[0] == csr.shape[1] or self.comm.size > 1 else (0, csr.shape[1])
        self.state += 1

    def stateGet(self):
        return self.state

    def isAssembled(self):
        return self.csr is not None

    def getOwnershipRange(self):
        return self.row_range

    def getOwnershipRangeColumn(self):
        return self.col_range

    def getLocalSize(self):
        return (self.csr.shape[0], self.col_range[1] - self.col_range[0])

    def getSize(self):
        return (int(self.comm.allreduce(self.csr.shape[0])), self.csr.shape[1])

    def getValuesCSR(self):
        return self.csr.indptr, self.csr.indices, self.csr.data

    def mult(self, x, y):
        if self.comm.size > 1:
            raise NotImplementedError("the stand-in performs no distributed linear algebra")
        y.array[:] = self.csr @ x.array

    def setOptionsPrefix(self, p):
        self._prefix = p

    def getOptionsPrefix(self):
        return self._prefix

    def setOption(self, opt, flag):
        self._opts[opt] = flag

    def getDiagonal(self, result=None):
        d = self.csr.diagonal(k=self.row_range[0])      # local row i holds global row r0 + i
        if result is None:
            return Vec(d, self.comm)
        result.array[:] = d
        return result

    def getVecLeft(self):
        return Vec(np.zeros(self.csr.shape[0]), self.comm)

    def createSubMatrix(self, isrow, iscol=None, submat=None):
        """Deep sub-matrix; with ``submat`` given the existing object is refilled
        (MAT_REUSE_MATRIX: same pattern expected, new values).  As MatCreateSubMatrix on MPIAIJ:
        rows = this rank's part of ``isrow`` (which it must own), columns renumbered by their
        position in the concatenation of all ranks' ``iscol`` -- collective."""
        iscol = isrow if iscol is None else iscol
        r0 = self.row_range[0]
        if self.comm.size == 1:
            sub = _submatrix(self.csr, isrow.getIndices() - r0, iscol.getIndices())
            rr = (0, len(isrow.getIndices()))
            cr = (0, len(iscol.getIndices()))
        else:
            parts = self.comm.allgather(np.asarray(iscol.getIndices(), dtype=np.int64))
            allcols = np.concatenate(parts)
            off = int(sum(len(p) for p in parts[:self.comm.rank]))
            sub = _submatrix(self.csr, isrow.getIndices() - r0, allcols)
            rbeg = int(self.comm.exscan(len(isrow.getIndices())))
            rr = (rbeg, rbeg + len(isrow.getIndices()))
            cr = (off, off + len(parts[self.comm.rank]))
        if submat is None or submat.csr is None:
            out = Mat(comm=self.comm) if submat is None else submat
            out.set_csr(sub, rr, cr)
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
    TorchDistComm = TorchDistComm
    COMM_WORLD = COMM_WORLD


PETSc = _PETSc
