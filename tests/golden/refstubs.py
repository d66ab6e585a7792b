"""Stand-ins that let the reference's OWN Python files for the hot path
(/root/reference/fenapack/preconditioners.py and field_split_backend.py) be executed in this
container, where petsc4py / DOLFIN are not installed.  Test infrastructure only: used by
tests/golden/make_reference_golden.py to produce golden vectors *with the reference's code*.

What is real and what is a stand-in in such a run:
  real      PCDPC_BRM1/2.apply, PCDRPC_BRM1/2.apply, BasePCD(R)PC.create/setUp/init_pcd/
            get_work_vecs, PCDInterface.setup_ksp / setup_ksp_Ap / setup_ksp_Mp / setup_mat_Kp /
            setup_mat_Mu / setup_mat_Bt / setup_ksp_Rp / _build_approx_Ap / apply_pcd_bcs /
            apply_bcs / _get_deep_submat -- every line of the reference's Python on this path
  stand-in  the petsc4py objects those lines call (Vec, Mat, KSP, IS: numpy/scipy float64; KSP
            PREONLY + CHOLESKY = sparse LU to machine precision, the reference's default inner
            solver, preconditioners.py:43-49), dolfin.timed/Timer/PETScMatrix/DirichletBC, the
            form assembly (matrices come from oracle/fem.py, embedded in the mixed space so the
            reference's createSubMatrix calls cut them back out), and SubfieldBC -- C++ in the
            reference (fenapack/SubfieldBC.h:92-182), restated here in Python.
"""
import importlib.util
import sys
import types

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

REF = "/root/reference/fenapack"


# ------------------------------------------------------------------ petsc4py.PETSc
class Comm:
    def tompi4py(self):
        return self


COMM = Comm()


class Vec:
    def __init__(self, array):
        self.array = np.array(array, dtype=np.float64)
        self.comm = COMM

    def duplicate(self):
        return Vec(np.zeros_like(self.array))

    def copy(self, result=None):
        if result is None:
            return Vec(self.array.copy())
        result.array[:] = self.array
        return result

    def axpy(self, alpha, x):
        self.array += alpha * x.array

    def scale(self, alpha):
        self.array *= alpha

    def reciprocal(self):
        self.array = 1.0 / self.array

    def sqrtabs(self):
        self.array = np.sqrt(np.abs(self.array))

    def setValues(self, idx, vals):
        self.array[np.asarray(idx, dtype=np.int64)] = vals

    def assemble(self):
        pass


class IS:
    def __init__(self, indices):
        self.indices = np.asarray(indices, dtype=np.int64)
        self.comm = COMM

    def getIndices(self):
        return self.indices


class Mat:
    class Option:
        SPD = "spd"

    def __init__(self, csr=None):
        self.csr = None if csr is None else sp.csr_matrix(csr)
        self.comm = COMM
        self.prefix = None
        self.options = {}

    def create(self, comm=None):
        return self

    @property
    def type(self):
        return None if self.csr is None else "seqaij"

    def isAssembled(self):
        return self.csr is not None

    def getSize(self):
        return self.csr.shape

    def setOption(self, opt, flag):
        self.options[opt] = flag

    def setOptionsPrefix(self, p):
        self.prefix = p

    def getOptionsPrefix(self):
        return self.prefix

    def createSubMatrix(self, isrow, iscol=None, submat=None):
        iscol = isrow if iscol is None else iscol
        sub = self.csr[isrow.indices][:, iscol.indices].tocsr()
        if submat is None or submat.csr is None:
            return Mat(sub)
        submat.csr = sub                     # MAT_REUSE_MATRIX: same object, new values
        return submat

    def createSubMatrixVirtual(self, mat, isrow, iscol=None):
        iscol = isrow if iscol is None else iscol
        self.csr = mat.csr[isrow.indices][:, iscol.indices].tocsr()
        return self

    def getDiagonal(self, result=None):
        d = self.csr.diagonal()
        if result is None:
            return Vec(d)
        result.array[:] = d
        return result

    def getVecLeft(self):
        return Vec(np.zeros(self.csr.shape[0]))

    def duplicate(self):
        return Mat(self.csr * 0.0)

    def copy(self, result=None):
        if result is None:
            return Mat(self.csr.copy())
        result.csr = self.csr.copy()
        return result

    def diagonalScale(self, L=None, R=None):
        if L is not None:
            self.csr = (sp.diags(L.array) @ self.csr).tocsr()
        if R is not None:
            self.csr = (self.csr @ sp.diags(R.array)).tocsr()

    def transposeMatMult(self, other, result=None):
        prod = (self.csr.T @ other.csr).tocsr()
        if result is None:
            return Mat(prod)
        result.csr = prod
        return result

    def mult(self, x, y):
        y.array[:] = self.csr @ x.array


class PC:
    class Type:
        CHOLESKY = "cholesky"
        LU = "lu"

    def __init__(self):
        self.type = None
        self.factor_solver_type = None

    def setType(self, t):
        self.type = t

    def setFactorSolverType(self, t):
        self.factor_solver_type = t


class KSP:
    class Type:
        PREONLY = "preonly"

    def __init__(self):
        self.pc = PC()
        self.type = None
        self.prefix = None
        self.ops = (Mat(), Mat())
        self._lu = None
        self.comm = COMM
        self.solves = 0

    def create(self, comm=None):
        return self

    def setType(self, t):
        self.type = t

    def setOptionsPrefix(self, p):
        self.prefix = p

    def getOptionsPrefix(self):
        return self.prefix

    def setFromOptions(self):
        pass

    def getOperators(self):
        return self.ops

    def setOperators(self, A, P=None):
        self.ops = (A, A if P is None else P)
        self._lu = None

    def setUp(self):
        assert self.type == KSP.Type.PREONLY and self.pc.type in (PC.Type.CHOLESKY, PC.Type.LU)
        self._lu = spla.splu(sp.csc_matrix(self.ops[1].csr))

    def solve(self, b, x):
        if self._lu is None:
            self.setUp()
        x.array[:] = self._lu.solve(b.array)
        self.solves += 1


class Sys:
    @staticmethod
    def getVersion():
        return (3, 12, 0)

    @staticmethod
    def getVersionInfo():
        return {"release": True}


class _PCHandle:
    """The `pc` argument PETSc passes to the python context's methods."""
    comm = COMM

    def __init__(self, prefix):
        self._prefix = prefix

    def getOptionsPrefix(self):
        return self._prefix


# ------------------------------------------------------------------ dolfin
def timed(name):
    def deco(f):
        return f
    return deco


class Timer:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class PETScMatrix:
    def __init__(self, comm=None):
        self._mat = Mat()

    def mat(self):
        return self._mat


class DirichletBC:
    """Carries what DirichletBC.get_boundary_values() returns: {mixed-space dof: value}."""

    def __init__(self, values):
        self.values = dict(values)

    def get_boundary_values(self):
        return dict(self.values)


# ------------------------------------------------------------------ fenapack._field_split_utils
class SubfieldBC:
    """Python restatement of fenapack/SubfieldBC.h:92-182 (C++, not runnable here): the BC's
    mixed-space dofs that belong to the index set are mapped to their position in the split
    vector (:145-155); apply() sets those entries (:162-182, VecSetValues INSERT + assembly)."""

    def __init__(self, bc, iset):
        pos = {int(g): k for k, g in enumerate(iset.getIndices())}
        pairs = sorted((pos[g], v) for g, v in bc.get_boundary_values().items() if g in pos)
        self.idx = np.array([p for p, _ in pairs], dtype=np.int64)
        self.val = np.array([v for _, v in pairs], dtype=np.float64)

    def apply(self, vec):
        vec.setValues(self.idx, self.val)
        vec.assemble()


# ------------------------------------------------------------------ fenapack.assembling
class _Form:
    def __init__(self, constant, phantom=False):
        self._c, self._p = constant, phantom

    def is_constant(self):
        return self._c

    def is_phantom(self):
        return self._p


class PCDAssembler:
    """Hands the reference's PCDInterface mixed-space matrices: each split block produced by
    oracle/fem.py is embedded at (is_row, is_col) of an N x N matrix, so that the reference's own
    createSubMatrix calls (field_split_backend.py:331-334) extract it again."""

    def __init__(self, n_mixed, is_u, is_p, blocks, bc_mixed):
        self.n, self.is_u, self.is_p, self.blocks = n_mixed, is_u, is_p, blocks
        self._bc = DirichletBC(bc_mixed)
        self.calls = []
        self.forms = {"ap": _Form(True), "mp": _Form(True), "kp": _Form(False), "fp": _Form(False),
                      "mu": _Form(True), "gp": _Form(True, phantom=True)}

    def _embed(self, key, rows, cols):
        B = sp.coo_matrix(self.blocks[key])
        return sp.csr_matrix((B.data, (rows[B.row], cols[B.col])), shape=(self.n, self.n))

    def _assemble(self, key, rows, cols, A):
        self.calls.append(key)
        A.mat().csr = self._embed(key, rows, cols)

    def ap(self, A):
        self._assemble("ap", self.is_p, self.is_p, A)

    def mp(self, A):
        self._assemble("mp", self.is_p, self.is_p, A)

    def kp(self, A):
        self._assemble("kp", self.is_p, self.is_p, A)

    def mu(self, A):
        self._assemble("mu", self.is_u, self.is_u, A)

    def get_pcd_form(self, key):
        return self.forms[key]

    def pcd_bcs(self):
        return [self._bc]


# ------------------------------------------------------------------ module wiring
def load_reference_modules():
    """Import the reference's preconditioners.py and field_split_backend.py by path with the
    stand-in modules above in place of dolfin / petsc4py / the rest of the fenapack package
    (whose __init__ would JIT-compile C++ against DOLFIN)."""
    saved = {k: sys.modules.get(k) for k in ("dolfin", "petsc4py", "petsc4py.PETSc", "fenapack", "fenapack.utils",
                                             "fenapack._field_split_utils", "fenapack.assembling",
                                             "fenapack.preconditioners", "fenapack.field_split_backend")}
    dolfin = types.ModuleType("dolfin")
    dolfin.timed, dolfin.Timer, dolfin.PETScMatrix, dolfin.DirichletBC = timed, Timer, PETScMatrix, DirichletBC
    petsc = types.ModuleType("petsc4py.PETSc")
    for name, obj in (("Vec", Vec), ("Mat", Mat), ("KSP", KSP), ("PC", PC), ("IS", IS), ("Sys", Sys), ("Comm", Comm)):
        setattr(petsc, name, obj)
    petsc4py = types.ModuleType("petsc4py")
    petsc4py.PETSc = petsc
    pkg = types.ModuleType("fenapack")
    pkg.__path__ = []
    utils = types.ModuleType("fenapack.utils")
    utils.get_default_factor_solver_type = lambda comm: "mumps"
    utils.pc_set_factor_solver_type = lambda pc, t: pc.setFactorSolverType(t)
    fsu = types.ModuleType("fenapack._field_split_utils")
    fsu.SubfieldBC = SubfieldBC
    asm = types.ModuleType("fenapack.assembling")
    asm.PCDAssembler = PCDAssembler
    sys.modules.update({"dolfin": dolfin, "petsc4py": petsc4py, "petsc4py.PETSc": petsc, "fenapack": pkg,
                        "fenapack.utils": utils, "fenapack._field_split_utils": fsu, "fenapack.assembling": asm})
    try:
        mods = {}
        for name in ("preconditioners", "field_split_backend"):
            spec = importlib.util.spec_from_file_location("fenapack." + name, f"{REF}/{name}.py")
            mod = importlib.util.module_from_spec(spec)
            sys.modules["fenapack." + name] = mod
            spec.loader.exec_module(mod)
            mods[name] = mod
        return mods["preconditioners"], mods["field_split_backend"]
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


# ------------------------------------------------------------------ fenapack/assembling.py
class HostTensor:
    """Matrix or vector handed to the assembler (the role of dolfin.PETScMatrix / PETScVector)."""

    def __init__(self, n=None):
        self.csr = None
        self.array = None if n is None else np.zeros(n)

    def set_csr(self, csr):
        self.csr = sp.csr_matrix(csr)


class HostBC:
    """Dirichlet condition on mixed-space dofs; serves the reference (``apply``) and the
    drop-in (``dofs`` / ``values``)."""

    def __init__(self, dofs, values):
        self._dofs = np.asarray(dofs, dtype=np.int64)
        self._vals = np.asarray(values, dtype=np.float64)

    def dofs(self):
        return self._dofs

    def values(self):
        return self._vals

    def get_boundary_values(self):
        return {int(d): float(v) for d, v in zip(self._dofs, self._vals)}

    def apply(self, tensor):
        """DirichletBC.apply(A): BC rows zeroed, unit diagonal."""
        A = sp.lil_matrix(tensor.csr)
        for d in self._dofs:
            A.rows[d], A.data[d] = [int(d)], [1.0]
        tensor.set_csr(A.tocsr())


class _SystemAssembler:
    """dolfin.SystemAssembler semantics on host callables: symmetric elimination of the BC dofs,
    rhs lifted; ``assemble(b, x)`` is the Newton variant (BC value g - x)."""

    def __init__(self, a, L, bcs):
        self.a, self.L, self.bcs = a, L, list(bcs) if bcs is not None else []

    def assemble(self, *tensors):
        dofs = np.concatenate([bc.dofs() for bc in self.bcs]) if self.bcs else np.zeros(0, dtype=np.int64)
        g = np.concatenate([bc.values() for bc in self.bcs]) if self.bcs else np.zeros(0)
        A = sp.csr_matrix(self.a())
        mats = [t for t in tensors if t.array is None]
        vecs = [t for t in tensors if t.array is not None]
        if vecs:
            b = vecs[0]
            if len(vecs) == 2:
                g = g - vecs[1].array[dofs]
            lift = np.zeros(A.shape[1])
            lift[dofs] = g
            rhs = np.array(self.L(), dtype=np.float64) - A @ lift
            rhs[dofs] = g
            b.array[:] = rhs
        if mats:
            D = sp.lil_matrix(A)
            mask = np.zeros(A.shape[0], dtype=bool)
            mask[dofs] = True
            C = sp.coo_matrix(A)
            keep = ~(mask[C.row] | mask[C.col])
            E = sp.coo_matrix((C.data[keep], (C.row[keep], C.col[keep])), shape=A.shape).tolil()
            for d in dofs:
                E[d, d] = 1.0
            del D
            mats[0].set_csr(E.tocsr())


def _assemble(form, tensor=None):
    tensor.set_csr(form())
    return tensor


def load_reference_assembling():
    """The reference's fenapack/assembling.py with dolfin.SystemAssembler / dolfin.assemble replaced
    by the host stand-ins above."""
    saved = {k: sys.modules.get(k) for k in ("dolfin", "fenapack", "fenapack.assembling")}
    dolfin = types.ModuleType("dolfin")
    dolfin.SystemAssembler, dolfin.assemble = _SystemAssembler, _assemble
    pkg = types.ModuleType("fenapack")
    pkg.__path__ = []
    sys.modules.update({"dolfin": dolfin, "fenapack": pkg})
    try:
        spec = importlib.util.spec_from_file_location("fenapack.assembling", f"{REF}/assembling.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules["fenapack.assembling"] = mod
        spec.loader.exec_module(mod)
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
