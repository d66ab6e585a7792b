"""PCD-fieldsplit-preconditioned GMRES -- drop-in for fenapack/field_split.py.

``PCDKSP`` keeps the reference interface (``PCDKSP(comm)``, ``setOperators``,
``setOptionsPrefix`` before ``init_pcd`` only, ``setFromOptions``,
``init_pcd(pcd_assembler, pcd_pc_class=None)`` exactly once, ``solve(b, x)``)
and the fixed configuration of the reference constructor (field_split.py:46-57:
GMRES, right preconditioning, FIELDSPLIT / SCHUR / UPPER / USER).  In the
reference those are PETSc objects wired together on the host; here the whole
Krylov loop is device resident: ``solve`` makes ONE call into the C ABI
(``fnp_solve_monolithic``), so only ``b`` and ``x`` cross PCIe per linear solve
("PCIe").

With real petsc4py a maintainer may instead keep PETSc's own KSP/PCFIELDSPLIT and
attach ``PCDPC_BRM1/2`` as the python context of the "p" sub-PC -- the
per-apply compatibility mode described in INTEGRATION.md.
"""
from __future__ import annotations

import importlib
import os

import numpy as np

from . import capi
from ._backend import PETSc, mat_state
from ._comm import HostComm
from .field_split_backend import PCDInterface
from .preconditioners import BasePCDPC, PCDPC_BRM1
from .utils import allow_only_one_call

_OUTER_KEYS = ("ksp_type", "ksp_gmres_restart", "ksp_rtol", "ksp_atol", "ksp_max_it")
_U_KEYS = ("ksp_type", "ksp_max_it", "pc_type", "pc_hypre_type", "pc_amg_threshold", "pc_amg_levels",
           "pc_amg_coarse_size", "pc_amg_smooth_steps", "pc_amg_eig_ratio", "pc_amg_coarse_drop", "pc_amg_prolongator_truncation", "pc_amg_replicate_size", "pc_amg_lag",
           "pc_amg_refresh")


def dofmap_dofs_is(dofmap, comm=None):
    """Index set of the dofs owned by a sub-space (reference _field_split_utils.py:39-50):
    ``dofmap.dofs()`` lists the OWNED dofs of this rank in the global monolithic numbering."""
    return PETSc.IS(np.asarray(dofmap.dofs(), dtype=np.int64), comm)


class _SubPC(object):
    """Stand-in for the "p" sub-PC that PETSc's fieldsplit would own."""

    def __init__(self, comm, prefix):
        self.comm = comm
        self._prefix = prefix
        self._ctx = None

    def getOptionsPrefix(self):
        return self._prefix

    def setPythonContext(self, ctx):
        self._ctx = ctx
        ctx.create(self)

    def getPythonContext(self):
        return self._ctx


class PCDKSP(object):
    """GMRES with right fieldsplit preconditioning using upper Schur factorization
    and PCD Schur complement approximation, solved on the GPU."""

    def __init__(self, comm=None, device=None):
        """``comm``: the communicator of the solver (reference field_split.py:46): every rank owns a
        contiguous range of the monolithic numbering, drives one GPU, and the ranks' device contexts
        are joined through NCCL (``fnp_create_dist``).  ``device``: CUDA device of this rank (default:
        ``LOCAL_RANK`` of the launcher, else the rank)."""
        self.comm = comm if comm is not None else PETSc.COMM_WORLD
        self._hc = HostComm(self.comm)
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", self._hc.rank)) if self._hc.size > 1 else 0
        self._device = device
        self._prefix = ""
        self._A = self._P = None
        self._ctx = None
        self._outer_opts = {"ksp_type": "gmres"}       # PETSc.KSP.Type.GMRES, field_split.py:52
        self._u_opts = {}
        self._its = 0
        self._rnorm = 0.0
        self._reason = 0
        self._state = None
        self._setup_done = False

    # -- configuration -----------------------------------------------------------
    def setOptionsPrefix(self, prefix):
        self._prefix = prefix or ""

    def getOptionsPrefix(self):
        return self._prefix

    def _forbid_setOptionsPrefix(self, prefix):
        raise RuntimeError("Options prefix cannot be set now. Set it before init_pcd.")

    def setOperators(self, A, P=None):
        self._A, self._P = A, (P if P is not None else A)

    def getOperators(self):
        return self._A, self._P

    def setTolerances(self, rtol=None, atol=None, max_it=None):
        if rtol is not None:
            self._outer_opts["ksp_rtol"] = rtol
        if atol is not None:
            self._outer_opts["ksp_atol"] = atol
        if max_it is not None:
            self._outer_opts["ksp_max_it"] = max_it
        if self._ctx is not None:
            self._ctx.set_options(self._outer_opts)

    def setFromOptions(self):
        opts = PETSc.Options(self._prefix)
        for key in _OUTER_KEYS:
            val = opts.getString(key, None)
            if val is not None:
                self._outer_opts[key] = val
        # library tuning knobs (<prefix>fnp_*: not PETSc names, include/fenapack_cuda.h) pass through
        for name, val in PETSc.Options().getAll().items():
            name = name.lstrip("-")
            if name.startswith(self._prefix + "fnp_"):
                self._outer_opts[name[len(self._prefix):]] = val
        uopts = PETSc.Options(self._prefix + "fieldsplit_u_")
        for key in _U_KEYS:
            val = uopts.getString(key, None)
            if val is not None:
                if key == "pc_type" and val in ("lu", "cholesky"):
                    raise RuntimeError("fieldsplit_u_pc_type=%s: sparse direct solves are not provided on the "
                                       "device; use richardson + amg (demo_navier-stokes-pcd.py:153-156)" % val)
                self._u_opts["fieldsplit_u_" + key] = val
        if self._ctx is not None:
            self._ctx.set_options(self._outer_opts)
            self._ctx.set_options(self._u_opts)

    # -- two-phase initialisation (reference field_split.py:61-144) ---------------
    @allow_only_one_call
    def init_pcd(self, pcd_assembler, pcd_pc_class=None):
        """Initialize from a ``PCDAssembler``.  Needs to be called after
        ``setOperators``; calls ``setFromOptions`` for all sub-solvers."""
        if self._A is None:
            raise RuntimeError("init_pcd: setOperators must be called first")
        V = pcd_assembler.function_space()
        is0 = dofmap_dofs_is(V.sub(0).dofmap(), self.comm)
        is1 = dofmap_dofs_is(V.sub(1).dofmap(), self.comm)
        self.is_u, self.is_p = is0, is1
        # From now on forbid setting options prefix
        self.setOptionsPrefix = self._forbid_setOptionsPrefix
        self.setFromOptions()

        # device context owning the whole block-triangular apply; one per rank, joined through
        # NCCL when the communicator has several ranks (the unique id travels over the caller's
        # communicator, as the reference's sub-solvers inherit ``comm``, field_split.py:75-77)
        hc = self._hc
        if hc.size > 1:
            uid = hc.bcast(capi.nccl_unique_id() if hc.rank == 0 else None, root=0)
            ctx = capi.Context(self._device, nccl_id=uid, rank=hc.rank, nranks=hc.size)
        else:
            ctx = capi.Context(self._device)
        # ownership ranges of the two split numberings: position in the owned index set plus the
        # exclusive scan of the local sizes (SubfieldBC.h:138-140)
        n_u, n_p = is0.getLocalSize(), is1.getLocalSize()
        u_begin, p_begin = hc.exscan(n_u), hc.exscan(n_p)
        ctx.set_layout(n_u, n_p, u_begin, hc.allreduce(n_u), p_begin, hc.allreduce(n_p))
        rstart = self._A.getOwnershipRange()[0] if hasattr(self._A, "getOwnershipRange") else 0
        if np.any(is0.getIndices() < rstart) or np.any(is1.getIndices() < rstart) or \
                is0.getLocalSize() + is1.getLocalSize() != self._local_size():
            raise RuntimeError("PCDKSP.init_pcd: the sub-space index sets do not partition this rank's rows of the operator")
        ctx.set_index_sets(is0.getIndices() - rstart, is1.getIndices() - rstart)
        self._ctx = ctx

        # PCD PC class: option > argument > default (field_split.py:109-124)
        pcd_pc_prefix = self._prefix + "fieldsplit_p_"
        sub_pc = _SubPC(self.comm, pcd_pc_prefix)
        name = PETSc.Options(pcd_pc_prefix).getString("pc_python_type", "")
        if name:
            mod, _, cls = name.rpartition(".")
            if mod == "fenapack":
                mod = "fenapack_b200"
            pcd_pc = getattr(importlib.import_module(mod), cls)()
        elif pcd_pc_class is not None:
            pcd_pc = pcd_pc_class()
        else:
            pcd_pc = PCDPC_BRM1()
        if not isinstance(pcd_pc, BasePCDPC):
            raise TypeError("PCD PC class must derive from BasePCDPC")
        sub_pc.setPythonContext(pcd_pc)
        pcd_pc._attach(ctx)
        pcd_pc.setFromOptions(sub_pc)
        ctx.set_options(self._outer_opts)
        ctx.set_options(self._u_opts)
        self._sub_pc, self._pcd_pc = sub_pc, pcd_pc

        pcd_interface = PCDInterface(pcd_assembler, self._A, is0, is1, deep_submats=True, device=self._device)
        try:
            pcd_pc.init_pcd(pcd_interface)
        except Exception:
            print("Initialization of PCD PC from PCDAssembler failed!")
            print("Maybe wrong PCD PC class or PCDAssembler instance.")
            raise
        self._setup()

    # -- set-up / value refresh -------------------------------
    def _local_size(self):
        A = self._A
        if hasattr(A, "getLocalSize"):
            return A.getLocalSize()[0]
        return A.getSize()[0]

    def _blocks(self):
        """The blocks PCFIELDSPLIT works with.  The Krylov MatMult uses the system matrix A in full
        (A00, A01, A10 and, for pressure-stabilised discretisations, A11); the triangular apply cuts
        its blocks from the preconditioning matrix P (PETSc's default useAmat = false): P00 for the
        velocity solve and P01 for the coupling -- uploaded separately only when they differ from A's."""
        A, P = self._A, self._P
        out = {capi.MAT_A00: A.createSubMatrix(self.is_u, self.is_u),
               capi.MAT_A01: A.createSubMatrix(self.is_u, self.is_p),
               capi.MAT_A10: A.createSubMatrix(self.is_p, self.is_u)}
        a11 = A.createSubMatrix(self.is_p, self.is_p)
        a11_nonzero = bool(np.any(a11.getValuesCSR()[2] != 0.0))
        if self._hc.allreduce(1 if a11_nonzero else 0) or getattr(self, "_has_a11", False):
            out[capi.MAT_A11] = a11
            self._has_a11 = True
        if P is not A:
            out[capi.MAT_P00] = P.createSubMatrix(self.is_u, self.is_u)
            p01 = P.createSubMatrix(self.is_u, self.is_p)
            ia, ja, va = out[capi.MAT_A01].getValuesCSR()
            ip, jp, vp = p01.getValuesCSR()
            differs = not (np.array_equal(ia, ip) and np.array_equal(ja, jp) and np.array_equal(va, vp))
            if self._hc.allreduce(1 if differs else 0) or getattr(self, "_has_p01", False):
                out[capi.MAT_P01] = p01
                self._has_p01 = True
        return out

    def _operator_state(self):
        return (mat_state(self._A), mat_state(self._P))

    def _setup(self):
        """PCSetUp chain: (re-)extract the blocks of the in-place re-assembled operators -- pattern
        once, values whenever they changed -- then the python PC's setUp.  A00 / P00 change with
        every Newton or time step; the coupling blocks normally do not, but are compared and
        re-uploaded when they do (e.g. SUPG pressure terms), never silently kept."""
        ctx = self._ctx
        first = not self._setup_done
        if first:
            self._uploaded = {}
        for which, mat in self._blocks().items():
            indptr, indices, data = mat.getValuesCSR()
            if which not in self._uploaded:
                if not first:
                    raise RuntimeError("PCDKSP: block %d of the operators appeared after init_pcd "
                                       "(its pattern was empty at the first set-up)" % which)
                ctx.set_pattern(which, indptr, indices)
                ctx.set_values(which, data)
                self._uploaded[which] = None if which in (capi.MAT_A00, capi.MAT_P00) else np.array(data, copy=True)
            elif which in (capi.MAT_A00, capi.MAT_P00):
                ctx.set_values(which, data)
            else:
                changed = self._hc.allreduce(0 if np.array_equal(self._uploaded[which], data) else 1)
                if changed:
                    ctx.set_values(which, data)
                    self._uploaded[which] = np.array(data, copy=True)
        self._pcd_pc.setUp(self._sub_pc)
        ctx.setup()
        self._state = self._operator_state()
        self._setup_done = True

    # -- solve --------------------------------------------------------------------
    def solve(self, b, x):
        if self._ctx is None:
            raise RuntimeError("PCDKSP.solve: init_pcd has not been called")
        if self._operator_state() != self._state:
            self._setup()
        ba = b.getArray() if hasattr(b, "getArray") else np.asarray(b)
        sol, its, rn, nap = self._ctx.solve_monolithic(ba)
        xa = x.getArray() if hasattr(x, "getArray") else x
        xa[:] = sol
        self._its, self._rnorm = its, rn
        self._reason = self._ctx.converged_reason()       # KSPConvergedReason from the library (rtol 2, atol 3, its -3)
        return its

    def getIterationNumber(self):
        return self._its

    def getResidualNorm(self):
        return self._rnorm

    def getConvergedReason(self):
        return self._reason

    def getConvergenceHistory(self):
        return self._ctx.residual_history()

    def device_context(self):
        return self._ctx


class PCDKSPPython(object):
    """``KSPPYTHON`` context around ``PCDKSP``: the entry point for a PETSc driver that OWNS its KSP
    -- a SNES (defcon's ``SNUFLSolver``, reference demo/defcon/navier-stokes.py:252-280, which in the
    reference replaces ``solver.snes.ksp`` by a ``PCDKSP``) or a TS.  With petsc4py::

        ksp = snes.ksp
        ksp.setType(PETSc.KSP.Type.PYTHON)
        ksp.setPythonContext(PCDKSPPython(pcd_assembler))      # or -ksp_python_type with set_assembler()

    PETSc then calls ``create / setFromOptions / setUp / solve`` (petsc4py's python-KSP protocol); the
    operators are the ones the driver keeps re-assembling in place, so every ``setUp`` after a Jacobian
    update triggers the value refresh, exactly as ``PCDNewtonSolver`` does through ``PCDKrylovSolver``."""

    def __init__(self, pcd_assembler=None, pcd_pc_class=None, device=None, ksp_factory=None):
        self._assembler, self._pc_class, self._device = pcd_assembler, pcd_pc_class, device
        self._factory = ksp_factory or PCDKSP
        self._inner = None

    def set_assembler(self, pcd_assembler, pcd_pc_class=None):
        self._assembler, self._pc_class = pcd_assembler, pcd_pc_class

    def inner(self):
        return self._inner

    # -- python-KSP protocol (called by PETSc) -----------------------------------
    def create(self, ksp):
        self._inner = self._factory(comm=getattr(ksp, "comm", None), device=self._device)

    def setFromOptions(self, ksp):
        if self._inner._ctx is None:                       # the prefix is frozen by init_pcd
            self._inner.setOptionsPrefix(ksp.getOptionsPrefix() or "")
        self._inner.setFromOptions()

    def setUp(self, ksp):
        A, P = ksp.getOperators()
        self._inner.setOperators(A, P)
        if self._inner._ctx is None:
            if self._assembler is None:
                raise RuntimeError("PCDKSPPython: no PCDAssembler given (constructor argument or set_assembler)")
            self._inner.init_pcd(self._assembler, self._pc_class)

    def solve(self, ksp, b, x):
        try:
            rtol, atol, _, max_it = ksp.getTolerances()
            self._inner.setTolerances(rtol=rtol, atol=atol, max_it=max_it)
        except AttributeError:
            pass
        its = self._inner.solve(b, x)
        for name, val in (("setIterationNumber", its), ("setResidualNorm", self._inner.getResidualNorm()),
                          ("setConvergedReason", self._inner.getConvergedReason())):
            if hasattr(ksp, name):
                getattr(ksp, name)(val)

    def view(self, ksp, viewer=None):
        print("PCDKSPPython: PCD-preconditioned (F)GMRES on the device, %d iterations in the last solve"
              % (self._inner.getIterationNumber() if self._inner is not None else 0))


class PCDKrylovSolver(object):
    """Counterpart of the reference's ``dolfin.PETScKrylovSolver`` subclass
    (field_split.py:153-187)."""

    def __init__(self, comm=None, device=None):
        self._ksp = PCDKSP(comm=comm, device=device)
        self.parameters = {"relative_tolerance": 1e-6, "absolute_tolerance": 1e-50, "maximum_iterations": 10000,
                           "error_on_nonconvergence": True}

    def init_pcd(self, pcd_assembler, pcd_pc_class=None):
        self._ksp.init_pcd(pcd_assembler, pcd_pc_class=pcd_pc_class)

    def ksp(self):
        return self._ksp

    def set_options_prefix(self, prefix):
        self._ksp.setOptionsPrefix(prefix)

    def get_options_prefix(self):
        return self._ksp.getOptionsPrefix()

    def set_from_options(self):
        self._ksp.setFromOptions()

    def set_operators(self, A, P):
        self._ksp.setOperators(A, P)

    def solve(self, x, b):
        """dolfin.PETScKrylovSolver.solve(x, b) argument order."""
        self._ksp.setTolerances(rtol=self.parameters["relative_tolerance"],
                                atol=self.parameters["absolute_tolerance"],
                                max_it=self.parameters["maximum_iterations"])
        its = self._ksp.solve(b, x)
        if self._ksp.getConvergedReason() < 0 and self.parameters["error_on_nonconvergence"]:
            raise RuntimeError("PCDKrylovSolver: Krylov solver did not converge in %d iterations" % its)
        return its
