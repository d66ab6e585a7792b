"""PCD Schur-complement preconditioners as petsc4py "python" PC contexts, backed
by libfenapack_cuda -- the drop-in for fenapack/preconditioners.py.

Same classes, same protocol (``create / setFromOptions / setUp / apply`` called
by PETSc, ``init_pcd`` called once by PCDKSP), same option prefixes
(``<pc prefix>PCD_Ap_``, ``<pc prefix>PCD_Mp_``, reference
preconditioners.py:28-34).  What changed is where the arithmetic runs: the
reference issues seven petsc4py calls per apply (copy, BC scatter, two KSP
solves, MatMult, axpy, scale; preconditioners.py:128-135) -- here one call into
the C ABI (``fnp_schur_apply``) runs the whole sequence on the GPU with the
vector updates fused into the SpMV kernels.

Differences a maintainer has to know:
  * the reference's default inner KSPs are PREONLY + Cholesky through a CPU
    sparse direct package (preconditioners.py:43-49).  Sparse direct solves are
    not provided on the device; the defaults here are the reference's
    "iterative" set-up (demo_navier-stokes-pcd.py:157-165) and a request for
    ``preonly`` + ``cholesky``/``lu`` is rejected loudly.
  * ``hypre``/``boomeramg`` select the library's smoothed-aggregation V-cycle.
"""
from __future__ import annotations

import numpy as np

from . import capi
from ._backend import PETSc

_INNER_KEYS = ("ksp_type", "ksp_max_it", "ksp_rtol", "pc_type", "pc_hypre_type",
               "ksp_chebyshev_eigenvalues", "pc_amg_threshold", "pc_amg_levels",
               "pc_amg_coarse_size", "pc_amg_smooth_steps", "pc_amg_eig_ratio", "pc_amg_coarse_drop", "pc_amg_prolongator_truncation", "pc_amg_replicate_size", "pc_amg_lag",
               "pc_amg_refresh")


def _vec_array(v, readonly=False):
    try:
        return v.getArray(readonly=readonly)
    except TypeError:
        return v.getArray()


def _mat_csr(mat):
    indptr, indices, data = mat.getValuesCSR()
    return (np.ascontiguousarray(indptr, dtype=np.int32), np.ascontiguousarray(indices, dtype=np.int32),
            np.ascontiguousarray(data, dtype=np.float64))


class BasePCDPC(object):
    """Base python context of the PCD preconditioners (reference
    preconditioners.py:25-85)."""

    variant = None   # "BRM1" | "BRM2"

    # -- python-PC protocol ------------------------------------------------
    def create(self, pc):
        prefix = pc.getOptionsPrefix() or ""
        self._pc_prefix = prefix
        self._prefix_Ap = prefix + "PCD_Ap_"
        self._prefix_Mp = prefix + "PCD_Mp_"
        self._device_opts = {}
        self._ctx = None              # device context (own, Schur-only) or the one PCDKSP attaches
        self._owns_ctx = True
        self._patterns_set = False

    def setFromOptions(self, pc):
        """Collect ``<prefix>PCD_Ap_*`` / ``<prefix>PCD_Mp_*`` from the PETSc options
        database (reference :37-39 forwards them to the two inner KSPs)."""
        for sub, petsc_prefix in (("fieldsplit_p_PCD_Ap_", self._prefix_Ap), ("fieldsplit_p_PCD_Mp_", self._prefix_Mp)):
            opts = PETSc.Options(petsc_prefix)
            for key in _INNER_KEYS:
                val = opts.getString(key, None)
                if val is not None:
                    self._device_opts[sub + key] = val
        self._check_inner_solver_options()
        if self._ctx is not None:
            self._ctx.set_options(self._device_opts)

    def _check_inner_solver_options(self):
        for name, val in self._device_opts.items():
            if name.endswith("pc_type") and val in ("cholesky", "lu", "icc", "ilu"):
                raise RuntimeError(
                    f"{name}={val}: sparse direct/incomplete factorisations are CPU-only in the reference "
                    "and are not provided by libfenapack_cuda; use the iterative set-up "
                    "(richardson/cg + amg, chebyshev + jacobi)")

    def _variant_name(self):
        return "fenapack.PCDPC_" + self.variant

    def init_pcd(self, pcd_interface):
        """Initialize by PCDInterface instance (once; reference :64-68)."""
        if hasattr(self, "interface"):
            raise RuntimeError("Reinitialization of PCDPC not allowed")
        self.interface = pcd_interface

    def _attach(self, ctx):
        """Used by PCDKSP in full-device mode: share its context instead of a Schur-only one."""
        self._ctx = ctx
        self._owns_ctx = False
        ctx.set_options(self._device_opts)

    def _ensure_ctx(self, n_p):
        if self._ctx is None:
            if getattr(getattr(self.interface.is_p, "comm", None), "size", 1) > 1:
                raise RuntimeError("stand-alone python-PC mode is single rank; multi-rank runs go through PCDKSP, "
                                   "which attaches its distributed device context")
            self._ctx = capi.Context(self.interface.device if hasattr(self.interface, "device") else 0)
            self._ctx.set_options(self._device_opts)
            self._ctx.set_layout(0, n_p)

    def setUp(self, pc):
        """Per PCSetUp (reference :71-85): Mp and Ap once, Kp whenever its form is
        not constant (value-only refresh of the device copy), BC index list once."""
        itf = self.interface
        Mp = itf.setup_mat_Mp(mat=getattr(self, "mat_Mp", None))
        Ap = itf.setup_mat_Ap(mat=getattr(self, "mat_Ap", None))
        Kp = itf.setup_mat_Kp(mat=getattr(self, "mat_Kp", None))
        if Mp is not None:
            self.mat_Mp = Mp
        if Ap is not None:
            self.mat_Ap = Ap
        if Kp is not None:        # updated only if not constant
            self.mat_Kp = Kp
            self.mat_Kp.setOptionsPrefix(self._pc_prefix + "PCD_Kp_")
        n_p = self.mat_Mp.getLocalSize()[0] if hasattr(self.mat_Mp, "getLocalSize") else self.mat_Mp.getSize()[0]
        self._ensure_ctx(n_p)
        ctx = self._ctx
        ctx.set_option("fieldsplit_p_pc_python_type", self._variant_name())
        for which, mat, fresh in ((capi.MAT_MP, self.mat_Mp, Mp), (capi.MAT_AP, self.mat_Ap, Ap),
                                  (capi.MAT_KP, self.mat_Kp, Kp)):
            if fresh is None:
                continue
            rp, ci, va = _mat_csr(mat)
            if not self._patterns_set or which not in ctx._shapes:
                ctx.set_pattern(which, rp, ci)
            ctx.set_values(which, va)
        if not self._patterns_set:
            idx, vals = itf.pcd_bc_indices()
            ctx.set_bc(idx, vals)
            self._patterns_set = True
        # Fetch bcs apply function (kept for API compatibility; the device applies them itself)
        self.bcs_applier = itf.apply_pcd_bcs
        if self._owns_ctx:
            ctx.setup()

    def apply(self, pc, x, y):
        xa = _vec_array(x, readonly=True)
        ya = _vec_array(y)
        if xa is ya or np.shares_memory(xa, ya):
            raise ValueError("PCD apply: x and y must be different vectors")
        self._ctx.schur_apply(xa, out=ya)

    # -- reference helper kept for compatibility -----------------------------
    def get_work_vecs(self, v, num):
        """Return ``num`` work vecs initially duplicated from v (reference :52-61);
        the device path needs none, but callers may rely on the contract."""
        try:
            vecs = self._work_vecs
            assert len(vecs) == num
        except AttributeError:
            self._work_vecs = vecs = tuple(v.duplicate() for i in range(num))
        except AssertionError:
            raise ValueError("Changing number of work vecs not allowed")
        return vecs


class PCDPC_BRM1(BasePCDPC):
    r"""``y = -Mp^{-1} (x + Kp Ap^{-1} bc(x))`` -- reference preconditioners.py:88-135.
    The identity term is kept separate from ``Kp Ap^{-1}`` exactly as the
    reference insists (:111-119); the Dirichlet values are inserted into the
    right-hand side of the Laplacian solve only."""
    variant = "BRM1"


class PCDPC_BRM2(BasePCDPC):
    r"""``y = -(I + Ap^{-1} bc(Kp .)) Mp^{-1} x`` -- reference preconditioners.py:139-169."""
    variant = "BRM2"


class BasePCDRPC(BasePCDPC):
    """Base python context of the pressure convection diffusion reaction (PCDR)
    preconditioners (reference preconditioners.py:173-208): one more inner solver,
    ``Rp = B diag(Mu)^-1 B^T`` with prefix ``<pc prefix>PCD_Rp_``, built by the library from
    the discrete pressure gradient ``Bt`` and the diagonal of the velocity mass matrix."""

    def create(self, pc):
        super(BasePCDRPC, self).create(pc)
        self._prefix_Rp = (pc.getOptionsPrefix() or "") + "PCD_Rp_"

    def setFromOptions(self, pc):
        opts = PETSc.Options(self._prefix_Rp)
        for key in _INNER_KEYS:
            val = opts.getString(key, None)
            if val is not None:
                self._device_opts["fieldsplit_p_PCD_Rp_" + key] = val
        super(BasePCDRPC, self).setFromOptions(pc)

    def _variant_name(self):
        return "fenapack.PCDRPC_" + self.variant

    def _ensure_ctx(self, n_p):
        if self._ctx is None:
            # Schur-only context that also holds Bt (u rows x p columns) for Rp
            n_u = self.mat_Bt.getSize()[0]
            self._ctx = capi.Context(self.interface.device if hasattr(self.interface, "device") else 0)
            self._ctx.set_options(self._device_opts)
            self._ctx.set_layout(n_u, n_p)

    def setUp(self, pc):
        itf = self.interface
        # velocity mass matrix and discrete pressure gradient (reference :199-206)
        Mu = itf.setup_mat_Mu(mat=getattr(self, "mat_Mu", None))
        if Mu is not None:
            self.mat_Mu = Mu
            self.mat_Mu.setOptionsPrefix(self._pc_prefix + "PCD_Mu_")
        Bt = itf.setup_mat_Bt(mat=getattr(self, "mat_Bt", None))
        if Bt is not None:
            self.mat_Bt = Bt
            self.mat_Bt.setOptionsPrefix(self._pc_prefix + "PCD_Bt_")
        owns = self._owns_ctx
        self._owns_ctx = False                      # the base class must not call setup() yet
        try:
            super(BasePCDRPC, self).setUp(pc)
        finally:
            self._owns_ctx = owns
        ctx = self._ctx
        if Mu is not None:
            ctx.set_mu_diag(self.mat_Mu.csr.diagonal() if hasattr(self.mat_Mu, "csr") else self.mat_Mu.getDiagonal().getArray())
        if Bt is not None and owns:                 # PCDKSP uploads A01 itself in full-device mode
            rp, ci, va = _mat_csr(self.mat_Bt)
            if capi.MAT_A01 not in ctx._shapes:
                ctx.set_pattern(capi.MAT_A01, rp, ci)
            ctx.set_values(capi.MAT_A01, va)
        if owns:
            ctx.setup()


class PCDRPC_BRM1(BasePCDRPC):
    r"""``y = -Rp^{-1} x - Mp^{-1} (I + Kp Ap^{-1}) x`` -- reference preconditioners.py:212-262."""
    variant = "BRM1"


class PCDRPC_BRM2(BasePCDRPC):
    r"""``y = -Rp^{-1} x - (I + Ap^{-1} Kp) Mp^{-1} x`` -- reference preconditioners.py:266-298."""
    variant = "BRM2"
