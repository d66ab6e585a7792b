"""Form bookkeeping of the PCD preconditioner -- drop-in for fenapack/assembling.py.

Assembly stays on the host (BASELINE.json north_star).  With DOLFIN the
``a, L, mp, ap, kp ...`` arguments are UFL forms and the work is delegated to
``dolfin.SystemAssembler`` / ``dolfin.assemble`` exactly as the reference does
(assembling.py:85-180).  Without DOLFIN (this image) every "form" is a *host
assembler callable* ``form() -> scipy.sparse matrix`` (or vector) on the full
mixed space in monolithic numbering, and a boundary condition is any object with
``dofs()`` (monolithic indices) and ``values()``.  Either way the contract seen
by PCDInterface is the same: tensors on the mixed space W.

BC semantics kept from the reference: the system / preconditioner matrix and the
right-hand side get ``bcs`` the way ``SystemAssembler`` applies them (symmetric
elimination, assembling.py:84-90,125-147; idempotent, so host callables that already
return constrained tensors may be combined with an empty or a repeated ``bcs`` list);
``ap`` gets ``bcs_pcd`` applied symmetrically (assembling.py:151-155); ``mp, mu, fp,
kp`` get none (:158-171); ``gp`` gets the velocity BCs row-wise, ``bc.apply(Bt)``
(:174-180).  Checked against the reference's own PCDAssembler in
tests/test_reference_golden.py.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

try:  # pragma: no cover
    import dolfin  # type: ignore
    HAVE_DOLFIN = True
except Exception:
    dolfin = None
    HAVE_DOLFIN = False


class PCDForm(object):
    """Wrapper of a PCD operator form with the two flags of the reference
    (assembling.py:201-230): ``const`` (assemble once) and ``phantom`` (not
    assembled, taken from the system matrix instead)."""

    def __init__(self, form, const=False, phantom=False):
        assert isinstance(const, bool) and isinstance(phantom, bool)
        self._form = form
        # public properties, as in the reference (assembling.py:222-224)
        self.constant = const
        self.phantom = phantom

    def dolfin_form(self):
        return self._form

    @property
    def ufl(self):
        return self._form

    def is_constant(self):
        return self.constant

    def is_phantom(self):
        return self.phantom


def _bc_dofs_values(bcs):
    """Concatenated (dofs, values) of a list of BC objects (``dofs()``, ``values()``)."""
    if not bcs:
        return np.zeros(0, dtype=np.int64), np.zeros(0)
    dofs = np.concatenate([np.asarray(bc.dofs(), dtype=np.int64) for bc in bcs])
    vals = np.concatenate([np.broadcast_to(np.asarray(bc.values(), dtype=np.float64), np.shape(bc.dofs())) for bc in bcs])
    return dofs, vals


def _dirichlet_rows(A, dofs):
    """``DirichletBC.apply(A)`` (MatZeroRows with unit diagonal): rows of ``dofs`` zeroed,
    diagonal set to one (added when the form's pattern lacks it; it lies outside every
    off-diagonal split block).  The zeroed entries leave the pattern; ``gp`` is a constant
    form, assembled once, so no later refresh depends on them."""
    A = sp.csr_matrix(A, copy=True)
    mask = np.zeros(A.shape[0], dtype=bool)
    mask[dofs] = True
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    A.data[mask[rows]] = 0.0
    A = (A + sp.diags(mask.astype(np.float64), shape=A.shape, format="csr")).tocsr()
    A.sort_indices()
    return A


def _symmetric_dirichlet(A, dofs):
    """Rows and columns of ``dofs`` zeroed, unit diagonal (what
    ``SystemAssembler`` does to the matrix); the pattern is kept."""
    A = sp.csr_matrix(A, copy=True)
    mask = np.zeros(A.shape[0], dtype=bool)
    mask[dofs] = True
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    kill = mask[rows] | mask[A.indices]
    A.data[kill] = 0.0
    A.data[kill & (rows == A.indices)] = 1.0
    return A


class PCDAssembler(object):
    """Collects the forms of the system and of the PCD operators and assembles
    them on demand.  Signature and defaults as the reference (assembling.py:35-37,
    98-106): ``ap, mp, mu, gp`` constant, ``fp, kp`` not, ``gp`` phantom."""

    def __init__(self, a, L, bcs, a_pc=None, mp=None, mu=None, ap=None, fp=None, kp=None, gp=None,
                 bcs_pcd=[], function_space=None):
        self._a, self._L, self._a_pc = a, L, a_pc
        self._bcs = list(bcs) if isinstance(bcs, (list, tuple)) else [bcs]
        self._bcs_pcd = bcs_pcd if bcs_pcd is None else (list(bcs_pcd) if isinstance(bcs_pcd, (list, tuple)) else [bcs_pcd])
        self._W = function_space
        self._forms = {
            "L": PCDForm(L),
            "ap": PCDForm(ap, const=True), "mp": PCDForm(mp, const=True), "mu": PCDForm(mu, const=True),
            "fp": PCDForm(fp), "kp": PCDForm(kp), "gp": PCDForm(gp, const=True, phantom=True),
        }
        # user may pass ready-made PCDForm instances to override the flags
        for key, f in (("ap", ap), ("mp", mp), ("mu", mu), ("fp", fp), ("kp", kp), ("gp", gp)):
            if isinstance(f, PCDForm):
                self._forms[key] = f
        if HAVE_DOLFIN and function_space is None and hasattr(a, "arguments"):  # pragma: no cover
            self._W = a.arguments()[0].function_space()
            self._assembler = dolfin.SystemAssembler(a, L, self._bcs)
            self._assembler_pc = dolfin.SystemAssembler(a_pc, L, self._bcs) if a_pc is not None else None

    # -- accessors -----------------------------------------------------------
    def function_space(self):
        return self._W

    def get_pcd_form(self, key):
        """Return form wrapped in ``PCDForm`` (reference assembling.py:108-114)."""
        form = self._forms.get(key)
        if form is None:
            raise AttributeError("Form '%s' requested by PCD not available" % key)
        return form

    def get_dolfin_form(self, key):
        """The wrapped form, ``None`` when it was not given (reference assembling.py:117-119)."""
        return self.get_pcd_form(key).dolfin_form()

    def pcd_bcs(self):
        """Artificial BCs of the PCD operator; only ``bcs_pcd=None`` is an error
        (reference assembling.py:184-189 -- the default ``[]`` is returned as is)."""
        if self._bcs_pcd is None:
            raise AttributeError("BCs requested by PCD not available")
        return self._bcs_pcd

    def bcs(self):
        return self._bcs

    # -- system --------------------------------------------------------------
    def _assemble(self, form, key=None):
        if form is None:
            raise AttributeError("Form '%s' requested by PCD not available" % key)
        return form() if callable(form) else form

    def _system(self, a_form, b=None, x=None):
        """What ``SystemAssembler(a, L, bcs)`` produces: matrix with the BC rows and columns
        eliminated, rhs lifted by the eliminated columns and set to the BC values (to
        ``g - x`` at the BC dofs in the Newton variant ``assemble(b, x)``)."""
        dofs, g = _bc_dofs_values(self._bcs)
        A = sp.csr_matrix(self._assemble(a_form, "a"))
        rhs = None
        if b is not None:
            rhs = np.array(self._assemble(self._L, "L"), dtype=np.float64)
            if dofs.size:
                if x is not None:
                    g = g - np.asarray(x.array if hasattr(x, "array") else x)[dofs]
                lift = np.zeros(A.shape[1])
                lift[dofs] = g
                rhs = rhs - A @ lift
                rhs[dofs] = g
        return (_symmetric_dirichlet(A, dofs) if dofs.size else A), rhs

    def system_matrix(self, A):
        A.set_csr(self._system(self._a)[0])

    def rhs_vector(self, b, x=None):
        b.array[:] = self._system(self._a, b, x)[1]

    def pc_matrix(self, P):
        if self._a_pc is None:
            return None
        P.set_csr(self._system(self._a_pc)[0])
        return P

    # -- PCD operators on the mixed space --------------------------------------
    def ap(self, Ap):
        A = self._assemble(self.get_dolfin_form("ap"), "ap")
        dofs, _ = _bc_dofs_values(self.pcd_bcs())
        Ap.set_csr(_symmetric_dirichlet(A, dofs))

    def mp(self, Mp):
        Mp.set_csr(self._assemble(self.get_dolfin_form("mp"), "mp"))

    def mu(self, Mu):
        Mu.set_csr(self._assemble(self.get_dolfin_form("mu"), "mu"))

    def fp(self, Fp):
        Fp.set_csr(self._assemble(self.get_dolfin_form("fp"), "fp"))

    def kp(self, Kp):
        Kp.set_csr(self._assemble(self.get_dolfin_form("kp"), "kp"))

    def gp(self, Bt):
        """Discrete pressure gradient with the velocity BC rows constrained (``bc.apply(Bt)``
        for every velocity BC, reference assembling.py:174-180)."""
        B = self._assemble(self.get_dolfin_form("gp"), "gp")
        dofs, _ = _bc_dofs_values(self._bcs)
        Bt.set_csr(_dirichlet_rows(B, dofs) if dofs.size else B)
