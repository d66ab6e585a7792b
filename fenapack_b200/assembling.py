"""Form bookkeeping of the PCD preconditioner -- drop-in for fenapack/assembling.py.

Assembly stays on the host (BASELINE.json north_star).  In the reference the
``a, L, mp, ap, kp ...`` arguments are UFL forms handed to ``dolfin.SystemAssembler`` /
``dolfin.assemble`` (assembling.py:85-180).  DOLFIN is not available in this image and that
branch is NOT implemented here (UFL forms raise ``NotImplementedError``): every "form" is a
*host assembler callable* ``form() -> scipy.sparse matrix`` (or vector) on the mixed space in
monolithic numbering -- this rank's rows when the function space reports an ownership range,
the whole tensor otherwise -- and a boundary condition is any object with ``dofs()``
(monolithic indices, all constrained dofs) and ``values()``.  A FEniCS user wraps
``lambda: as_scipy(assemble(form))``; the contract seen by PCDInterface is the same:
tensors on the mixed space W.

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


def _dirichlet_rows(A, dofs, row0=0):
    """``DirichletBC.apply(A)`` (MatZeroRows with unit diagonal): rows of ``dofs`` zeroed,
    diagonal set to one (added when the form's pattern lacks it; it lies outside every
    off-diagonal split block).  The zeroed entries leave the pattern; ``gp`` is a constant
    form, assembled once, so no later refresh depends on them.  ``A`` holds the rows
    ``row0 ..`` of the global matrix (row partition); ``dofs`` are global ids."""
    A = sp.csr_matrix(A, copy=True)
    mask = np.zeros(A.shape[1], dtype=bool)
    mask[dofs] = True
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr)) + row0
    A.data[mask[rows]] = 0.0
    loc = np.arange(A.shape[0])
    own = mask[loc + row0]
    D = sp.csr_matrix((own.astype(np.float64)[own], (loc[own], loc[own] + row0)), shape=A.shape)
    A = (A + D).tocsr()
    A.sort_indices()
    return A


def _symmetric_dirichlet(A, dofs, row0=0):
    """Rows and columns of ``dofs`` zeroed, unit diagonal (what
    ``SystemAssembler`` does to the matrix); the pattern is kept.  Row-partition aware as
    ``_dirichlet_rows``."""
    A = sp.csr_matrix(A, copy=True)
    mask = np.zeros(A.shape[1], dtype=bool)
    mask[dofs] = True
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr)) + row0
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
        if function_space is None and hasattr(L, "arguments"):  # pragma: no cover - UFL forms need DOLFIN
            # reference assembling.py:98: the test function's space
            self._W = L.arguments()[0].function_space()
        if not all(callable(f) or f is None or isinstance(f, PCDForm) or sp.issparse(f) or isinstance(f, np.ndarray)
                   for f in (a, L, a_pc, mp, mu, ap, fp, kp, gp)):
            raise NotImplementedError(
                "PCDAssembler: UFL forms need DOLFIN's SystemAssembler / assemble (reference assembling.py:85-180), "
                "which is not available here; pass host assembler callables returning scipy matrices instead")
        # row partition: every callable returns this rank's rows [row0, row1) of the tensor on the mixed
        # space (DOLFIN: GenericDofMap.ownership_range); one rank owns everything
        self._rows = None
        dm = getattr(self._W, "dofmap", None)
        if callable(dm) and hasattr(dm(), "ownership_range"):
            self._rows = tuple(int(v) for v in dm().ownership_range())

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
        row0 = self._row0()
        rhs = None
        if b is not None:
            rhs = np.array(self._assemble(self._L, "L"), dtype=np.float64)
            if dofs.size:
                if x is not None:
                    xa = np.asarray(x.array if hasattr(x, "array") else x)
                    if xa.size != A.shape[1]:
                        # row-partitioned iterate: the lifting needs the values at every constrained dof
                        xa = np.concatenate(x.comm.allgather(xa))
                    g = g - xa[dofs]
                lift = np.zeros(A.shape[1])
                lift[dofs] = g
                rhs = rhs - A @ lift
                own = (dofs >= row0) & (dofs < row0 + A.shape[0])
                rhs[dofs[own] - row0] = g[own]
        return (_symmetric_dirichlet(A, dofs, row0) if dofs.size else A), rhs

    def _row0(self):
        return self._rows[0] if self._rows is not None else 0

    def _fill(self, mat, csr):
        if self._rows is not None:
            mat.set_csr(csr, row_range=self._rows)
        else:
            mat.set_csr(csr)

    def system_matrix(self, A):
        self._fill(A, self._system(self._a)[0])

    def rhs_vector(self, b, x=None):
        b.array[:] = self._system(self._a, b, x)[1]

    def pc_matrix(self, P):
        if self._a_pc is None:
            return None
        self._fill(P, self._system(self._a_pc)[0])
        return P

    # -- PCD operators on the mixed space --------------------------------------
    def ap(self, Ap):
        A = self._assemble(self.get_dolfin_form("ap"), "ap")
        dofs, _ = _bc_dofs_values(self.pcd_bcs())
        self._fill(Ap, _symmetric_dirichlet(A, dofs, self._row0()))

    def mp(self, Mp):
        self._fill(Mp, self._assemble(self.get_dolfin_form("mp"), "mp"))

    def mu(self, Mu):
        self._fill(Mu, self._assemble(self.get_dolfin_form("mu"), "mu"))

    def fp(self, Fp):
        self._fill(Fp, self._assemble(self.get_dolfin_form("fp"), "fp"))

    def kp(self, Kp):
        self._fill(Kp, self._assemble(self.get_dolfin_form("kp"), "kp"))

    def gp(self, Bt):
        """Discrete pressure gradient with the velocity BC rows constrained (``bc.apply(Bt)``
        for every velocity BC, reference assembling.py:174-180)."""
        B = self._assemble(self.get_dolfin_form("gp"), "gp")
        dofs, _ = _bc_dofs_values(self._bcs)
        self._fill(Bt, _dirichlet_rows(B, dofs, self._row0()) if dofs.size else B)
