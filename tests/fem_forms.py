"""DOLFIN-less "forms" for the drop-in API tests: host assembler callables on the
mixed P2/P1 space of the reference's backward-facing-step problem
(test/bench/test_pcd_scaling.py:30-148), built on the oracle assembler."""
import numpy as np
import scipy.sparse as sp

from fenapack_b200.petsc_shim import Vec
from oracle import fem, problems


def struct_add(*mats):
    """Sum of sparse matrices that keeps the union of the *stored* patterns (scipy's
    ``+`` drops entries that cancel to zero, which would make the pattern depend on
    the values -- DOLFIN never does that)."""
    coo = [sp.coo_matrix(m) for m in mats]
    out = sp.coo_matrix((np.concatenate([c.data for c in coo]),
                         (np.concatenate([c.row for c in coo]), np.concatenate([c.col for c in coo]))),
                        shape=mats[0].shape).tocsr()
    out.sort_indices()
    return out


def permute(A, perm_rows, perm_cols):
    """P A Q^T for permutations given as index maps, preserving stored zeros."""
    c = sp.coo_matrix(A)
    out = sp.coo_matrix((c.data, (perm_rows[c.row], perm_cols[c.col])), shape=A.shape).tocsr()
    out.sort_indices()
    return out


class _DofMap:
    def __init__(self, dofs):
        self._d = np.asarray(dofs, dtype=np.int64)

    def dofs(self):
        return self._d


class _Sub:
    def __init__(self, dofs):
        self._dm = _DofMap(dofs)

    def dofmap(self):
        return self._dm


class MixedSpace:
    def __init__(self, is_u, is_p):
        self._subs = (_Sub(is_u), _Sub(is_p))

    def sub(self, i):
        return self._subs[i]

    def dim(self):
        return self._subs[0].dofmap().dofs().size + self._subs[1].dofmap().dofs().size


class DirichletDofs:
    def __init__(self, dofs, values):
        self._dofs, self._vals = np.asarray(dofs, dtype=np.int64), values

    def dofs(self):
        return self._dofs

    def values(self):
        return self._vals


class BFSModel:
    """Picard-linearised Navier-Stokes on the BFS mesh; ``w`` (monolithic Vec) is
    the current iterate all forms are evaluated at."""

    def __init__(self, level=2, nu=0.02, variant="BRM1", idt=0.0):
        self.space = S = problems.bfs_space(level)
        self.nu, self.variant, self.idt = nu, variant, idt
        self.asm = fem.Assembler(S)
        self.is_u, self.is_p = problems.interleaved_index_sets(S)
        self.N = S.nu_dofs + S.n1
        self.W = MixedSpace(self.is_u, self.is_p)
        self.w = Vec(np.zeros(self.N))
        dirichlet, vals, _, _, _ = problems.bfs_boundary(S)
        self.bc_u = (2 * dirichlet[:, None] + np.arange(2)[None, :]).ravel()      # split-u numbering
        self.g_u_steady = vals.ravel()
        self.g_u = self.g_u_steady.copy()
        vx = S.verts[:, 0]
        mask = np.isclose(vx, -1.0) if variant == "BRM1" else np.isclose(vx, 5.0)
        self.bc_pcd = DirichletDofs(self.is_p[np.flatnonzero(mask)], 0.0)
        self.K = self.asm.velocity_block(self.asm.p2_scalar(nu=nu, mass_coeff=idt))
        self.A10 = self.asm.divergence()
        self.A01 = self.A10.T.tocsr()
        # permutation split -> monolithic
        n_u = S.nu_dofs
        self.Pm = sp.csr_matrix((np.ones(self.N), (np.concatenate([self.is_u, self.is_p]), np.arange(self.N))),
                                shape=(self.N, self.N))
        self.n_u = n_u
        self.split2mono = np.concatenate([self.is_u, self.is_p])
        # backward Euler: (1/dt) M (u - u0) in the residual (demo_unsteady-navier-stokes-pcd.py:118-127)
        self.Mu_dt = self.asm.velocity_block(self.asm.p2_scalar(mass_coeff=idt)) if idt else None
        self.w0 = Vec(np.zeros(self.N))

    def set_time(self, t, t_ramp=1.0):
        """Inflow ramp of the unsteady demo (demo_unsteady-navier-stokes-pcd.py:80-84):
        the parabolic profile is scaled by a smooth ramp in time."""
        f = 1.0 if t >= t_ramp else 0.5 * (1.0 - np.cos(np.pi * t / t_ramp))
        self.g_u = f * self.g_u_steady

    def wind(self):
        return self.w.array[self.is_u].reshape(-1, 2)

    def _jacobian_free(self, stabilised=False):
        delta = self.asm.sd_parameter(self.wind(), self.nu) if stabilised else None
        C = self.asm.velocity_block(self.asm.p2_scalar(nu=0.0, wind=self.wind(), delta_sd=delta))
        A00 = struct_add(self.K, C)
        n_u = self.n_u
        blocks = [sp.coo_matrix(A00), sp.coo_matrix(self.A01), sp.coo_matrix(self.A10)]
        offs = [(0, 0), (0, n_u), (n_u, 0)]
        J = sp.coo_matrix((np.concatenate([b.data for b in blocks]),
                           (np.concatenate([b.row + o[0] for b, o in zip(blocks, offs)]),
                            np.concatenate([b.col + o[1] for b, o in zip(blocks, offs)]))),
                          shape=(self.N, self.N)).tocsr()
        J.sort_indices()
        return J

    def _system(self, stabilised=False):
        J = self._jacobian_free(stabilised)
        ws = np.concatenate([self.w.array[self.is_u], self.w.array[self.is_p]])
        F = J @ ws
        if self.Mu_dt is not None:
            F[:self.n_u] -= self.Mu_dt @ self.w0.array[self.is_u]
        dxbc = ws[self.bc_u] - self.g_u
        lift = np.zeros(self.N)
        lift[self.bc_u] = dxbc
        b = F - J @ lift
        b[self.bc_u] = dxbc
        Jbc = fem.apply_dirichlet_symmetric(J, self.bc_u, compress=False)
        return Jbc, b

    # forms (monolithic numbering) ------------------------------------------------
    def a(self):
        J, _ = self._system()
        return permute(J, self.split2mono, self.split2mono)

    def L(self):
        _, b = self._system()
        return self.Pm @ b

    def a_pc(self):
        """Jacobian with the streamline-diffusion term for the AMG 00-block
        (demo_navier-stokes-pcd.py:123-125)."""
        J, _ = self._system(stabilised=True)
        return permute(J, self.split2mono, self.split2mono)

    def _embed_p(self, Mp):
        c = sp.coo_matrix(Mp)
        out = sp.coo_matrix((c.data, (self.is_p[c.row], self.is_p[c.col])), shape=(self.N, self.N)).tocsr()
        out.sort_indices()
        return out

    def mp(self):
        return self._embed_p(self.asm.p1_mass(1.0 / self.nu))

    def mu(self):
        """Velocity mass matrix (1/dt) (u, v) on the mixed space (PCDR only), no BCs."""
        Mu = self.asm.velocity_block(self.asm.p2_scalar(mass_coeff=self.idt))
        c = sp.coo_matrix(Mu)
        out = sp.coo_matrix((c.data, (self.is_u[c.row], self.is_u[c.col])), shape=(self.N, self.N)).tocsr()
        out.sort_indices()
        return out

    def ap(self):
        return self._embed_p(self.asm.p1_laplace())

    def kp(self):
        parts = [self.asm.p1_convection(self.wind(), 1.0 / self.nu)]
        if self.variant == "BRM2":
            parts.append(-self.asm.p1_boundary_flux_mass(self.wind(), lambda x: np.isclose(x[:, 0], -1.0), 1.0 / self.nu))
        # structural P1 pattern, so that refreshes are value-only
        return self._embed_p(struct_add(*parts))


# ---------------------------------------------------------------------------------------------
# row-partitioned view of a model (multi-rank drop-in tests): every rank assembles the whole
# tensor on the host (test harness only) and hands over its contiguous range of rows of the
# monolithic numbering, the way DOLFIN's tensors arrive in an MPI run
# ---------------------------------------------------------------------------------------------
class _OwnedDofMap(_DofMap):
    def __init__(self, dofs, rows):
        super().__init__(dofs)
        self._rows = rows

    def ownership_range(self):
        return self._rows


class PartitionedSpace(MixedSpace):
    def __init__(self, is_u, is_p, rows):
        r0, r1 = rows
        self._rows = rows
        self._subs = (_Sub(is_u[(is_u >= r0) & (is_u < r1)]), _Sub(is_p[(is_p >= r0) & (is_p < r1)]))

    def dofmap(self):
        return _OwnedDofMap(np.arange(*self._rows), self._rows)


class PartitionedModel:
    """Rows [r0, r1) of every form of ``model`` (global column ids)."""

    def __init__(self, model, rank, nranks):
        self.m = model
        N = model.N
        self.rows = (N * rank // nranks, N * (rank + 1) // nranks)
        self.W = PartitionedSpace(model.is_u, model.is_p, self.rows)
        self.bc_pcd = model.bc_pcd                     # all constrained dofs, global ids

    def _rows_of(self, f):
        r0, r1 = self.rows

        def g():
            T = f()
            return T[r0:r1] if isinstance(T, np.ndarray) else sp.csr_matrix(T)[r0:r1, :]
        return g

    def __getattr__(self, name):
        if name in ("a", "L", "a_pc", "mp", "mu", "ap", "kp"):
            return self._rows_of(getattr(self.m, name))
        raise AttributeError(name)
