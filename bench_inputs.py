"""Synthetic-input generator for bench.py: P2/P1 Oseen operators on structured box
meshes, assembled with torch on the GPU (or CPU) -- inputs only, never the
measured path.  DOLFIN assembly stays on the host in production; this stands in
for it at sizes (tens of millions of dofs) where a numpy assembler is too slow.

Discretisation = oracle/fem.py (Kuhn triangulation, P2/P1, same forms:
demo_navier-stokes-pcd.py:112-137), but with the *lattice numbering*: scalar P2
node (X, Y, Z) of the (2nx+1)(2ny+1)(2nz+1) half-step lattice has index
(Z*(2ny+1)+Y)*(2nx+1)+X, pressure vertex (x, y, z) has (z*(ny+1)+y)*(nx+1)+x and
velocity dof = 3*node + component.  z-slabs of the lattice are contiguous index
ranges, which is what the row partition over GPUs uses.  tests/test_bench_inputs.py
checks the result against oracle/fem.py through the node permutation.
"""
from __future__ import annotations

from itertools import permutations

import numpy as np
import torch
from scipy.special import roots_jacobi

PAIRS = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]


def _tet_quadrature(n=3):
    def gj(alpha):
        x, w = roots_jacobi(n, alpha, 0.0)
        return 0.5 * (x + 1.0), w * 0.5 ** (alpha + 1)
    u, wu = gj(2)
    v, wv = gj(1)
    t, wt = gj(0)
    U, V, T = np.meshgrid(u, v, t, indexing="ij")
    W = wu[:, None, None] * wv[None, :, None] * wt[None, None, :]
    x, y, z = U.ravel(), (V * (1 - U)).ravel(), (T * (1 - U) * (1 - V)).ravel()
    lam = np.column_stack([1 - x - y - z, x, y, z])
    w = W.ravel()
    return lam, w / w.sum()


def _p2_tables(lam):
    nq = lam.shape[0]
    phi = np.zeros((nq, 10))
    dphi = np.zeros((nq, 10, 4))
    for i in range(4):
        phi[:, i] = lam[:, i] * (2 * lam[:, i] - 1)
        dphi[:, i, i] = 4 * lam[:, i] - 1
    for k, (i, j) in enumerate(PAIRS):
        phi[:, 4 + k] = 4 * lam[:, i] * lam[:, j]
        dphi[:, 4 + k, i] = 4 * lam[:, j]
        dphi[:, 4 + k, j] = 4 * lam[:, i]
    return phi, dphi


class BoxTaylorHood:
    """Geometry tables of the uniform Kuhn-triangulated box."""

    def __init__(self, nx, ny, nz, lengths=(1.0, 1.0, 1.0), device="cpu"):
        self.n = (nx, ny, nz)
        self.L = lengths
        self.h = np.array([lengths[0] / nx, lengths[1] / ny, lengths[2] / nz])
        self.dev = torch.device(device)
        self.lam, self.w = _tet_quadrature(3)
        self.phi, self.dphi = _p2_tables(self.lam)
        # six tetrahedra of the unit cube: vertex offsets (in cells)
        offs = []
        for perm in permutations(range(3)):
            v = np.zeros((4, 3), dtype=np.int64)
            for s in range(3):
                v[s + 1] = v[s]
                v[s + 1, perm[s]] += 1
            offs.append(v)
        self.tet_vert = np.stack(offs)                                  # [6, 4, 3] (x,y,z) offsets
        # P2 local node offsets on the half-step lattice: vertices 2*v, edges v_a+v_b
        node = np.zeros((6, 10, 3), dtype=np.int64)
        node[:, :4] = 2 * self.tet_vert
        for k, (a, b) in enumerate(PAIRS):
            node[:, 4 + k] = self.tet_vert[:, a] + self.tet_vert[:, b]
        self.tet_node = node
        # constant geometry per tet type
        X = self.tet_vert * self.h[None, None, :]
        J = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))
        Jinv = np.linalg.inv(J)
        self.grad_lam = np.concatenate([-Jinv.sum(axis=1, keepdims=True), Jinv], axis=1)   # [6, 4, 3]
        self.vol = np.abs(np.linalg.det(J)) / 6.0                                            # [6]
        self.gphi = np.einsum("qlk,tkd->tqld", self.dphi, self.grad_lam)                     # [6, nq, 10, 3]
        self.n2 = (2 * nx + 1) * (2 * ny + 1) * (2 * nz + 1)
        self.n1 = (nx + 1) * (ny + 1) * (nz + 1)

    def t(self, a, dtype=torch.float64):
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=self.dev)

    def node_id(self, X, Y, Z):
        nx, ny, nz = self.n
        return (Z * (2 * ny + 1) + Y) * (2 * nx + 1) + X

    def vert_id(self, x, y, z):
        nx, ny, nz = self.n
        return (z * (ny + 1) + y) * (nx + 1) + x

    def node_coords_of(self, ids):
        nx, ny, nz = self.n
        X = ids % (2 * nx + 1)
        Y = (ids // (2 * nx + 1)) % (2 * ny + 1)
        Z = ids // ((2 * nx + 1) * (2 * ny + 1))
        return X, Y, Z


def _coo_to_csr(rows, cols, vals, nrows_local, row0, ncols):
    """Sum duplicates; returns numpy (rowptr int32, col int32, val f64) of local rows.
    Deterministic: duplicates are accumulated in a fixed order (stable sort, then one
    scatter pass per duplicate rank -- no atomics), so that repeated runs assemble
    bit-identical operators and the solver's iteration counts are reproducible."""
    key = (rows - row0) * ncols + cols
    key, order = torch.sort(key, stable=True)
    vals = vals[order]
    ukey, inv, counts = torch.unique_consecutive(key, return_inverse=True, return_counts=True)
    first = torch.cumsum(counts, 0) - counts
    rank = torch.arange(key.numel(), device=key.device) - first[inv]
    out = torch.zeros(ukey.numel(), dtype=torch.float64, device=vals.device)
    for r in range(int(counts.max().item()) if key.numel() else 0):
        sel = torch.nonzero(rank == r, as_tuple=True)[0]
        out[inv[sel]] += vals[sel]
    r = ukey // ncols
    c = ukey - r * ncols
    cnt = torch.bincount(r, minlength=nrows_local)
    rowptr = torch.zeros(nrows_local + 1, dtype=torch.int64, device=vals.device)
    rowptr[1:] = torch.cumsum(cnt, 0)
    return rowptr.to(torch.int32).cpu().numpy(), c.to(torch.int32).cpu().numpy(), out.cpu().numpy()


def recirculating_wind(x, y, z):
    """Same analytic wind as oracle/problems.py (x-z recirculation)."""
    xs, zs = 2 * x - 1, 2 * z - 1
    return torch.stack([2 * zs * (1 - xs * xs), torch.zeros_like(x), -2 * xs * (1 - zs * zs)], dim=-1)


def poiseuille_wind(x, y, z):
    return torch.stack([16 * y * (1 - y) * z * (1 - z), torch.zeros_like(x), torch.zeros_like(x)], dim=-1)


class OseenBoxProblem:
    """Blocks of the Oseen system + PCD operators for the local row range of one rank.

    kind = "cavity": unit cube, lid z=1 moves with (1,0,0), no-slip elsewhere,
                     PCD Dirichlet set = pressure vertex 0, wind = recirculating.
    kind = "channel": box [0,L]x[0,1]^2, parabolic inflow at x=0, natural outflow
                     at x=L, PCD Dirichlet set = inlet (BRM1) / outlet (BRM2) vertices.
    The rank owns lattice planes [Z0, Z1) of the P2 lattice and [z0, z1) of the
    vertex lattice (balanced split of 2nz+1 and nz+1 planes).
    """

    def __init__(self, nx, ny, nz, kind="cavity", nu=0.02, variant="BRM2", lengths=None,
                 rank=0, nranks=1, device="cpu", layers_per_chunk=4):
        if lengths is None:
            lengths = (1.0, 1.0, 1.0) if kind == "cavity" else (4.0, 1.0, 1.0)
        self.g = g = BoxTaylorHood(nx, ny, nz, lengths, device)
        self.kind, self.nu, self.variant = kind, nu, variant
        self.rank, self.nranks = rank, nranks
        PZ = 2 * nz + 1
        self.Z0, self.Z1 = (PZ * rank) // nranks, (PZ * (rank + 1)) // nranks
        self.z0, self.z1 = ((nz + 1) * rank) // nranks, ((nz + 1) * (rank + 1)) // nranks
        plane2 = (2 * nx + 1) * (2 * ny + 1)
        plane1 = (nx + 1) * (ny + 1)
        self.node_begin, self.node_end = self.Z0 * plane2, self.Z1 * plane2
        self.p_begin, self.p_end = self.z0 * plane1, self.z1 * plane1
        self.n_u_global, self.n_p_global = 3 * g.n2, g.n1
        self.u_begin, self.n_u = 3 * self.node_begin, 3 * (self.node_end - self.node_begin)
        self.n_p = self.p_end - self.p_begin
        self.cheb_bounds = (0.5, 2.5)
        self.wind_fn = recirculating_wind if kind == "cavity" else poiseuille_wind
        self._assemble(layers_per_chunk)

    # -- boundary data -------------------------------------------------------
    def _node_bc(self, X, Y, Z):
        """(is Dirichlet node, boundary velocity [.,3]) for lattice nodes."""
        nx, ny, nz = self.g.n
        onb = (X == 0) | (X == 2 * nx) | (Y == 0) | (Y == 2 * ny) | (Z == 0) | (Z == 2 * nz)
        val = torch.zeros(X.shape + (3,), dtype=torch.float64, device=X.device)
        if self.kind == "cavity":
            val[..., 0] = (Z == 2 * nz).to(torch.float64)
            return onb, val
        walls = (Y == 0) | (Y == 2 * ny) | (Z == 0) | (Z == 2 * nz)
        outlet = (X == 2 * nx) & ~walls
        isbc = onb & ~outlet
        y = Y.to(torch.float64) * (0.5 * self.g.h[1])
        z = Z.to(torch.float64) * (0.5 * self.g.h[2])
        inlet = (X == 0)
        val[..., 0] = torch.where(inlet, 16 * y * (1 - y) * z * (1 - z), torch.zeros_like(y))
        return isbc, val

    def _vert_pcd_bc(self, x, y, z):
        nx = self.g.n[0]
        if self.kind == "cavity":
            return (x == 0) & (y == 0) & (z == 0)
        return (x == 0) if self.variant == "BRM1" else (x == nx)

    # -- assembly --------------------------------------------------------------
    def _assemble(self, layers):
        g = self.g
        nx, ny, nz = g.n
        dev = g.dev
        nu = self.nu
        phi, w = g.t(g.phi), g.t(g.w)
        lamq = g.t(g.lam)
        # cell layers needed: every cell touching an owned P2 plane or an owned vertex plane
        k_lo = max(0, min((self.Z0 - 1) // 2 if self.Z0 > 0 else 0, self.z0 - 1))
        k_hi = min(nz, max((self.Z1 - 1) // 2 + 1, self.z1))
        S_parts, B_parts, K_parts, M_parts, A_parts = [], [], [], [], []
        bu = torch.zeros(self.n_u, dtype=torch.float64, device=dev)
        bp = torch.zeros(self.n_p, dtype=torch.float64, device=dev)
        # constant local matrices per tet type
        Kloc = torch.stack([g.t(g.vol[t] * np.einsum("q,qid,qjd->ij", g.w, g.gphi[t], g.gphi[t])) for t in range(6)])
        Bloc = torch.stack([g.t(-g.vol[t] * np.einsum("q,qi,qjd->ijd", g.w, g.lam, g.gphi[t]).reshape(4, 30)) for t in range(6)])
        M1 = (np.ones((4, 4)) + np.eye(4)) / 20.0
        Mloc = torch.stack([g.t(g.vol[t] * M1 / nu) for t in range(6)])
        Aloc = torch.stack([g.t(g.vol[t] * g.grad_lam[t] @ g.grad_lam[t].T) for t in range(6)])
        gphi = g.t(g.gphi)
        glam = g.t(g.grad_lam)
        vol = g.t(g.vol)
        tet_node = g.t(g.tet_node, torch.int64)
        tet_vert = g.t(g.tet_vert, torch.int64)

        n2row = self.node_end - self.node_begin
        S_r, S_c, S_v = [], [], []
        for kb in range(k_lo, k_hi, layers):
            ke = min(k_hi, kb + layers)
            kk, jj, ii = torch.meshgrid(torch.arange(kb, ke, device=dev), torch.arange(ny, device=dev),
                                        torch.arange(nx, device=dev), indexing="ij")
            base = torch.stack([ii.reshape(-1), jj.reshape(-1), kk.reshape(-1)], dim=-1)     # [ncube, 3]
            for t in range(6):
                nodes3 = 2 * base[:, None, :] + tet_node[t][None]                           # [nc, 10, 3] lattice coords
                nid = g.node_id(nodes3[..., 0], nodes3[..., 1], nodes3[..., 2])            # [nc, 10]
                verts3 = base[:, None, :] + tet_vert[t][None]
                vid = g.vert_id(verts3[..., 0], verts3[..., 1], verts3[..., 2])            # [nc, 4]
                isbc, gval = self._node_bc(nodes3[..., 0], nodes3[..., 1], nodes3[..., 2])  # [nc,10], [nc,10,3]
                # wind at the P2 nodes -> quadrature points
                xyz = nodes3.to(torch.float64) * g.t(0.5 * g.h)
                Wn = self.wind_fn(xyz[..., 0], xyz[..., 1], xyz[..., 2])                    # [nc, 10, 3]
                wq = torch.einsum("ql,cld->cqd", phi, Wn)
                wg = torch.einsum("cqd,qjd->cqj", wq, gphi[t])
                Sl = nu * Kloc[t][None] + vol[t] * torch.einsum("q,qi,cqj->cij", w, phi, wg)   # [nc,10,10]
                # rhs lifting:  b_i -= sum_j S_ij g_j  (free rows), pressure rows: b_p -= B g
                lift_u = -torch.einsum("cij,cjd->cid", Sl, gval)                             # [nc,10,3]
                lift_p = -torch.einsum("ij,cj->ci", Bloc[t], gval.reshape(gval.shape[0], 30))
                rown = nid[:, :, None].expand(-1, -1, 10)
                coln = nid[:, None, :].expand(-1, 10, -1)
                keep = (rown >= self.node_begin) & (rown < self.node_end)
                free = ~isbc[:, :, None] & ~isbc[:, None, :]
                m = keep & free
                S_r.append(rown[m]); S_c.append(coln[m]); S_v.append(Sl[m])
                # velocity rhs
                rsel = (nid >= self.node_begin) & (nid < self.node_end) & ~isbc
                rows_u = (3 * (nid[rsel] - self.node_begin))[:, None] + torch.arange(3, device=dev)[None]
                bu.index_add_(0, rows_u.reshape(-1), lift_u[rsel].reshape(-1))
                # A10 block: rows = owned vertices, cols = velocity dofs (not Dirichlet)
                cdof = (3 * nid[:, :, None] + torch.arange(3, device=dev)[None, None]).reshape(-1, 30)
                cbc = isbc[:, :, None].expand(-1, -1, 3).reshape(-1, 30)
                rowv = vid[:, :, None].expand(-1, -1, 30)
                colv = cdof[:, None, :].expand(-1, 4, -1)
                mB = (rowv >= self.p_begin) & (rowv < self.p_end) & ~cbc[:, None, :]
                B_parts.append((rowv[mB], colv[mB], Bloc[t][None].expand(rowv.shape[0], -1, -1)[mB]))
                psel = (vid >= self.p_begin) & (vid < self.p_end)
                bp.index_add_(0, (vid[psel] - self.p_begin), lift_p[psel])
                # A01 block: rows = owned velocity dofs (not Dirichlet), cols = vertices
                rowu = cdof[:, :, None].expand(-1, -1, 4)
                colp = vid[:, None, :].expand(-1, 30, -1)
                nid30 = nid[:, :, None].expand(-1, -1, 3).reshape(-1, 30)
                mT = ((nid30 >= self.node_begin) & (nid30 < self.node_end) & ~cbc)[:, :, None].expand(-1, -1, 4)
                A_parts.append((rowu[mT], colp[mT], Bloc[t].T[None].expand(rowu.shape[0], -1, -1)[mT]))
                # pressure operators
                vx, vy, vz = verts3[..., 0], verts3[..., 1], verts3[..., 2]
                pbc = self._vert_pcd_bc(vx, vy, vz)
                rowp = vid[:, :, None].expand(-1, -1, 4)
                colpp = vid[:, None, :].expand(-1, 4, -1)
                own = (rowp >= self.p_begin) & (rowp < self.p_end)
                wgl = torch.einsum("cqd,jd->cqj", wq, glam[t])
                Kl = (vol[t] / nu) * torch.einsum("q,qi,cqj->cij", w, lamq, wgl)
                K_parts.append((rowp[own], colpp[own], Kl[own]))
                M_parts.append((rowp[own], colpp[own], Mloc[t][None].expand(rowp.shape[0], -1, -1)[own]))
                freep = own & ~pbc[:, :, None] & ~pbc[:, None, :]
                S_parts.append((rowp[freep], colpp[freep], Aloc[t][None].expand(rowp.shape[0], -1, -1)[freep]))
            # flush velocity COO of this chunk into a partial CSR to bound memory
        cat = lambda parts, i: torch.cat([p[i] for p in parts])
        # ---- scalar velocity operator + Dirichlet identity rows ----------------
        own_nodes = torch.arange(self.node_begin, self.node_end, device=dev)
        X, Y, Z = g.node_coords_of(own_nodes)
        isbc_own, gval_own = self._node_bc(X, Y, Z)
        bcn = own_nodes[isbc_own]
        r = torch.cat(S_r + [bcn]); c = torch.cat(S_c + [bcn])
        v = torch.cat(S_v + [torch.ones(bcn.numel(), dtype=torch.float64, device=dev)])
        del S_r, S_c, S_v
        rp, ci, va = _coo_to_csr(r, c, v, n2row, self.node_begin, g.n2)
        del r, c, v
        # expand the scalar operator to 3 components: dof = 3*node + comp
        cnt = np.diff(rp)
        rp3 = np.zeros(3 * n2row + 1, dtype=np.int64)
        rp3[1:] = np.cumsum(np.repeat(cnt, 3))
        assert rp3[-1] < 2 ** 31
        ci3 = np.empty(rp3[-1], dtype=np.int32)
        va3 = np.empty(rp3[-1])
        rows = np.repeat(np.arange(n2row), cnt)
        within = np.arange(ci.size) - np.repeat(rp[:-1], cnt)
        for comp in range(3):
            pos = rp3[3 * rows + comp] + within
            ci3[pos] = 3 * ci + comp
            va3[pos] = va
        self.A00 = (rp3.astype(np.int32), ci3, va3)
        self.S00 = (rp, ci, va)          # the scalar operator: A00 = S00 (x) I_3, local node rows, global node columns
        bu_h = bu.cpu().numpy()
        gv = gval_own[isbc_own].cpu().numpy()
        bidx = (3 * (bcn - self.node_begin).cpu().numpy()[:, None] + np.arange(3)[None]).ravel()
        bu_h[bidx] = gv.ravel()
        self.b_u = bu_h
        self.b_p = bp.cpu().numpy()
        self.A10 = _coo_to_csr(cat(B_parts, 0), cat(B_parts, 1), cat(B_parts, 2), self.n_p, self.p_begin, 3 * g.n2)
        del B_parts
        self.A01 = _coo_to_csr(cat(A_parts, 0), cat(A_parts, 1), cat(A_parts, 2), self.n_u, self.u_begin, g.n1)
        del A_parts
        self.Kp = _coo_to_csr(cat(K_parts, 0), cat(K_parts, 1), cat(K_parts, 2), self.n_p, self.p_begin, g.n1)
        self.Mp = _coo_to_csr(cat(M_parts, 0), cat(M_parts, 1), cat(M_parts, 2), self.n_p, self.p_begin, g.n1)
        own_v = torch.arange(self.p_begin, self.p_end, device=dev)
        x = own_v % (nx + 1); y = (own_v // (nx + 1)) % (ny + 1); z = own_v // ((nx + 1) * (ny + 1))
        pbc_own = own_v[self._vert_pcd_bc(x, y, z)]
        r = torch.cat([cat(S_parts, 0), pbc_own]); c = torch.cat([cat(S_parts, 1), pbc_own])
        v = torch.cat([cat(S_parts, 2), torch.ones(pbc_own.numel(), dtype=torch.float64, device=dev)])
        self.Ap = _coo_to_csr(r, c, v, self.n_p, self.p_begin, g.n1)
        self.bc_idx = (pbc_own - self.p_begin).to(torch.int32).cpu().numpy()
        self.bc_val = np.zeros(self.bc_idx.size)
        self.ndofs_global = self.n_u_global + self.n_p_global

    def scipy(self, name):
        import scipy.sparse as sp
        rp, ci, va = getattr(self, name)
        ncols = {"A00": self.n_u_global, "A10": self.n_u_global, "S00": self.n_u_global // 3}.get(name, self.n_p_global)
        return sp.csr_matrix((va, ci, rp), shape=(rp.size - 1, ncols))


class MergedProblem:
    """The rows of several consecutive z-slabs (OseenBoxProblem instances generated as ranks
    s0 .. s0+k-1 of a finer partition) as ONE local problem: row pointers chained, column ids
    are global already.  Bounds the generator's device memory at large n (one slab of the 128^3
    cavity needs as much as the whole 64^3 one)."""

    OPS = ("A00", "A01", "A10", "Ap", "Mp", "Kp", "S00")

    def __init__(self, parts):
        p0, pl = parts[0], parts[-1]
        for k in ("kind", "nu", "variant", "n_u_global", "n_p_global", "cheb_bounds", "ndofs_global"):
            setattr(self, k, getattr(p0, k))
        self.u_begin, self.p_begin = p0.u_begin, p0.p_begin
        self.n_u = sum(p.n_u for p in parts)
        self.n_p = sum(p.n_p for p in parts)
        assert pl.u_begin + pl.n_u - p0.u_begin == self.n_u and pl.p_begin + pl.n_p - p0.p_begin == self.n_p
        for name in self.OPS:
            rps, cis, vas = zip(*[getattr(p, name) for p in parts])
            nnz = np.cumsum([0] + [int(r[-1]) for r in rps])
            assert nnz[-1] < 2 ** 31, "operator exceeds 2^31 entries on one rank"
            rp = np.concatenate([rps[0]] + [r[1:].astype(np.int64) + nnz[i + 1] for i, r in enumerate(rps[1:])]).astype(np.int32)
            setattr(self, name, (rp, np.concatenate(cis), np.concatenate(vas)))
            for p in parts:
                setattr(p, name, None)          # free the slab copies as we go
        self.b_u = np.concatenate([p.b_u for p in parts])
        self.b_p = np.concatenate([p.b_p for p in parts])
        off = np.cumsum([0] + [p.n_p for p in parts])
        self.bc_idx = np.concatenate([p.bc_idx.astype(np.int64) + off[i] for i, p in enumerate(parts)]).astype(np.int32)
        self.bc_val = np.concatenate([p.bc_val for p in parts])

    scipy = OseenBoxProblem.scipy


def generate(nx, ny, nz, kind="cavity", nu=0.02, variant="BRM2", rank=0, nranks=1, device="cpu",
             cells_per_slab=64 ** 3):
    """Local problem of `rank`, generated in z-slabs of about `cells_per_slab` cells."""
    k = max(1, int(np.ceil(nx * ny * nz / nranks / cells_per_slab)))
    k = min(k, max(1, (nz + 1) // nranks))        # every slab keeps at least one vertex plane
    if k == 1:
        return OseenBoxProblem(nx, ny, nz, kind=kind, nu=nu, variant=variant, rank=rank, nranks=nranks, device=device)
    parts = []
    for s in range(k):
        parts.append(OseenBoxProblem(nx, ny, nz, kind=kind, nu=nu, variant=variant, rank=rank * k + s,
                                     nranks=nranks * k, device=device))
        if str(device).startswith("cuda"):
            torch.cuda.empty_cache()
    return MergedProblem(parts)
