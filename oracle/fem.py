"""Minimal vectorised P2/P1 (Taylor-Hood) simplex assembler -- TEST INFRASTRUCTURE.

Stands in for DOLFIN, which assembles every operator of the reference on the
host (fenapack/assembling.py:127-180) and is not installed here.  The forms are
the ones of the reference demo and bench:

  system Jacobian (Picard/Oseen)   demo/navier-stokes-pcd/demo_navier-stokes-pcd.py:112-118
  mp = (1/nu) p q dx               demo_navier-stokes-pcd.py:130
  kp = (1/nu) (u_.grad p) q dx     demo_navier-stokes-pcd.py:131
  ap = grad p . grad q dx          demo_navier-stokes-pcd.py:132
  BRM2 Robin term on the inlet     demo_navier-stokes-pcd.py:133-136
  streamline-diffusion term        demo_navier-stokes-pcd.py:123-125, fenapack/stabilization.py:66-67
  reaction term (1/dt) for kp      demo/unsteady-navier-stokes-pcd/demo_unsteady-navier-stokes-pcd.py:138

Boundary-condition semantics follow fenapack/assembling.py:151-171: ``ap`` gets
the PCD Dirichlet conditions applied symmetrically (SystemAssembler), ``mp`` and
``kp`` get none; the system matrix gets the velocity conditions symmetrically.

Numbering produced here ("split" numbering): velocity dof = d*node + component
(P2 nodes = vertices followed by edges), pressure dof = vertex.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from itertools import permutations

import numpy as np
import scipy.sparse as sp
from scipy.special import roots_jacobi

# --------------------------------------------------------------------------
# meshes
# --------------------------------------------------------------------------


def rectangle_mesh(nx, ny, x0=0.0, y0=0.0, x1=1.0, y1=1.0):
    """Structured triangulation, every square cut by its "right" diagonal
    (lower-left to upper-right), as DOLFIN's ``UnitSquareMesh(n, n)``."""
    xs = np.linspace(x0, x1, nx + 1)
    ys = np.linspace(y0, y1, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    verts = np.column_stack([X.ravel(), Y.ravel()])
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (iy * (nx + 1) + ix).ravel()
    v1 = v0 + 1
    v2 = v0 + (nx + 1)
    v3 = v2 + 1
    cells = np.concatenate([np.column_stack([v0, v1, v3]), np.column_stack([v0, v2, v3])])
    return verts, cells.astype(np.int64)


def box_mesh(nx, ny, nz, lengths=(1.0, 1.0, 1.0)):
    """Structured tetrahedral mesh, six tetrahedra per brick, all sharing the
    brick diagonal (Kuhn triangulation; the layout DOLFIN's ``BoxMesh`` uses)."""
    xs = np.linspace(0.0, lengths[0], nx + 1)
    ys = np.linspace(0.0, lengths[1], ny + 1)
    zs = np.linspace(0.0, lengths[2], nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    verts = np.column_stack([X.ravel(), Y.ravel(), Z.ravel()])
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = ((iz * (ny + 1) + iy) * (nx + 1) + ix).ravel()
    step = np.array([1, nx + 1, (nx + 1) * (ny + 1)], dtype=np.int64)
    tets = []
    for perm in permutations(range(3)):
        a = base
        b = a + step[perm[0]]
        c = b + step[perm[1]]
        d = c + step[perm[2]]
        tets.append(np.column_stack([a, b, c, d]))
    return verts, np.concatenate(tets).astype(np.int64)


def load_dolfin_xml(path):
    """Read a DOLFIN XML triangle mesh (demo/data/mesh_lshape.xml)."""
    root = ET.parse(path).getroot()
    mesh = root.find("mesh")
    vs = mesh.find("vertices")
    verts = np.zeros((int(vs.get("size")), 2))
    for v in vs:
        verts[int(v.get("index"))] = (float(v.get("x")), float(v.get("y")))
    cs = mesh.find("cells")
    cells = np.zeros((int(cs.get("size")), 3), dtype=np.int64)
    for c in cs:
        cells[int(c.get("index"))] = (int(c.get("v0")), int(c.get("v1")), int(c.get("v2")))
    return verts, cells


# The 20-vertex / 22-triangle backward-facing-step mesh of the reference
# (demo/data/mesh_lshape.xml:5-51), restated as data so that nothing has to
# read /root/reference at run time.  tests/test_oracle_fem.py checks it against
# the XML file when the reference tree is present.
LSHAPE_VERTS = np.array(
    [(-1, 0), (-1, 1), (0, 0), (0, 1), (1, 0), (1, 1), (0, -1), (1, -1), (2, 0), (2, 1),
     (3, 0), (3, 1), (4, 0), (4, 1), (5, 0), (5, 1), (2, -1), (3, -1), (4, -1), (5, -1)],
    dtype=float)
LSHAPE_CELLS = np.array(
    [(0, 2, 1), (3, 1, 2), (2, 4, 3), (5, 3, 4), (4, 2, 7), (6, 7, 2), (4, 8, 5), (9, 5, 8),
     (8, 10, 9), (11, 9, 10), (10, 12, 11), (13, 11, 12), (12, 14, 13), (15, 13, 14),
     (7, 16, 4), (8, 4, 16), (16, 17, 8), (10, 8, 17), (17, 18, 10), (12, 10, 18),
     (18, 19, 12), (14, 12, 19)], dtype=np.int64)


def lshape_mesh(level):
    verts, cells = LSHAPE_VERTS.copy(), LSHAPE_CELLS.copy()
    for _ in range(level):
        verts, cells = refine_uniform(verts, cells)
    return verts, cells


def _unique_edges(cells):
    """Edges of a simplicial mesh: (edges[ne,2], cell_edges[nc,nle]) with local
    edge order (0,1),(0,2),(0,3),(1,2),(1,3),(2,3) (triangle: first three that exist)."""
    nv_loc = cells.shape[1]
    pairs = [(i, j) for i in range(nv_loc) for j in range(i + 1, nv_loc)]
    a = np.concatenate([cells[:, i] for i, _ in pairs])
    b = np.concatenate([cells[:, j] for _, j in pairs])
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    nv = int(cells.max()) + 1
    key = lo * nv + hi
    uniq, inv = np.unique(key, return_inverse=True)
    edges = np.column_stack([uniq // nv, uniq % nv])
    cell_edges = inv.reshape(len(pairs), cells.shape[0]).T.copy()
    return edges, cell_edges, pairs


def refine_uniform(verts, cells):
    """Regular (red) refinement of a triangle mesh: 4 children per triangle."""
    assert cells.shape[1] == 3
    edges, ce, _ = _unique_edges(cells)
    nv = verts.shape[0]
    mids = 0.5 * (verts[edges[:, 0]] + verts[edges[:, 1]])
    verts2 = np.vstack([verts, mids])
    m01, m02, m12 = nv + ce[:, 0], nv + ce[:, 1], nv + ce[:, 2]
    v0, v1, v2 = cells[:, 0], cells[:, 1], cells[:, 2]
    cells2 = np.concatenate([
        np.column_stack([v0, m01, m02]),
        np.column_stack([m01, v1, m12]),
        np.column_stack([m02, m12, v2]),
        np.column_stack([m01, m12, m02]),
    ])
    return verts2, cells2


# --------------------------------------------------------------------------
# quadrature and basis functions on the reference simplex (barycentric form)
# --------------------------------------------------------------------------


def simplex_quadrature(d, n):
    """Collapsed Gauss-Jacobi rule on the unit d-simplex, exact for total
    degree 2n-1.  Returns barycentric points [nq, d+1] and weights summing to 1."""
    if d == 0:
        return np.ones((1, 1)), np.ones(1)

    def gj(alpha):
        x, w = roots_jacobi(n, alpha, 0.0)
        return 0.5 * (x + 1.0), w * 0.5 ** (alpha + 1)

    if d == 1:
        t, w = gj(0)
        lam = np.column_stack([1 - t, t])
    elif d == 2:
        u, wu = gj(1)
        v, wv = gj(0)
        U, V = np.meshgrid(u, v, indexing="ij")
        W = np.outer(wu, wv)
        x, y = U.ravel(), (V * (1 - U)).ravel()
        lam = np.column_stack([1 - x - y, x, y])
        w = W.ravel()
    elif d == 3:
        u, wu = gj(2)
        v, wv = gj(1)
        t, wt = gj(0)
        U, V, T = np.meshgrid(u, v, t, indexing="ij")
        W = wu[:, None, None] * wv[None, :, None] * wt[None, None, :]
        x = U.ravel()
        y = (V * (1 - U)).ravel()
        z = (T * (1 - U) * (1 - V)).ravel()
        lam = np.column_stack([1 - x - y - z, x, y, z])
        w = W.ravel()
    else:
        raise ValueError(d)
    return lam, w / w.sum()


def p2_basis(lam, pairs):
    """P2 Lagrange basis in barycentric form: values [nq, nloc] and derivatives
    with respect to the barycentric coordinates [nq, nloc, d+1]."""
    nq, nb = lam.shape
    nloc = nb + len(pairs)
    phi = np.zeros((nq, nloc))
    dphi = np.zeros((nq, nloc, nb))
    for i in range(nb):
        phi[:, i] = lam[:, i] * (2 * lam[:, i] - 1)
        dphi[:, i, i] = 4 * lam[:, i] - 1
    for k, (i, j) in enumerate(pairs):
        phi[:, nb + k] = 4 * lam[:, i] * lam[:, j]
        dphi[:, nb + k, i] = 4 * lam[:, j]
        dphi[:, nb + k, j] = 4 * lam[:, i]
    return phi, dphi


# --------------------------------------------------------------------------
# function-space bookkeeping
# --------------------------------------------------------------------------


@dataclass
class TaylorHoodSpace:
    verts: np.ndarray          # [nv, d]
    cells: np.ndarray          # [nc, d+1]
    edges: np.ndarray = field(init=False)
    cell_nodes: np.ndarray = field(init=False)   # [nc, nloc2] P2 node ids
    pairs: list = field(init=False)
    grad_lam: np.ndarray = field(init=False)     # [nc, d+1, d]
    vol: np.ndarray = field(init=False)          # [nc]

    def __post_init__(self):
        self.edges, ce, self.pairs = _unique_edges(self.cells)
        self.cell_nodes = np.hstack([self.cells, self.nv + ce])
        d = self.dim
        X = self.verts[self.cells]                       # [nc, d+1, d]
        J = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))  # columns = edge vectors
        det = np.linalg.det(J)
        Jinv = np.linalg.inv(J)                          # rows = grad of lam_1..lam_d
        self.grad_lam = np.concatenate([-Jinv.sum(axis=1, keepdims=True), Jinv], axis=1)
        fact = {2: 2.0, 3: 6.0}[d]
        self.vol = np.abs(det) / fact

    @property
    def dim(self):
        return self.verts.shape[1]

    @property
    def nv(self):
        return self.verts.shape[0]

    @property
    def n1(self):               # P1 dofs
        return self.nv

    @property
    def n2(self):               # scalar P2 nodes
        return self.nv + self.edges.shape[0]

    @property
    def nu_dofs(self):
        return self.dim * self.n2

    @property
    def node_coords(self):
        mids = 0.5 * (self.verts[self.edges[:, 0]] + self.verts[self.edges[:, 1]])
        return np.vstack([self.verts, mids])

    def cell_diameter(self):
        """Circumdiameter-free stand-in for DOLFIN's ``Cell.h()`` = longest edge."""
        X = self.verts[self.cells]
        h = np.zeros(X.shape[0])
        for i, j in self.pairs:
            h = np.maximum(h, np.linalg.norm(X[:, i] - X[:, j], axis=1))
        return h

    def boundary_facets(self):
        """(cell, local vertex opposite to the facet) of every boundary facet."""
        nb = self.dim + 1
        nv = self.nv
        keys, owner, opp = [], [], []
        for k in range(nb):
            others = [i for i in range(nb) if i != k]
            f = np.sort(self.cells[:, others], axis=1)
            key = f[:, 0]
            for c in range(1, f.shape[1]):
                key = key * nv + f[:, c]
            keys.append(key)
            owner.append(np.arange(self.cells.shape[0]))
            opp.append(np.full(self.cells.shape[0], k))
        keys, owner, opp = map(np.concatenate, (keys, owner, opp))
        _, first, counts = np.unique(keys, return_index=True, return_counts=True)
        sel = first[counts == 1]
        return owner[sel], opp[sel]


def _coo_to_csr(rows, cols, vals, shape):
    A = sp.coo_matrix((vals.ravel(), (rows.ravel(), cols.ravel())), shape=shape).tocsr()
    A.sort_indices()
    return A


def _scatter_idx(rn, cn):
    """Row/col index arrays [nc, nr, ncol] for local matrices."""
    rows = np.repeat(rn[:, :, None], cn.shape[1], axis=2)
    cols = np.repeat(cn[:, None, :], rn.shape[1], axis=1)
    return rows, cols


def _chunks(n, size):
    for s in range(0, n, size):
        yield slice(s, min(n, s + size))


# --------------------------------------------------------------------------
# forms
# --------------------------------------------------------------------------


class Assembler:
    """Assembles the operators of the PCD-preconditioned Oseen problem on one
    TaylorHoodSpace.  All matrices are returned in scipy CSR with sorted
    indices and the *structural* sparsity pattern (value-independent, so that a
    re-assembly with another wind has the same pattern -- the property the
    reference relies on for MAT_REUSE_MATRIX, fenapack/field_split_backend.py:331-334)."""

    def __init__(self, space: TaylorHoodSpace, qorder=3, chunk=200_000):
        self.V = space
        d = space.dim
        self.lam, self.w = simplex_quadrature(d, qorder)
        self.phi, self.dphi = p2_basis(self.lam, space.pairs)
        self.chunk = chunk

    # -- helpers ----------------------------------------------------------
    def _grad_phi(self, sl):
        # [nc, nq, nloc, d] = dphi[q, l, k] * grad_lam[c, k, d]
        return np.einsum("qlk,ckd->cqld", self.dphi, self.V.grad_lam[sl], optimize=True)

    def wind_at_quad(self, wind, sl):
        """wind: [n2, d] nodal P2 values -> [nc, nq, d]."""
        Wc = wind[self.V.cell_nodes[sl]]                 # [nc, nloc, d]
        return np.einsum("ql,cld->cqd", self.phi, Wc, optimize=True)

    # -- scalar P2 operators (one velocity component) ----------------------
    def p2_scalar(self, nu=0.0, wind=None, mass_coeff=0.0, delta_sd=None):
        """nu*stiffness + convection(wind) + mass_coeff*mass [+ streamline diffusion]."""
        V = self.V
        n2 = V.n2
        rows, cols, vals = [], [], []
        for sl in _chunks(V.cells.shape[0], self.chunk):
            g = self._grad_phi(sl)
            vol = V.vol[sl]
            loc = np.zeros((g.shape[0], g.shape[2], g.shape[2]))
            if nu != 0.0:
                loc += nu * np.einsum("q,c,cqid,cqjd->cij", self.w, vol, g, g, optimize=True)
            if mass_coeff != 0.0:
                M = np.einsum("q,qi,qj->ij", self.w, self.phi, self.phi)
                loc += mass_coeff * vol[:, None, None] * M[None]
            if wind is not None:
                wq = self.wind_at_quad(wind, sl)
                wg = np.einsum("cqd,cqjd->cqj", wq, g, optimize=True)     # w . grad phi_j
                loc += np.einsum("q,c,qi,cqj->cij", self.w, vol, self.phi, wg, optimize=True)
                if delta_sd is not None:
                    loc += np.einsum("q,c,cqi,cqj->cij", self.w, vol * delta_sd[sl], wg, wg, optimize=True)
            r, c = _scatter_idx(V.cell_nodes[sl], V.cell_nodes[sl])
            rows.append(r.ravel()); cols.append(c.ravel()); vals.append(loc.ravel())
        return _coo_to_csr(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals), (n2, n2))

    def sd_parameter(self, wind, nu):
        """Streamline-diffusion parameter per cell (fenapack/stabilization.py:66-67)
        evaluated with the wind at the cell midpoint."""
        V = self.V
        d = V.dim
        lam_mid = np.full((1, d + 1), 1.0 / (d + 1))
        phi_mid, _ = p2_basis(lam_mid, V.pairs)
        wmid = np.einsum("l,cld->cd", phi_mid[0], wind[V.cell_nodes])
        wn = np.linalg.norm(wmid, axis=1)
        h = V.cell_diameter()
        pe = 0.5 * wn * h / nu
        with np.errstate(divide="ignore", invalid="ignore"):
            delta = np.where(pe > 1.0, 0.5 * h * (1.0 - 1.0 / pe) / wn, 0.0)
        return delta

    def velocity_block(self, scalar):
        """Expand a scalar P2 operator to the d-component block (dof = d*node+comp)."""
        d = self.V.dim
        return sp.kron(scalar, sp.identity(d, format="csr"), format="csr")

    def newton_coupling(self, wind):
        """int (phi_j e_c . grad) w_r  phi_i : the extra Newton term of
        ``derivative(F, w)`` (demo_navier-stokes-pcd.py:119-120), couples components."""
        V = self.V
        d = V.dim
        rows, cols, vals = [], [], []
        for sl in _chunks(V.cells.shape[0], self.chunk):
            g = self._grad_phi(sl)
            Wc = wind[V.cell_nodes[sl]]                               # [nc, nloc, d]
            gw = np.einsum("cqld,clr->cqrd", g, Wc, optimize=True)    # d w_r / d x_d at q
            m = np.einsum("q,c,qi,qj,cqrd->cirjd", self.w, V.vol[sl], self.phi, self.phi, gw, optimize=True)
            cn = V.cell_nodes[sl]
            rdof = (d * cn[:, :, None] + np.arange(d)[None, None, :]).reshape(cn.shape[0], -1)
            r, c = _scatter_idx(rdof, rdof)
            rows.append(r.ravel()); cols.append(c.ravel()); vals.append(m.reshape(cn.shape[0], -1).ravel())
        n = V.nu_dofs
        return _coo_to_csr(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals), (n, n))

    def divergence(self):
        """A10: rows = pressure dofs, cols = velocity dofs;  -int q div u."""
        V = self.V
        d = V.dim
        nb = d + 1
        psi = self.lam                                       # P1 basis values [nq, nb]
        rows, cols, vals = [], [], []
        for sl in _chunks(V.cells.shape[0], self.chunk):
            g = self._grad_phi(sl)                           # [nc, nq, nloc, d]
            loc = -np.einsum("q,c,qi,cqjd->cijd", self.w, V.vol[sl], psi, g, optimize=True)
            cn = V.cell_nodes[sl]
            cdof = (d * cn[:, :, None] + np.arange(d)[None, None, :]).reshape(cn.shape[0], -1)
            r, c = _scatter_idx(V.cells[sl], cdof)
            rows.append(r.ravel()); cols.append(c.ravel()); vals.append(loc.reshape(cn.shape[0], nb, -1).ravel())
        return _coo_to_csr(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals), (V.n1, V.nu_dofs))

    # -- P1 (pressure) operators -----------------------------------------
    def p1_mass(self, coeff=1.0):
        V = self.V
        d = V.dim
        nb = d + 1
        M = (np.ones((nb, nb)) + np.eye(nb)) / ((d + 1) * (d + 2))
        loc = coeff * V.vol[:, None, None] * M[None]
        r, c = _scatter_idx(V.cells, V.cells)
        return _coo_to_csr(r, c, loc, (V.n1, V.n1))

    def p1_laplace(self):
        V = self.V
        loc = V.vol[:, None, None] * np.einsum("cid,cjd->cij", V.grad_lam, V.grad_lam)
        r, c = _scatter_idx(V.cells, V.cells)
        return _coo_to_csr(r, c, loc, (V.n1, V.n1))

    def p1_convection(self, wind, coeff=1.0):
        """coeff * int (wind . grad p_j) q_i."""
        V = self.V
        rows, cols, vals = [], [], []
        for sl in _chunks(V.cells.shape[0], self.chunk):
            wq = self.wind_at_quad(wind, sl)
            wg = np.einsum("cqd,cjd->cqj", wq, V.grad_lam[sl], optimize=True)
            loc = coeff * np.einsum("q,c,qi,cqj->cij", self.w, V.vol[sl], self.lam, wg, optimize=True)
            r, c = _scatter_idx(V.cells[sl], V.cells[sl])
            rows.append(r.ravel()); cols.append(c.ravel()); vals.append(loc.ravel())
        return _coo_to_csr(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals), (V.n1, V.n1))

    def p1_boundary_flux_mass(self, wind, facet_mask_fn, coeff=1.0, qorder=3):
        """coeff * int_{Gamma} (wind . n) p q ds over boundary facets whose
        vertices all satisfy ``facet_mask_fn(coords)``  (BRM2 Robin term)."""
        V = self.V
        d = V.dim
        owner, opp = V.boundary_facets()
        nb = d + 1
        keep = np.ones(owner.size, dtype=bool)
        for k in range(nb):
            on = facet_mask_fn(V.verts[V.cells[owner, k]])
            keep &= on | (opp == k)
        owner, opp = owner[keep], opp[keep]
        lamf, wf = simplex_quadrature(d - 1, qorder)
        rows, cols, vals = [], [], []
        for k in range(nb):
            sel = owner[opp == k]
            if sel.size == 0:
                continue
            others = [i for i in range(nb) if i != k]
            lam = np.zeros((lamf.shape[0], nb))
            lam[:, others] = lamf
            phi, _ = p2_basis(lam, V.pairs)
            gk = V.grad_lam[sel, k, :]
            gnorm = np.linalg.norm(gk, axis=1)
            normal = -gk / gnorm[:, None]
            area = d * V.vol[sel] * gnorm
            wq = np.einsum("ql,cld->cqd", phi, wind[V.cell_nodes[sel]])
            wn = np.einsum("cqd,cd->cq", wq, normal)
            loc = coeff * np.einsum("q,c,cq,qi,qj->cij", wf, area, wn, lam, lam, optimize=True)
            r, c = _scatter_idx(V.cells[sel], V.cells[sel])
            rows.append(r.ravel()); cols.append(c.ravel()); vals.append(loc.ravel())
        if not rows:
            return sp.csr_matrix((V.n1, V.n1))
        return _coo_to_csr(np.concatenate(rows), np.concatenate(cols), np.concatenate(vals), (V.n1, V.n1))


# --------------------------------------------------------------------------
# Dirichlet conditions on assembled operators
# --------------------------------------------------------------------------


def apply_dirichlet_symmetric(A, bc_idx, compress=True):
    """Symmetric Dirichlet application as ``dolfin.SystemAssembler`` does it
    (fenapack/assembling.py:85-91,151-155): rows and columns of constrained dofs
    are zeroed and the diagonal is set to one.  With ``compress`` the zeroed
    entries are dropped from the pattern (a fixed, value-independent set)."""
    A = A.tocsr()
    n = A.shape[0]
    isbc = np.zeros(n, dtype=bool)
    isbc[bc_idx] = True
    A = A.tocoo()
    kill = isbc[A.row] | isbc[A.col]
    diag = A.row == A.col
    data = np.where(kill, 0.0, A.data)
    data = np.where(kill & diag, 1.0, data)
    if compress:
        keep = ~kill | diag
        out = sp.coo_matrix((data[keep], (A.row[keep], A.col[keep])), shape=A.shape).tocsr()
    else:
        out = sp.coo_matrix((data, (A.row, A.col)), shape=A.shape).tocsr()
    out.sort_indices()
    return out


def zero_rows(A, idx, compress=True):
    A = A.tocoo()
    mask = np.zeros(A.shape[0], dtype=bool)
    mask[idx] = True
    kill = mask[A.row]
    if compress:
        out = sp.coo_matrix((A.data[~kill], (A.row[~kill], A.col[~kill])), shape=A.shape).tocsr()
    else:
        out = sp.coo_matrix((np.where(kill, 0.0, A.data), (A.row, A.col)), shape=A.shape).tocsr()
    out.sort_indices()
    return out


def zero_cols(A, idx, compress=True):
    return zero_rows(A.T.tocsr(), idx, compress).T.tocsr()
