"""Scenario builders for the PCD-preconditioned Oseen problem -- TEST INFRASTRUCTURE.

Problem definitions restate the reference bench/demo set-ups:
  backward-facing step   test/bench/test_pcd_scaling.py:30-148, demo_navier-stokes-pcd.py:50-137
  unsteady BFS           demo/unsteady-navier-stokes-pcd/demo_unsteady-navier-stokes-pcd.py:95-140
and add the box-domain cases named in BASELINE.json (lid-driven cavities, 3D
channel).  Every builder returns a ``PCDProblem`` holding the blocks of the
linearised (Oseen) system in split numbering plus the PCD operators exactly as
``PCDAssembler`` (fenapack/assembling.py:127-180) would hand them to the PC.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from . import fem


@dataclass
class PCDProblem:
    name: str
    dim: int
    nu: float
    variant: str                # "BRM1" | "BRM2"
    A00: sp.csr_matrix          # velocity block of the system matrix (with velocity BCs)
    A01: sp.csr_matrix          # = -B^T with BC rows zeroed
    A10: sp.csr_matrix          # = -B   with BC columns zeroed
    Mp: sp.csr_matrix
    Ap: sp.csr_matrix           # with PCD Dirichlet BCs applied symmetrically
    Kp: sp.csr_matrix
    bc_idx: np.ndarray          # PCD Dirichlet dofs in pressure ("p" split) numbering
    bc_val: np.ndarray
    b_u: np.ndarray
    b_p: np.ndarray
    P00: sp.csr_matrix | None = None    # stabilised velocity block for the "iterative" set-up
    is_u: np.ndarray | None = None      # monolithic numbering of the split dofs
    is_p: np.ndarray | None = None
    cheb_bounds: tuple = (0.5, 2.0)
    mu_diag: np.ndarray | None = None   # diagonal of the velocity mass matrix (1/dt) (u, v), PCDR only
    meta: dict = field(default_factory=dict)

    @property
    def n_u(self):
        return self.A00.shape[0]

    @property
    def n_p(self):
        return self.Mp.shape[0]

    def system_matrix(self):
        return sp.bmat([[self.A00, self.A01], [self.A10, None]], format="csr")

    def rhs(self):
        return np.concatenate([self.b_u, self.b_p])


def interleaved_index_sets(space: fem.TaylorHoodSpace):
    """A DOLFIN-like monolithic numbering: per vertex [u_0..u_{d-1}, p], then
    per edge [u_0..u_{d-1}].  Returns (is_u, is_p): positions of the split dofs in
    the monolithic vector -- the role of ``dofmap_dofs_is``
    (fenapack/_field_split_utils.py:39-50)."""
    d, nv, ne = space.dim, space.nv, space.edges.shape[0]
    is_p = (d + 1) * np.arange(nv, dtype=np.int64) + d
    vert_u = ((d + 1) * np.arange(nv, dtype=np.int64)[:, None] + np.arange(d)[None, :]).ravel()
    edge_u = (d + 1) * nv + np.arange(d * ne, dtype=np.int64)
    return np.concatenate([vert_u, edge_u]), is_p


def _build(name, space, nu, variant, wind, vel_bc_nodes, vel_bc_vals, pcd_bc_mask,
           inlet_mask_fn=None, idt=0.0, stabilise=False, newton=False, cheb_bounds=None,
           qorder=3, pcdr=False):
    """Common assembly path.  ``wind``: [n2, d] nodal P2 wind.  ``vel_bc_nodes``:
    P2 node ids with Dirichlet velocity, ``vel_bc_vals``: [len, d]."""
    d = space.dim
    asm = fem.Assembler(space, qorder=qorder)
    S = asm.p2_scalar(nu=nu, wind=wind, mass_coeff=idt)
    A00 = asm.velocity_block(S)
    if newton:
        A00 = (A00 + asm.newton_coupling(wind)).tocsr()
        A00.sort_indices()
    A10 = asm.divergence()
    A01 = A10.T.tocsr()
    A01.sort_indices()
    P00 = None
    if stabilise:
        delta = asm.sd_parameter(wind, nu)
        if np.any(delta > 0):
            P00 = asm.velocity_block(asm.p2_scalar(nu=nu, wind=wind, mass_coeff=idt, delta_sd=delta))
            if newton:
                P00 = (P00 + asm.newton_coupling(wind)).tocsr()
                P00.sort_indices()

    # velocity Dirichlet conditions, symmetric (SystemAssembler)
    bc_dofs = (d * vel_bc_nodes[:, None] + np.arange(d)[None, :]).ravel()
    g = np.zeros(space.nu_dofs)
    g[bc_dofs] = vel_bc_vals.ravel()
    b_u = -(A00 @ g)
    b_p = -(A10 @ g)
    b_u[bc_dofs] = g[bc_dofs]
    A00 = fem.apply_dirichlet_symmetric(A00, bc_dofs)
    if P00 is not None:
        P00 = fem.apply_dirichlet_symmetric(P00, bc_dofs)
    A01 = fem.zero_rows(A01, bc_dofs)
    A10 = fem.zero_cols(A10, bc_dofs)

    # PCD operators (fenapack/assembling.py:151-171)
    Mp = asm.p1_mass(1.0 / nu)
    Kp = asm.p1_convection(wind, 1.0 / nu)
    mu_diag = None
    if idt != 0.0 and not pcdr:
        # PCD for unsteady problems: reaction term in Kp (demo_unsteady-navier-stokes-pcd.py:138)
        Kp = (Kp + asm.p1_mass(idt / nu)).tocsr()
    if pcdr:
        # PCDR: Kp stays pure convection, the reaction enters through Rp built from
        # mu = (1/dt) (u, v) (demo_unsteady-navier-stokes-pcdr.py:137-139); no BCs on mu
        mu_diag = np.repeat(asm.p2_scalar(mass_coeff=idt).diagonal(), d)
    if variant == "BRM2" and inlet_mask_fn is not None:
        Kp = (Kp - asm.p1_boundary_flux_mass(wind, inlet_mask_fn, 1.0 / nu)).tocsr()
    Kp.sort_indices()
    bc_idx = np.flatnonzero(pcd_bc_mask).astype(np.int32)
    Ap = fem.apply_dirichlet_symmetric(asm.p1_laplace(), bc_idx)

    is_u, is_p = interleaved_index_sets(space)
    if cheb_bounds is None:
        cheb_bounds = (0.5, 2.0) if d == 2 else (0.5, 2.5)
    return PCDProblem(name=name, dim=d, nu=nu, variant=variant, A00=A00, A01=A01, A10=A10,
                      Mp=Mp, Ap=Ap, Kp=Kp, bc_idx=bc_idx, bc_val=np.zeros(bc_idx.size),
                      b_u=b_u, b_p=b_p, P00=P00, is_u=is_u, is_p=is_p, cheb_bounds=cheb_bounds, mu_diag=mu_diag,
                      meta={"n2": space.n2, "n1": space.n1, "cells": space.cells.shape[0],
                            "ndofs": space.nu_dofs + space.n1})


def _on_boundary_nodes(space):
    """P2 node ids on the boundary (vertices and edge midpoints of boundary facets)."""
    owner, opp = space.boundary_facets()
    nb = space.dim + 1
    nodes = []
    for k in range(nb):
        sel = owner[opp == k]
        if sel.size == 0:
            continue
        loc = [i for i in range(nb) if i != k]
        loc += [nb + e for e, (i, j) in enumerate(space.pairs) if i != k and j != k]
        nodes.append(space.cell_nodes[sel][:, loc].ravel())
    return np.unique(np.concatenate(nodes))


# --------------------------------------------------------------------------
# backward-facing step (reference demo / bench problem)
# --------------------------------------------------------------------------


def bfs_space(level):
    return fem.TaylorHoodSpace(*fem.lshape_mesh(level))


def bfs_boundary(space):
    X = space.node_coords
    bnd = _on_boundary_nodes(space)
    xb = X[bnd]
    inlet = np.isclose(xb[:, 0], -1.0)
    outlet = np.isclose(xb[:, 0], 5.0)
    dirichlet = bnd[~outlet | (np.isclose(xb[:, 1], -1.0) | np.isclose(xb[:, 1], 1.0))]
    vals = np.zeros((dirichlet.size, 2))
    xd = X[dirichlet]
    on_in = np.isclose(xd[:, 0], -1.0)
    vals[on_in, 0] = 4.0 * xd[on_in, 1] * (1.0 - xd[on_in, 1])
    return dirichlet, vals, inlet, outlet, bnd


def backward_facing_step(level=4, nu=0.02, variant="BRM1", wind=None, idt=0.0,
                         stabilise=False, newton=False, pcdr=False):
    """One linearised step of the reference BFS problem around ``wind``
    ([n2, 2] nodal values; default: zero wind = Stokes, the first Picard step
    from the zero initial guess of the demo)."""
    space = bfs_space(level)
    dirichlet, vals, _, _, _ = bfs_boundary(space)
    if wind is None:
        wind = np.zeros((space.n2, 2))
    vx = space.verts[:, 0]
    pcd_mask = np.isclose(vx, -1.0) if variant == "BRM1" else np.isclose(vx, 5.0)
    return _build(f"bfs_l{level}", space, nu, variant, wind, dirichlet, vals, pcd_mask,
                  inlet_mask_fn=lambda x: np.isclose(x[:, 0], -1.0), idt=idt,
                  stabilise=stabilise, newton=newton, pcdr=pcdr), space


# --------------------------------------------------------------------------
# box domains
# --------------------------------------------------------------------------


def recirculating_wind(X):
    """Analytic enclosed-flow wind (2y(1-x^2), -2x(1-y^2)) on [-1,1]^2 mapped to
    the unit square/cube; third component zero in 3D."""
    x = 2.0 * X[:, 0] - 1.0
    y = 2.0 * X[:, -1] - 1.0 if X.shape[1] == 2 else 2.0 * X[:, 1] - 1.0
    W = np.zeros_like(X)
    if X.shape[1] == 2:
        W[:, 0] = 2.0 * y * (1 - x * x)
        W[:, 1] = -2.0 * x * (1 - y * y)
    else:
        z = 2.0 * X[:, 2] - 1.0
        W[:, 0] = 2.0 * z * (1 - x * x)
        W[:, 2] = -2.0 * x * (1 - z * z)
        W[:, 1] = 0.0 * y
    return W


def lid_driven_cavity(n, dim=2, nu=0.02, variant="BRM2", wind="recirculating", stabilise=True,
                      newton=False):
    """Unit square/cube, lid (top face) moving with unit speed in x, no-slip
    elsewhere.  Enclosed flow has no inlet/outlet, so the PCD Dirichlet set is a
    single pinned pressure dof (vertex 0) which makes Ap non-singular."""
    if dim == 2:
        space = fem.TaylorHoodSpace(*fem.rectangle_mesh(n, n))
    else:
        space = fem.TaylorHoodSpace(*fem.box_mesh(n, n, n))
    X = space.node_coords
    bnd = _on_boundary_nodes(space)
    vals = np.zeros((bnd.size, dim))
    top = np.isclose(X[bnd, dim - 1], 1.0)
    vals[top, 0] = 1.0
    W = recirculating_wind(X) if isinstance(wind, str) else wind
    pcd_mask = np.zeros(space.n1, dtype=bool)
    pcd_mask[0] = True
    return _build(f"cavity{dim}d_n{n}", space, nu, variant, W, bnd, vals, pcd_mask,
                  inlet_mask_fn=None, stabilise=stabilise, newton=newton), space


def channel(nx, ny, nz=None, length=4.0, nu=0.02, variant="BRM1", stabilise=True):
    """Box channel [0,L]x[0,1](x[0,1]): parabolic inflow at x=0, natural outflow
    at x=L, no-slip walls; wind = the Poiseuille profile (exact solution)."""
    if nz is None:
        space = fem.TaylorHoodSpace(*fem.rectangle_mesh(nx, ny, 0.0, 0.0, length, 1.0))
    else:
        space = fem.TaylorHoodSpace(*fem.box_mesh(nx, ny, nz, (length, 1.0, 1.0)))
    d = space.dim
    X = space.node_coords
    prof = 4.0 * X[:, 1] * (1 - X[:, 1])
    if d == 3:
        prof = prof * 4.0 * X[:, 2] * (1 - X[:, 2])
    W = np.zeros_like(X)
    W[:, 0] = prof
    bnd = _on_boundary_nodes(space)
    xb = X[bnd]
    walls = np.isclose(xb[:, 1], 0.0) | np.isclose(xb[:, 1], 1.0)
    if d == 3:
        walls |= np.isclose(xb[:, 2], 0.0) | np.isclose(xb[:, 2], 1.0)
    outlet = np.isclose(xb[:, 0], length) & ~walls
    dirichlet = bnd[~outlet]
    vals = np.zeros((dirichlet.size, d))
    xd = X[dirichlet]
    on_in = np.isclose(xd[:, 0], 0.0)
    vals[on_in, 0] = W[dirichlet][on_in, 0]
    vx = space.verts[:, 0]
    pcd_mask = np.isclose(vx, 0.0) if variant == "BRM1" else np.isclose(vx, length)
    return _build(f"channel{d}d_{nx}x{ny}" + (f"x{nz}" if nz else ""), space, nu, variant, W,
                  dirichlet, vals, pcd_mask, inlet_mask_fn=lambda x: np.isclose(x[:, 0], 0.0),
                  stabilise=stabilise), space
