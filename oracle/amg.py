"""Smoothed-aggregation AMG, CPU restatement -- TEST INFRASTRUCTURE.

The reference delegates Ap^-1 and the velocity-block solve of its "iterative"
set-up to hypre BoomerAMG (demo_navier-stokes-pcd.py:153-160), which is neither
vendored nor reproducible bit-wise.  BASELINE.json's north_star replaces it, on
both sides of the comparison, by a smoothed-aggregation V-cycle whose smoother,
restriction and prolongation are SpMV-class operations.  This module restates
that algorithm (Vanek/Mandel/Brezina) so that the product's hierarchy and
V-cycle can be checked to rounding:

  strength     |a_ij| >= theta * sqrt(|a_ii| |a_jj|)            (i != j)
  aggregation  greedy three-phase (root + strong neighbours; leftovers join the
               strongest neighbouring aggregate; remaining form own aggregates);
               rows without strong neighbours (Dirichlet rows) stay un-aggregated
  tentative    T[i, agg(i)] = 1/sqrt(|agg|)
  prolongator  P = (I - (4/3)/rho * D^-1 A) T, rho = power-iteration estimate of
               rho(D^-1 A) (20 steps, fixed start vector) times 1.1; entries below
               p_trunc * max|row| are dropped and rows rescaled to their row sum
  coarse op    A_c = P^T A P, then entries below coarse_drop*sqrt(|a_ii||a_jj|) are
               lumped onto the diagonal (SA coarse stencils of P2 operators in 3D
               have ~200 entries per row, ~90 % of them negligible)
  smoother     Chebyshev-Jacobi of ``smooth_steps`` steps on [rho/ratio, rho] (one
               step = damped Jacobi), same recurrence as the Mp solve
  coarsest     dense inverse
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from .petsc_algos import chebyshev_jacobi


def start_vector(n):
    """Deterministic pseudo-random start vector shared with the product
    (multiplicative hash of the index), values in [0.5, 1.5)."""
    i = np.arange(n, dtype=np.uint64)
    h = (i * np.uint64(2654435761) + np.uint64(12345)) & np.uint64(0xFFFFFFFF)
    return 0.5 + h.astype(np.float64) / 4294967296.0


def estimate_rho(A, dinv, steps=20, safety=1.1):
    v = start_vector(A.shape[0])
    v /= np.linalg.norm(v)
    rho = 0.0
    for _ in range(steps):
        w = dinv * (A @ v)
        rho = float(np.linalg.norm(w))
        if rho == 0.0:
            return 1.0
        v = w / rho
    return safety * rho


def strength_graph(A, theta):
    """CSR boolean graph of strong off-diagonal connections."""
    A = A.tocsr()
    d = np.abs(A.diagonal())
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    cols = A.indices
    strong = (rows != cols) & (np.abs(A.data) >= theta * np.sqrt(d[rows] * d[cols])) & (A.data != 0.0)
    S = sp.csr_matrix((np.abs(A.data[strong]), (rows[strong], cols[strong])), shape=A.shape)
    S.sort_indices()
    return S


def aggregate_greedy(S, use_c=None):
    """Three-phase greedy aggregation on the strength graph S (CSR, values =
    |a_ij|).  Returns agg[n] (aggregate id or -1) and the number of aggregates.
    For large graphs the identical loop in oracle/pcd_ref.c is used when the C
    library is built (tests/test_oracle.py checks the two against each other)."""
    n = S.shape[0]
    if use_c is None:
        use_c = n > 20000
    if use_c:
        from . import cref
        return cref.aggregate_greedy(S)
    ip, ix, vals = S.indptr, S.indices, S.data
    agg = np.full(n, -1, dtype=np.int64)
    nagg = 0
    has_nb = np.diff(ip) > 0
    # phase 1: root i with all strong neighbours still free
    for i in range(n):
        if agg[i] != -1 or not has_nb[i]:
            continue
        nb = ix[ip[i]:ip[i + 1]]
        if np.all(agg[nb] == -1):
            agg[i] = nagg
            agg[nb] = nagg
            nagg += 1
    # phase 2: leftovers join the aggregate of their strongest phase-1 neighbour
    agg1 = agg.copy()
    for i in range(n):
        if agg[i] != -1 or not has_nb[i]:
            continue
        best, bestv = -1, -1.0
        for k in range(ip[i], ip[i + 1]):
            j = ix[k]
            if agg1[j] != -1 and vals[k] > bestv:
                best, bestv = agg1[j], vals[k]
        if best != -1:
            agg[i] = best
    # phase 3: whatever is left forms aggregates with its free neighbours
    for i in range(n):
        if agg[i] != -1 or not has_nb[i]:
            continue
        agg[i] = nagg
        for k in range(ip[i], ip[i + 1]):
            j = ix[k]
            if agg[j] == -1 and has_nb[j]:
                agg[j] = nagg
        nagg += 1
    return agg, nagg


def tentative_prolongator(agg, nagg):
    n = agg.size
    rows = np.flatnonzero(agg >= 0)
    cols = agg[rows]
    counts = np.bincount(cols, minlength=nagg).astype(np.float64)
    vals = 1.0 / np.sqrt(counts[cols])
    T = sp.csr_matrix((vals, (rows, cols)), shape=(n, nagg))
    return T


@dataclass
class Level:
    A: sp.csr_matrix
    dinv: np.ndarray
    rho: float
    P: sp.csr_matrix | None = None     # to this level from the next coarser one
    R: sp.csr_matrix | None = None


@dataclass
class Hierarchy:
    levels: list = field(default_factory=list)
    coarse_inv: np.ndarray | None = None
    smooth_steps: int = 2
    eig_ratio: float = 10.0

    def operator_complexity(self):
        return sum(l.A.nnz for l in self.levels) / self.levels[0].A.nnz

    def _smooth(self, lvl, b):
        return chebyshev_jacobi(lvl.A, lvl.dinv, b, lvl.rho / self.eig_ratio, lvl.rho, self.smooth_steps)

    def vcycle(self, b, k=0):
        """One V-cycle with zero initial guess on level k."""
        lvl = self.levels[k]
        if k == len(self.levels) - 1:
            return self.coarse_inv @ b
        x = self._smooth(lvl, b)                 # pre-smoothing from zero guess
        r = b - lvl.A @ x
        x = x + lvl.P @ self.vcycle(lvl.R @ r, k + 1)
        r = b - lvl.A @ x
        return x + self._smooth(lvl, r)          # post-smoothing on the correction equation

    def __call__(self, b):
        return self.vcycle(b)


def filter_lumped(A, drop):
    """Sparsify a coarse operator: off-diagonal entries with
    |a_ij| < drop * sqrt(|a_ii| |a_jj|) are removed and added to the diagonal of
    their row (row sums, hence the action on constants, are preserved)."""
    if drop <= 0.0:
        return A
    A = A.tocsr()
    d = np.abs(A.diagonal())
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    cols = A.indices
    small = (rows != cols) & (np.abs(A.data) < drop * np.sqrt(d[rows] * d[cols]))
    lump = np.bincount(rows[small], weights=A.data[small], minlength=A.shape[0])
    data = A.data + np.where(rows == cols, lump[rows], 0.0)
    keep = ~small
    out = sp.csr_matrix((data[keep], (rows[keep], cols[keep])), shape=A.shape)
    out.sort_indices()
    return out


def truncate_prolongator(P, trunc):
    """Drop entries below trunc * max|row| and rescale every row to its former row
    sum (constants stay in the range of P)."""
    if trunc <= 0.0:
        return P
    n = P.shape[0]
    rows = np.repeat(np.arange(n), np.diff(P.indptr))
    rmax = np.zeros(n)
    np.maximum.at(rmax, rows, np.abs(P.data))
    rs0 = np.bincount(rows, weights=P.data, minlength=n)
    keep = np.abs(P.data) >= trunc * rmax[rows]
    rs1 = np.bincount(rows[keep], weights=P.data[keep], minlength=n)
    scale = np.where(rs1 != 0.0, rs0 / np.where(rs1 != 0.0, rs1, 1.0), 1.0)
    out = sp.csr_matrix((scale[rows[keep]] * P.data[keep], (rows[keep], P.indices[keep])), shape=P.shape)
    out.sort_indices()
    return out


def _block_of(begins, n):
    """Block (rank) index of every row for ownership offsets ``begins``."""
    return np.searchsorted(np.asarray(begins[1:]), np.arange(n), side="right")


def build_hierarchy(A, theta=0.08, max_levels=12, coarse_size=400, smooth_steps=2,
                    eig_ratio=10.0, omega_scale=4.0 / 3.0, coarse_drop=0.0, p_trunc=0.2, blocks=None,
                    replicate_size=0):
    """``blocks``: ownership offsets [0, n_1, ..., n] of a row partition.  With more
    than one block the aggregation and the prolongator smoothing are block local
    (no aggregate crosses a block boundary, P = T - omega D^-1 A_bd T with A_bd the
    block-diagonal part), exactly what the multi-rank library does; the Galerkin
    product uses the full A.  One block = the plain serial algorithm.  Levels with at most
    ``replicate_size`` rows are coarsened serially again (the library gathers them on
    every rank)."""
    H = Hierarchy(smooth_steps=smooth_steps, eig_ratio=eig_ratio)
    A = sp.csr_matrix(A)
    A.sort_indices()
    begins = [0, A.shape[0]] if blocks is None else [int(b) for b in blocks]
    H.begins = []
    while True:
        n = A.shape[0]
        nb = len(begins) - 1
        diag = A.diagonal()
        dinv = np.where(diag != 0.0, 1.0 / np.where(diag != 0.0, diag, 1.0), 0.0)
        if nb == 1:
            Abd = A
            rho = estimate_rho(A, dinv)
        else:
            blk = _block_of(begins, n)
            coo = A.tocoo()
            same = blk[coo.row] == blk[coo.col]
            Abd = sp.csr_matrix((coo.data[same], (coo.row[same], coo.col[same])), shape=A.shape)
            Abd.sort_indices()
            rho = max(estimate_rho(Abd[b0:b1, b0:b1].tocsr(), dinv[b0:b1])
                      for b0, b1 in zip(begins[:-1], begins[1:]) if b1 > b0)
        lvl = Level(A=A, dinv=dinv, rho=rho)
        H.levels.append(lvl)
        H.begins.append(list(begins))
        if n <= coarse_size or len(H.levels) >= max_levels:
            break
        S = strength_graph(Abd, theta * 0.5 ** (len(H.levels) - 1))
        if nb == 1:
            agg, nagg = aggregate_greedy(S)
            cbegins = [0, nagg]
        else:
            agg = np.full(n, -1, dtype=np.int64)
            cbegins = [0]
            for b0, b1 in zip(begins[:-1], begins[1:]):
                a, na = aggregate_greedy(S[b0:b1, b0:b1].tocsr())
                agg[b0:b1] = np.where(a >= 0, a + cbegins[-1], -1)
                cbegins.append(cbegins[-1] + na)
            nagg = cbegins[-1]
        if nagg == 0 or nagg >= n:
            break
        T = tentative_prolongator(agg, nagg)
        omega = omega_scale / rho
        P = (T - omega * (sp.diags(dinv) @ (Abd @ T))).tocsr()
        P.sort_indices()
        P = truncate_prolongator(P, p_trunc)
        R = P.T.tocsr()
        R.sort_indices()
        Ac = (R @ (A @ P)).tocsr()
        Ac.sort_indices()
        Ac = filter_lumped(Ac, coarse_drop)
        lvl.P, lvl.R = P, R
        A = Ac
        begins = cbegins
        if len(begins) > 2 and begins[-1] <= replicate_size:
            # the library replicates small levels on every rank and continues serially
            begins = [0, begins[-1]]
    H.coarse_inv = np.linalg.inv(H.levels[-1].A.toarray())
    return H


def build_hierarchy_kron(A, bs=1, blocks=None, S=None, **kw):
    """Hierarchy of A = S (x) I_bs (interleaved components) arranged the way the
    library does it in Kronecker mode: the scalar operator S is coarsened (node blocks
    for a multi-rank partition) and every level is expanded back to the full block.
    `S` may be handed in (large systems: extracting it from A and expanding level 0 again
    would only cost memory); level 0 then is A itself."""
    A = sp.csr_matrix(A)
    if bs == 1:
        return build_hierarchy(A, blocks=blocks, **kw)
    given = S is not None
    S = sp.csr_matrix(S) if given else A[::bs, :][:, ::bs].tocsr()
    Hs = build_hierarchy(S, blocks=None if blocks is None else [b // bs for b in blocks], **kw)
    eye = sp.identity(bs, format="csr")
    H = Hierarchy(smooth_steps=Hs.smooth_steps, eig_ratio=Hs.eig_ratio)
    for k, l in enumerate(Hs.levels):
        H.levels.append(Level(A=A if (given and k == 0) else sp.kron(l.A, eye, format="csr"), dinv=np.repeat(l.dinv, bs), rho=l.rho,
                              P=None if l.P is None else sp.kron(l.P, eye, format="csr"),
                              R=None if l.R is None else sp.kron(l.R, eye, format="csr")))
    H.coarse_inv = np.kron(Hs.coarse_inv, np.eye(bs))
    H.begins = Hs.begins
    return H


# --------------------------------------------------------------------------
# numeric refresh with frozen prolongators (prototype of the device-side refresh)
# --------------------------------------------------------------------------


def galerkin_plan(A, P):
    """(pattern of A_c, W) with ``A_c.data == W @ A.data`` for every matrix that has A's pattern:
    the Galerkin product P^T A P with P frozen is linear in the values of A, entry (I, J) of A_c
    collecting p_iI * a_ij * p_jJ over the fine entries (i, j).  W has one row per stored entry of
    A_c and one column per stored entry of A, so a value refresh of the coarse operator is ONE
    SpMV-class pass (deterministic, no atomics) -- the formulation the device-side refresh uses.
    The pattern of A_c is the structural product (no entry is lost to cancellation)."""
    A = sp.csr_matrix(A)
    P = sp.csr_matrix(P)
    A.sort_indices()
    P.sort_indices()
    n = A.shape[0]
    rows = np.repeat(np.arange(n), np.diff(A.indptr))
    cols = A.indices
    ci = np.diff(P.indptr)[rows]                      # |P row i| per fine entry
    cj = np.diff(P.indptr)[cols]                      # |P row j|
    nt = ci * cj                                      # terms per fine entry
    e = np.repeat(np.arange(A.nnz), nt)               # fine entry of every term
    t = np.arange(nt.sum()) - np.repeat(np.cumsum(nt) - nt, nt)       # running index inside the entry
    a = t // np.repeat(cj, nt)                        # which entry of P row i
    b = t % np.repeat(cj, nt)                         # which entry of P row j
    pi = P.indptr[rows[e]] + a
    pj = P.indptr[cols[e]] + b
    I, J, coef = P.indices[pi], P.indices[pj], P.data[pi] * P.data[pj]
    nc = P.shape[1]
    pat = sp.csr_matrix((np.ones(I.size), (I, J)), shape=(nc, nc))   # duplicates summed: structural pattern
    pat.sort_indices()
    # position of (I, J) in the CSR of the pattern
    key = I.astype(np.int64) * nc + J
    prow = np.repeat(np.arange(nc), np.diff(pat.indptr))
    pkey = prow.astype(np.int64) * nc + pat.indices
    pos = np.searchsorted(pkey, key)
    W = sp.csr_matrix((coef, (pos, e)), shape=(pat.nnz, A.nnz))
    W.sum_duplicates()
    return pat, W


def refresh_plans(H):
    """Plans of every level transition of a hierarchy built with ``coarse_drop = 0``."""
    return [galerkin_plan(l.A, l.P) for l in H.levels[:-1]]


def refresh_hierarchy(H, A_new, plans, new_rho=True):
    """Hierarchy for new values on level 0 (same pattern), prolongators frozen, coarse operators
    recomputed through the plans.  ``new_rho``: re-estimate the smoother bounds on every level."""
    A = sp.csr_matrix(A_new)
    A.sort_indices()
    out = Hierarchy(smooth_steps=H.smooth_steps, eig_ratio=H.eig_ratio)
    for k, old in enumerate(H.levels):
        diag = A.diagonal()
        dinv = np.where(diag != 0.0, 1.0 / np.where(diag != 0.0, diag, 1.0), 0.0)
        out.levels.append(Level(A=A, dinv=dinv, rho=estimate_rho(A, dinv) if new_rho else old.rho, P=old.P, R=old.R))
        if k + 1 < len(H.levels):
            pat, W = plans[k]
            A = sp.csr_matrix((W @ A.data, pat.indices, pat.indptr), shape=pat.shape)
    out.coarse_inv = np.linalg.inv(out.levels[-1].A.toarray())
    return out
