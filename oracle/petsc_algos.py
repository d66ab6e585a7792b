"""Restatement (numpy, float64) of the PETSc algorithm chain FENaPack selects --
TEST INFRASTRUCTURE.  PARITY UNPINNED for the PETSc algorithms: PETSc is not vendored
in /root/reference, not installed here, and no release is pinned; the recurrences
below are restated from the published algorithms and are marked "recalled" where the
PETSc source decides a detail.  PINNED for the functions that
restate the reference's own Python (brm1/brm2/pcdr_* apply, build_rp, apply_bcs):
tests/test_reference_golden.py checks them against outputs of the reference's code
run in this container (tests/golden/make_reference_golden.py).

  reference call site                                   restated here
  fenapack/preconditioners.py:124-135  (PCDPC_BRM1.apply)   brm1_apply
  fenapack/preconditioners.py:158-169  (PCDPC_BRM2.apply)   brm2_apply
  fenapack/preconditioners.py:251-262, 284-297 (PCDRPC_*)   pcdr_brm1_apply, pcdr_brm2_apply
  fenapack/field_split_backend.py:142-166 (_build_approx_Ap) build_rp
  fenapack/SubfieldBC.h:162-182        (VecSetValues INSERT) apply_bcs
  fenapack/field_split.py:52-57        (GMRES, right PC,    gmres_right / fgmres,
                                        fieldsplit SCHUR/UPPER) fieldsplit_upper_apply
  demo_navier-stokes-pcd.py:161-165    (chebyshev+jacobi)   chebyshev_jacobi
  demo_navier-stokes-pcd.py:153-160    (richardson+AMG)     richardson
  fenapack/preconditioners.py:43-49    (preonly+cholesky)   direct_solver (scipy splu)
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# --------------------------------------------------------------------------
# inner solvers
# --------------------------------------------------------------------------


def direct_solver(A):
    """PREONLY + LU/Cholesky (fenapack/preconditioners.py:43-49, field_split.py:96-98)."""
    lu = spla.splu(sp.csc_matrix(A))
    return lu.solve


def chebyshev_coefficients(emin, emax, steps):
    """Scalars of PETSc's KSPCHEBYSHEV recurrence (recalled).
    Returns (s, [omega_1 .. omega_{steps-1}])."""
    s = 2.0 / (emax + emin)
    alpha = 1.0 - s * emin
    mu = 1.0 / alpha
    omegaprod = 2.0 / alpha
    c0, c1 = 1.0, mu
    omegas = []
    for _ in range(steps - 1):
        c2 = 2.0 * mu * c1 - c0
        omegas.append(omegaprod * c1 / c2)
        c0, c1 = c1, c2
    return s, omegas


def chebyshev_jacobi(A, dinv, b, emin, emax, steps):
    """KSPCHEBYSHEV + PCJACOBI, zero initial guess, fixed number of steps
    (``steps`` = number of Jacobi applications = degree of the polynomial in
    D^-1 A; the reference sets ksp_max_it 5, demo_navier-stokes-pcd.py:162).
    No norms are evaluated: with bounds [0.5, 2] the contraction never reaches
    PETSc's default rtol within 5 steps, so the norms never change the result."""
    s, omegas = chebyshev_coefficients(emin, emax, steps)
    p0 = np.zeros_like(b)
    p1 = s * (dinv * b)
    for omega in omegas:
        r = b - A @ p1
        p2 = (1.0 - omega) * p0 + omega * p1 + (omega * s) * (dinv * r)
        p0, p1 = p1, p2
    return p1


def richardson(A, precond, b, max_it):
    """KSPRICHARDSON (scale 1), zero initial guess, ``max_it`` preconditioner
    applications: x1 = B b; x_{k+1} = x_k + B (b - A x_k)."""
    x = precond(b)
    for _ in range(max_it - 1):
        x = x + precond(b - A @ x)
    return x


def pcg(A, precond, b, max_it, rtol=0.0):
    """Preconditioned CG (KSPCG), zero initial guess.  Stops on the
    preconditioned residual norm <= rtol*||B b|| or after max_it iterations."""
    x = np.zeros_like(b)
    r = b.copy()
    z = precond(r)
    p = z.copy()
    rz = float(r @ z)
    z0 = np.sqrt(float(z @ z))
    its = 0
    for its in range(1, max_it + 1):
        Ap = A @ p
        alpha = rz / float(p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        z = precond(r)
        if rtol > 0.0 and np.sqrt(float(z @ z)) <= rtol * z0:
            break
        rz_new = float(r @ z)
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, its


# --------------------------------------------------------------------------
# the Schur-complement approximations (the reference's own arithmetic)
# --------------------------------------------------------------------------


def apply_bcs(vec, bc_idx, bc_val):
    """SubfieldBC::apply: vec[idx] = val (INSERT_VALUES), SubfieldBC.h:162-182."""
    vec[bc_idx] = bc_val
    return vec


def brm1_apply(x, solve_Ap, Kp, solve_Mp, bc_idx, bc_val):
    """y = -Mp^-1 (x + Kp Ap^-1 bc(x));  preconditioners.py:124-135, step by step."""
    z = x.copy()                       # x.copy(result=z)
    apply_bcs(z, bc_idx, bc_val)       # bcs_applier(z)
    y = solve_Ap(z)                    # ksp_Ap.solve(z, y)
    z = Kp @ y                         # mat_Kp.mult(y, z)
    z = z + x                          # z.axpy(1.0, x)
    y = solve_Mp(z)                    # ksp_Mp.solve(z, y)
    return -y                          # y.scale(-1.0)


def brm2_apply(x, solve_Ap, Kp, solve_Mp, bc_idx, bc_val):
    """y = -(I + Ap^-1 bc(Kp .)) Mp^-1 x;  preconditioners.py:158-169."""
    y = solve_Mp(x)                    # ksp_Mp.solve(x, y)
    z0 = y.copy()                      # y.copy(result=z0)
    z1 = Kp @ z0                       # mat_Kp.mult(z0, z1)
    apply_bcs(z1, bc_idx, bc_val)      # bcs_applier(z1)
    z0 = solve_Ap(z1)                  # ksp_Ap.solve(z1, z0)
    y = y + z0                         # y.axpy(1.0, z0)
    return -y                          # y.scale(-1.0)


def build_rp(Bt, mu_diag):
    """Approximate pressure Laplacian of the PCDR variants: Rp = B diag(Mu)^-1 B^T,
    built as (D^-1/2 Bt)^T (D^-1/2 Bt) exactly as PCDInterface._build_approx_Ap does
    (fenapack/field_split_backend.py:142-166)."""
    d = np.sqrt(np.abs(1.0 / mu_diag))
    S = sp.diags(d) @ sp.csr_matrix(Bt)
    Rp = (S.T @ S).tocsr()
    Rp.sort_indices()
    return Rp


def pcdr_brm1_apply(x, solve_Ap, Kp, solve_Mp, solve_Rp, bc_idx, bc_val):
    """y = -Rp^-1 x - Mp^-1 (I + Kp Ap^-1) x;  preconditioners.py:251-262, step by step."""
    z = x.copy()
    apply_bcs(z, bc_idx, bc_val)
    y = solve_Ap(z)
    z = Kp @ y
    z = z + x
    y = solve_Mp(z)
    z = solve_Rp(x)                    # ksp_Rp.solve(x, z)
    y = y + z                          # y.axpy(1.0, z)
    return -y


def pcdr_brm2_apply(x, solve_Ap, Kp, solve_Mp, solve_Rp, bc_idx, bc_val):
    """y = -Rp^-1 x - (I + Ap^-1 Kp) Mp^-1 x;  preconditioners.py:284-297."""
    y = solve_Mp(x)
    z0 = y.copy()
    z1 = Kp @ z0
    apply_bcs(z1, bc_idx, bc_val)
    z0 = solve_Ap(z1)
    y = y + z0
    z0 = solve_Rp(x)                   # ksp_Rp.solve(x, z0)
    y = y + z0
    return -y


def fieldsplit_upper_apply(x_u, x_p, schur_apply, A01, solve_A00):
    """PCFIELDSPLIT, SCHUR factorisation, UPPER (field_split.py:54-57; PETSc
    fieldsplit.c recalled):
        y_p = S^-1 x_p ;  y_u = A00^-1 (x_u - A01 y_p)."""
    y_p = schur_apply(x_p)
    y_u = solve_A00(x_u - A01 @ y_p)
    return y_u, y_p


# --------------------------------------------------------------------------
# outer Krylov method
# --------------------------------------------------------------------------


def _givens_update(H, cs, sn, g, j):
    """Apply previous rotations to column j of H, form the new one, update g."""
    for i in range(j):
        t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
        H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
        H[i, j] = t
    a, b = H[j, j], H[j + 1, j]
    rho = np.hypot(a, b)
    cs[j], sn[j] = (1.0, 0.0) if rho == 0.0 else (a / rho, b / rho)
    H[j, j] = rho
    H[j + 1, j] = 0.0
    g[j + 1] = -sn[j] * g[j]
    g[j] = cs[j] * g[j]


def fgmres(A, precond, b, rtol=1e-6, atol=1e-50, restart=150, max_it=10000, flexible=True,
           monitor=None):
    """Right-preconditioned restarted GMRES with classical Gram-Schmidt (no
    refinement -- PETSc's default), zero initial guess, convergence on the
    recurrence estimate of the true residual norm: ||r|| <= max(rtol*||b||, atol)
    (KSPGMRES as configured at field_split.py:52-53 and demo:146-148).

    flexible=True stores Z_j = M^-1 v_j (FGMRES); flexible=False rebuilds the
    update with one extra preconditioner application per cycle (PETSc's right
    GMRES).  Both give identical iterates when ``precond`` is a fixed linear
    operator.  Returns (x, iterations, residual history, number of PC applies)."""
    matvec = (lambda v: A @ v) if not callable(A) else A
    n = b.size
    x = np.zeros(n)
    bnorm = float(np.linalg.norm(b))
    tol = max(rtol * bnorm, atol)
    hist = [bnorm]
    its = 0
    napply = 0
    if bnorm <= tol:
        return x, 0, hist, 0
    r = b.copy()
    beta = bnorm
    while its < max_it:
        m = restart
        V = np.zeros((m + 1, n))
        Z = np.zeros((m, n)) if flexible else None
        H = np.zeros((m + 1, m))
        cs, sn, g = np.zeros(m), np.zeros(m), np.zeros(m + 1)
        V[0] = r / beta
        g[0] = beta
        j_done = 0
        converged = False
        for j in range(m):
            z = precond(V[j])
            napply += 1
            if flexible:
                Z[j] = z
            w = matvec(z)
            # classical Gram-Schmidt: all dots against the unmodified w, then one MAXPY
            h = V[: j + 1] @ w
            w = w - h @ V[: j + 1]
            hn = float(np.linalg.norm(w))
            H[: j + 1, j] = h
            H[j + 1, j] = hn
            if hn != 0.0:
                V[j + 1] = w / hn
            _givens_update(H, cs, sn, g, j)
            its += 1
            j_done = j + 1
            res = abs(g[j + 1])
            hist.append(res)
            if monitor is not None:
                monitor(its, res)
            if res <= tol or its >= max_it:
                converged = res <= tol
                break
        y = np.linalg.solve(np.triu(H[:j_done, :j_done]), g[:j_done])
        if flexible:
            x = x + y @ Z[:j_done]
        else:
            x = x + precond(y @ V[:j_done])
            napply += 1
        if converged or its >= max_it:
            break
        r = b - matvec(x)
        beta = float(np.linalg.norm(r))
        if beta <= tol:
            break
    return x, its, hist, napply


# --------------------------------------------------------------------------
# the assembled preconditioner of one PCDProblem
# --------------------------------------------------------------------------


class PCDPreconditioner:
    """The block-triangular PCD preconditioner of one ``PCDProblem`` with the
    reference's two inner-solver set-ups:

      ls="direct"     LU / Cholesky everywhere (the defaults at
                      preconditioners.py:43-49 and field_split.py:96-98)
      ls="iterative"  velocity: richardson x1 + AMG, Ap: richardson x2 + AMG,
                      Mp: chebyshev x5 + jacobi      (demo:153-165)
    ``amg_u`` / ``amg_p`` are callables b -> one V-cycle (oracle.amg.Hierarchy.vcycle).
    """

    def __init__(self, prob, ls="direct", amg_u=None, amg_p=None, cheb_steps=5,
                 ap_its=2, u_its=1, pcdr=False, amg_r=None, rp_its=1):
        self.prob = prob
        self.ls = ls
        self.pcdr = pcdr
        if pcdr:      # PCDR variants: Rp from the discrete gradient and the velocity mass diagonal
            self.Rp = build_rp(prob.A01, prob.mu_diag)
            self.solve_Rp = direct_solver(self.Rp) if ls == "direct" else \
                (lambda b: richardson(self.Rp, amg_r, b, rp_its))
        P00 = prob.P00 if prob.P00 is not None else prob.A00
        if ls == "direct":
            self.solve_A00 = direct_solver(P00)
            self.solve_Ap = direct_solver(prob.Ap)
            self.solve_Mp = direct_solver(prob.Mp)
        else:
            dinv = 1.0 / prob.Mp.diagonal()
            emin, emax = prob.cheb_bounds
            self.solve_Mp = lambda b: chebyshev_jacobi(prob.Mp, dinv, b, emin, emax, cheb_steps)
            self.solve_Ap = lambda b: richardson(prob.Ap, amg_p, b, ap_its)
            self.solve_A00 = lambda b: richardson(P00, amg_u, b, u_its)

    def schur_apply(self, x_p):
        if self.pcdr:
            f = pcdr_brm1_apply if self.prob.variant == "BRM1" else pcdr_brm2_apply
            return f(x_p, self.solve_Ap, self.prob.Kp, self.solve_Mp, self.solve_Rp, self.prob.bc_idx,
                     self.prob.bc_val)
        f = brm1_apply if self.prob.variant == "BRM1" else brm2_apply
        return f(x_p, self.solve_Ap, self.prob.Kp, self.solve_Mp, self.prob.bc_idx, self.prob.bc_val)

    def apply_split(self, x_u, x_p):
        return fieldsplit_upper_apply(x_u, x_p, self.schur_apply, self.prob.A01, self.solve_A00)

    def __call__(self, x):
        nu_ = self.prob.n_u
        y_u, y_p = self.apply_split(x[:nu_], x[nu_:])
        return np.concatenate([y_u, y_p])
