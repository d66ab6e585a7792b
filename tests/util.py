"""Shared helpers of the parity tests: build a device context from an oracle
``PCDProblem`` and the matching oracle preconditioner."""
import numpy as np
import scipy.sparse as sp

from fenapack_b200 import capi
from oracle import amg as oamg
from oracle import petsc_algos as pa

ITERATIVE_OPTIONS = {
    # the reference's "iterative" set-up, demo_navier-stokes-pcd.py:153-165
    "fieldsplit_u_ksp_type": "richardson",
    "fieldsplit_u_ksp_max_it": 1,
    "fieldsplit_u_pc_type": "hypre",
    "fieldsplit_u_pc_hypre_type": "boomeramg",
    "fieldsplit_p_PCD_Ap_ksp_type": "richardson",
    "fieldsplit_p_PCD_Ap_ksp_max_it": 2,
    "fieldsplit_p_PCD_Ap_pc_type": "hypre",
    "fieldsplit_p_PCD_Ap_pc_hypre_type": "boomeramg",
    "fieldsplit_p_PCD_Mp_ksp_type": "chebyshev",
    "fieldsplit_p_PCD_Mp_ksp_max_it": 5,
    "fieldsplit_p_PCD_Mp_pc_type": "jacobi",
    "ksp_gmres_restart": 150,
    "ksp_rtol": 1e-6,
}


def make_context(prob, extra_options=None, device=0, pcdr=False):
    ctx = capi.Context(device)
    opts = dict(ITERATIVE_OPTIONS)
    opts["fieldsplit_p_pc_python_type"] = ("fenapack.PCDRPC_" if pcdr else "fenapack.PCDPC_") + prob.variant
    if pcdr:        # demo_unsteady-navier-stokes-pcdr.py:167-170
        opts.update({"fieldsplit_p_PCD_Rp_ksp_type": "richardson", "fieldsplit_p_PCD_Rp_ksp_max_it": 1,
                     "fieldsplit_p_PCD_Rp_pc_type": "hypre", "fieldsplit_p_PCD_Rp_pc_hypre_type": "boomeramg"})
    opts["fieldsplit_p_PCD_Mp_ksp_chebyshev_eigenvalues"] = "%r, %r" % tuple(prob.cheb_bounds)
    opts.update(extra_options or {})
    ctx.set_options(opts)
    ctx.set_layout(prob.n_u, prob.n_p)
    ctx.set_matrix(capi.MAT_A00, prob.A00)
    ctx.set_matrix(capi.MAT_A01, prob.A01)
    ctx.set_matrix(capi.MAT_A10, prob.A10)
    ctx.set_matrix(capi.MAT_AP, prob.Ap)
    ctx.set_matrix(capi.MAT_MP, prob.Mp)
    ctx.set_matrix(capi.MAT_KP, prob.Kp)
    if prob.P00 is not None:
        ctx.set_matrix(capi.MAT_P00, prob.P00)
    ctx.set_bc(prob.bc_idx, prob.bc_val)
    if pcdr:
        ctx.set_mu_diag(prob.mu_diag)
    if prob.is_u is not None:
        ctx.set_index_sets(prob.is_u, prob.is_p)
    ctx.setup()
    return ctx


def oracle_hierarchy_from_device(ctx, which, smooth_steps=2, eig_ratio=10.0):
    """The oracle V-cycle (oracle/amg.py) on the hierarchy the library built."""
    levels, cinv = ctx.amg_hierarchy(which)
    bs = ctx.block_size(which)
    if bs > 1:      # Kronecker mode: the library coarsens S of A = S (x) I_bs; expand for the oracle
        eye = sp.identity(bs, format="csr")
        levels = [{k: (sp.kron(v, eye, format="csr") if k != "rho" else v) for k, v in e.items()} for e in levels]
        cinv = np.kron(cinv, np.eye(bs))
    H = oamg.Hierarchy(smooth_steps=smooth_steps, eig_ratio=eig_ratio)
    for e in levels:
        A = e["A"]
        d = A.diagonal()
        dinv = np.where(d != 0, 1.0 / np.where(d != 0, d, 1.0), 0.0)
        H.levels.append(oamg.Level(A=A, dinv=dinv, rho=e["rho"], P=e.get("P"), R=e.get("R")))
    H.coarse_inv = cinv
    return H


def oracle_preconditioner(prob, ctx, pcdr=False):
    Hu = oracle_hierarchy_from_device(ctx, capi.MAT_A00)
    Hp = oracle_hierarchy_from_device(ctx, capi.MAT_AP)
    if pcdr:
        Hr = oracle_hierarchy_from_device(ctx, capi.MAT_RP)
        return pa.PCDPreconditioner(prob, "iterative", amg_u=Hu, amg_p=Hp, pcdr=True, amg_r=Hr)
    return pa.PCDPreconditioner(prob, "iterative", amg_u=Hu, amg_p=Hp)


def relerr(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def oracle_hierarchy_like_library(A, bs=1, blocks=None, **kw):
    """The oracle's own hierarchy set-up arranged the way the library does it (Kronecker
    mode coarsens the scalar operator): oracle.amg.build_hierarchy_kron."""
    return oamg.build_hierarchy_kron(A, bs=bs, blocks=blocks, **kw)
