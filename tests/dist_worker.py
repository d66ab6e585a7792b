"""Multi-rank worker (one process per GPU, launched by torchrun from
tests/test_dist_gpu.py): row-partitioned operators, halo exchange, distributed AMG
and FGMRES through the C ABI, checked against the serial oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from fenapack_b200 import capi  # noqa: E402
from oracle import amg as oamg  # noqa: E402
from oracle import petsc_algos as pa  # noqa: E402
from oracle import problems  # noqa: E402
from util import ITERATIVE_OPTIONS, oracle_hierarchy_like_library, relerr  # noqa: E402


def split(n, world, align=1):
    b = [((n // align) * r // world) * align for r in range(world)] + [n]
    return b


def parity_check(rank, world, local, variant="BRM1", extra_options=None, p2p=None, repl=0, pcdr=False):
    """Row-partitioned library against the serial oracle on a small problem.  Collective over
    torch.distributed (already initialised).  Returns the measured discrepancies; raises on a
    violated tolerance.  pcdr: the PCDR variant on the unsteady BFS problem (Rp = Bt^T D^-1 Bt is
    assembled across the ranks)."""
    dim = 3
    if pcdr:
        dim = 2
        p0_, _ = problems.backward_facing_step(3, variant=variant, idt=5.0)
        x0 = pa.direct_solver(p0_.system_matrix())(p0_.rhs())
        prob, _ = problems.backward_facing_step(3, variant=variant, wind=x0[:p0_.n_u].reshape(-1, 2), idt=5.0,
                                                stabilise=True, pcdr=True)
    elif variant == "BRM1":
        prob, _ = problems.channel(12, 4, 6, variant=variant)
    else:
        prob, _ = problems.lid_driven_cavity(8, dim=3, variant="BRM2")
    ub, pb = split(prob.n_u, world, dim), split(prob.n_p, world)
    u0, u1, p0, p1 = ub[rank], ub[rank + 1], pb[rank], pb[rank + 1]

    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx = capi.Context(local, nccl_id=idt.cpu().numpy().tobytes(), rank=rank, nranks=world)
    opts = dict(ITERATIVE_OPTIONS)
    opts["fieldsplit_p_pc_python_type"] = ("fenapack.PCDRPC_" if pcdr else "fenapack.PCDPC_") + prob.variant
    if pcdr:        # demo_unsteady-navier-stokes-pcdr.py:167-170
        opts.update({"fieldsplit_p_PCD_Rp_ksp_type": "richardson", "fieldsplit_p_PCD_Rp_ksp_max_it": 1,
                     "fieldsplit_p_PCD_Rp_pc_type": "hypre", "fieldsplit_p_PCD_Rp_pc_hypre_type": "boomeramg"})
    opts["fieldsplit_p_PCD_Mp_ksp_chebyshev_eigenvalues"] = "%r, %r" % tuple(prob.cheb_bounds)
    opts.update(extra_options or {})
    ctx.set_options(opts)
    if p2p is not None:                    # exercise the peer-memory halo path as well
        ctx.set_option("fnp_halo_p2p", p2p)
    if repl:                               # replicated coarse tail of the hierarchies
        ctx.set_option("fieldsplit_u_pc_amg_replicate_size", repl)
        ctx.set_option("fieldsplit_p_PCD_Ap_pc_amg_replicate_size", repl)
    ctx.set_layout(u1 - u0, p1 - p0, u0, prob.n_u, p0, prob.n_p)
    P00 = prob.P00 if prob.P00 is not None else prob.A00
    mats = {capi.MAT_A00: (prob.A00, u0, u1), capi.MAT_A01: (prob.A01, u0, u1), capi.MAT_A10: (prob.A10, p0, p1),
            capi.MAT_AP: (prob.Ap, p0, p1), capi.MAT_MP: (prob.Mp, p0, p1), capi.MAT_KP: (prob.Kp, p0, p1)}
    if prob.P00 is not None:
        mats[capi.MAT_P00] = (prob.P00, u0, u1)
    for which, (A, r0, r1) in mats.items():
        ctx.set_matrix(which, A[r0:r1, :].tocsr())
    sel = (prob.bc_idx >= p0) & (prob.bc_idx < p1)
    ctx.set_bc(prob.bc_idx[sel] - p0, prob.bc_val[sel])
    if pcdr:
        ctx.set_mu_diag(prob.mu_diag[u0:u1])
    ctx.setup()

    out = {}
    rng = np.random.default_rng(0)
    # distributed SpMV of every operator
    worst = 0.0
    for which, (A, r0, r1) in mats.items():
        x = rng.standard_normal(A.shape[1])
        c0, c1 = (u0, u1) if A.shape[1] == prob.n_u else (p0, p1)
        y = ctx.spmv(which, x[c0:c1], r1 - r0)
        e = relerr(y, (A @ x)[r0:r1])
        assert e <= 1e-12, ("spmv", which, e)
        worst = max(worst, e)
    out["spmv_relerr"] = worst
    # the preconditioner against the oracle with the same block-local hierarchy
    bs = ctx.block_size(capi.MAT_P00 if prob.P00 is not None else capi.MAT_A00)
    assert bs == dim, "the Picard velocity block should be recognised as S (x) I_d"
    kw = {k[len("fieldsplit_u_pc_amg_"):]: int(v) for k, v in opts.items() if k == "fieldsplit_u_pc_amg_coarse_size"}
    kwp = {k[len("fieldsplit_p_PCD_Ap_pc_amg_"):]: int(v) for k, v in opts.items() if k == "fieldsplit_p_PCD_Ap_pc_amg_coarse_size"}
    Hu = oracle_hierarchy_like_library(P00, bs=bs, blocks=ub, replicate_size=repl, **kw)
    Hp = oamg.build_hierarchy(prob.Ap, blocks=pb, replicate_size=repl, **kwp)
    if pcdr:
        Rp = pa.build_rp(prob.A01, prob.mu_diag)
        v = rng.standard_normal(prob.n_p)
        out["rp_spmv_relerr"] = relerr(ctx.spmv(capi.MAT_RP, v[p0:p1], p1 - p0), (Rp @ v)[p0:p1])
        assert out["rp_spmv_relerr"] <= 1e-12, ("Rp", out["rp_spmv_relerr"])
        # the hierarchy of Rp from the values the library assembled (the cross-rank sum rounds differently from
        # the serial product, and aggregation decisions on a structured mesh are ties): gather its rows
        import scipy.sparse as sp
        parts = [None] * world
        dist.all_gather_object(parts, ctx.rp_local_rows(prob.n_p))
        Rp_lib = sp.vstack(parts).tocsr()
        Rp_lib.sort_indices()
        assert abs(Rp_lib - Rp).max() <= 1e-12 * abs(Rp).max()
        Hr = oamg.build_hierarchy(Rp_lib, blocks=pb, replicate_size=repl)
        pc = pa.PCDPreconditioner(prob, "iterative", amg_u=Hu, amg_p=Hp, pcdr=True, amg_r=Hr)
        out["rp_solve_relerr"] = relerr(ctx.rp_solve(v[p0:p1]), pc.solve_Rp(v)[p0:p1])
        assert out["rp_solve_relerr"] <= 1e-9, ("rp_solve", out["rp_solve_relerr"])
    else:
        pc = pa.PCDPreconditioner(prob, "iterative", amg_u=Hu, amg_p=Hp)
    b = rng.standard_normal(prob.n_p)
    out["ap_solve_relerr"] = relerr(ctx.ap_solve(b[p0:p1]), pc.solve_Ap(b)[p0:p1])
    assert out["ap_solve_relerr"] <= 1e-9, "ap_solve"
    bu = rng.standard_normal(prob.n_u)
    out["u_solve_relerr"] = relerr(ctx.u_solve(bu[u0:u1]), pc.solve_A00(bu)[u0:u1])
    assert out["u_solve_relerr"] <= 1e-9, "u_solve"
    xu, xp = rng.standard_normal(prob.n_u), rng.standard_normal(prob.n_p)
    yu, yp = ctx.pc_apply(xu[u0:u1], xp[p0:p1])
    ru, rp = pc.apply_split(xu, xp)
    out["pc_apply_relerr"] = max(relerr(yp, rp[p0:p1]), relerr(yu, ru[u0:u1]))
    assert out["pc_apply_relerr"] <= 1e-8, "pc_apply"
    # the outer solve
    A, rhs = prob.system_matrix(), prob.rhs()
    x_ref, its_ref, hist_ref, _ = pa.fgmres(A, pc, rhs, rtol=1e-6, restart=150)
    su, sp_, its, rn, nap = ctx.solve(prob.b_u[u0:u1], prob.b_p[p0:p1])
    assert abs(its - its_ref) <= 1, (its, its_ref)
    xs = np.concatenate([x_ref[:prob.n_u][u0:u1], x_ref[prob.n_u:][p0:p1]])
    out["solution_relerr"] = relerr(np.concatenate([su, sp_]), xs)
    assert out["solution_relerr"] <= 1e-5
    keys = ["spmv_relerr", "ap_solve_relerr", "u_solve_relerr", "pc_apply_relerr", "solution_relerr"]
    if not pcdr and not repl:
        # value refresh on several ranks: new velocity-block values (same pattern, Kronecker structure kept),
        # coarse operators recomputed on the device with frozen prolongators -- rank local, no communication
        import scipy.sparse as sp
        which = capi.MAT_P00 if prob.P00 is not None else capi.MAT_A00
        A2 = sp.csr_matrix(P00, copy=True)
        d = A2.diagonal()
        A2.setdiag(d * (1.0 + 0.3 * np.repeat(np.sin(np.arange(prob.n_u // dim)), dim) ** 2))
        A2.sort_indices()
        assert np.array_equal(A2.indices, P00.indices)
        ctx.set_values(which, A2[u0:u1, :].tocsr().data)
        ctx.setup()
        Hr = oamg.Hierarchy(smooth_steps=Hu.smooth_steps, eig_ratio=Hu.eig_ratio)
        Ak = A2
        for k, old in enumerate(Hu.levels):
            dk = Ak.diagonal()
            Hr.levels.append(oamg.Level(A=Ak, dinv=np.where(dk != 0, 1.0 / np.where(dk != 0, dk, 1.0), 0.0), rho=old.rho,
                                        P=old.P, R=old.R))
            if k + 1 < len(Hu.levels):
                Ak = (old.R @ Ak @ old.P).tocsr()
        Hr.coarse_inv = np.linalg.inv(Hr.levels[-1].A.toarray())
        out["refresh_u_solve_relerr"] = relerr(ctx.u_solve(bu[u0:u1]), Hr(bu)[u0:u1])
        assert out["refresh_u_solve_relerr"] <= 1e-9, ("refresh", out["refresh_u_solve_relerr"])
        keys.append("refresh_u_solve_relerr")
    out.update({"its": int(its), "oracle_its": int(its_ref), "ndofs": int(prob.n_u + prob.n_p),
                "problem": ("unsteady BFS level 3 PCDR " + variant) if pcdr else
                           ("channel 12x4x6 BRM1" if variant == "BRM1" else "cavity 8^3 BRM2")})
    # worst case over the ranks (each rank checked its own rows)
    t = torch.tensor([out[k] for k in keys], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    for k, v in zip(keys, t.tolist()):
        out[k] = v
    ctx.close()
    return out


def main():
    import faulthandler
    # a hang (a collective one rank never reaches) reports where it is stuck and ends the run
    faulthandler.dump_traceback_later(int(os.environ.get("FNP_TEST_HANG_S", "240")), exit=True)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    variant = sys.argv[1] if len(sys.argv) > 1 else "BRM1"
    out = parity_check(rank, world, local, variant, p2p=os.environ.get("FNP_P2P") or None,
                       repl=int(os.environ.get("FNP_REPL", "0")), pcdr=os.environ.get("FNP_PCDR") == "1")
    dist.barrier()
    if rank == 0:
        print(f"DIST OK world={world} variant={variant} its={out['its']} oracle_its={out['oracle_its']} {out}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
