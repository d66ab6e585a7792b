"""Experiment driver: 16-byte gathers and the SELL sorting window of the Kronecker SpMV kernel.
    python profiles/exp_sell_gather.py [n1]
Per (sigma, gather) configuration: isolated A00 SpMV time (CUDA events), one PC apply, one solve."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bench_inputs as bi  # noqa: E402
from fenapack_b200 import capi  # noqa: E402

n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prob = bi.OseenBoxProblem(n1, n1, n1, kind="cavity", variant="BRM2", device="cuda:0")
torch.cuda.empty_cache()
ops = (("A00", capi.MAT_A00), ("A01", capi.MAT_A01), ("A10", capi.MAT_A10), ("Ap", capi.MAT_AP),
       ("Mp", capi.MAT_MP), ("Kp", capi.MAT_KP))
xu = torch.randn(prob.n_u, dtype=torch.float64, device="cuda")
xp = torch.randn(prob.n_p, dtype=torch.float64, device="cuda")
yu, zu, zp = torch.empty_like(xu), torch.empty_like(xu), torch.empty_like(xp)
ref = None
bu = torch.from_numpy(prob.b_u).cuda()
bp = torch.from_numpy(prob.b_p).cuda()
su, sp_ = torch.empty_like(bu), torch.empty_like(bp)


def make(sigma, gather):
    ctx = capi.Context(0)
    opts = dict(bench.OPTIONS)
    opts["fieldsplit_p_pc_python_type"] = "fenapack.PCDPC_BRM2"
    opts["fnp_sell_sigma"] = sigma
    opts["fnp_sell_gather"] = gather
    ctx.set_options(opts)
    ctx.set_layout(prob.n_u, prob.n_p)
    for name, which in ops:
        rp, ci, va = getattr(prob, name)
        ctx.set_pattern(which, rp, ci)
        ctx.set_values(which, va)
    ctx.set_bc(prob.bc_idx, prob.bc_val)
    ctx.setup()
    return ctx


def time_spmv(ctx, which, x, y, reps=20):
    for _ in range(3):
        ctx.spmv_device(which, x.data_ptr(), y.data_ptr())
    ctx.synchronize()
    ctx.tic()
    for _ in range(reps):
        ctx.spmv_device(which, x.data_ptr(), y.data_ptr())
    return ctx.toc() / reps


def time_solve(ctx, tag):
    for _ in range(3):
        ctx.pc_apply_device(xu.data_ptr(), xp.data_ptr(), zu.data_ptr(), zp.data_ptr())
    ctx.synchronize()
    ctx.tic()
    for _ in range(10):
        ctx.pc_apply_device(xu.data_ptr(), xp.data_ptr(), zu.data_ptr(), zp.data_ptr())
    pc = ctx.toc() / 10
    for _ in range(2):
        its, rn, nap = ctx.solve_device(bu.data_ptr(), bp.data_ptr(), su.data_ptr(), sp_.data_ptr())
    ctx.synchronize()
    ctx.tic()
    for _ in range(3):
        its, rn, nap = ctx.solve_device(bu.data_ptr(), bp.data_ptr(), su.data_ptr(), sp_.data_ptr())
    print(f"{tag}: pc_apply {pc:.3f} ms; solve {ctx.toc() / 3:.2f} ms, {its} its", flush=True)


yp = torch.empty_like(xp)
ctx = make(1024, 0)
out = []
for g in (0, 4, 7, 15, 45, 47):
    ctx.set_option("fnp_sell_gather", g)
    ms = time_spmv(ctx, capi.MAT_A00, xu, yu)
    if ref is None:
        ref = yu.clone()
    out.append(f"g{g} {ms:.4f} ({float((yu - ref).abs().max()):.0e})")
for g in (0, 4, 79):      # 79 = 15 + 64: default cache policy (L2 resident) for operators <= 64 MB
    ctx.set_option("fnp_sell_gather", g)
    out.append(f"| A01 g{g} {time_spmv(ctx, capi.MAT_A01, xp, yu):.4f} Ap g{g} {time_spmv(ctx, capi.MAT_AP, xp, yp):.4f}")
for g in (15, 31):            # bit 16: L2 prefetch in the CSR sub-warp kernel (A10, 181 nnz per row)
    ctx.set_option("fnp_sell_gather", g)
    out.append(f"| A10 g{g} {time_spmv(ctx, capi.MAT_A10, xu, yp):.4f}")
print(" ".join(out), flush=True)
for g in (7, 15, 31, 47, 79):
    ctx.set_option("fnp_sell_gather", g)
    time_solve(ctx, f"g{g}")
    sol = torch.cat([su, sp_]).clone()
    if g == 7:
        sol_ref = sol
    else:
        print(f"   solution max diff vs g7: {float((sol - sol_ref).abs().max()):.1e}", flush=True)
ctx.close()
