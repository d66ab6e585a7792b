import sys, time
t0 = time.time()
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from oracle import problems, petsc_algos as pa
from fenapack_b200 import capi
from util import make_context
p0, _ = problems.backward_facing_step(3, variant="BRM2")
x = pa.direct_solver(p0.system_matrix())(p0.rhs())
prob, _ = problems.backward_facing_step(3, variant="BRM2", wind=x[:p0.n_u].reshape(-1, 2), stabilise=True)
print("built", time.time() - t0, flush=True)
rng = np.random.default_rng(0)
res = {}
for tag, extra in (("default", {}), ("reordered", {"fnp_reorder_nodes": 384})):
    ctx = make_context(prob, extra)
    xu, xp = rng.standard_normal(prob.n_u), rng.standard_normal(prob.n_p)
    e = [np.linalg.norm(ctx.spmv(capi.MAT_A00, xu, prob.n_u) - prob.A00 @ xu), np.linalg.norm(ctx.spmv(capi.MAT_A01, xp, prob.n_u) - prob.A01 @ xp),
         np.linalg.norm(ctx.spmv(capi.MAT_A10, xu, prob.n_p) - prob.A10 @ xu)]
    su, sp_, its, rn, nap = ctx.solve(prob.b_u, prob.b_p)
    A, b = prob.system_matrix(), prob.rhs()
    tr = np.linalg.norm(b - A @ np.concatenate([su, sp_])) / np.linalg.norm(b)
    xm, itm, _, _ = ctx.solve_monolithic(np.concatenate([prob.b_u, prob.b_p])[np.argsort(np.concatenate([prob.is_u, prob.is_p]))])
    print(tag, "bs", ctx.block_size(capi.MAT_A00), "spmv err", ["%.1e" % v for v in e], "its", its, itm, "true rel res %.2e" % tr, flush=True)
    ctx.close()
print("total", time.time() - t0)
