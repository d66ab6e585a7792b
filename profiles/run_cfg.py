"""The remaining BASELINE.json configurations on one B200, through the C ABI (inputs from the oracle's
numpy assembler -- the stand-in for DOLFIN assembly, which stays on the host):

    python profiles/run_cfg.py cfg2 [n]          2D lid-driven cavity P2/P1 (n=471: 2.0 M dofs), PCD BRM2
    python profiles/run_cfg.py cfg3 [level] [steps]   unsteady BFS (level 7: 1.63 M dofs), backward Euler dt=0.2,
                                                 Picard steps with per-step value refresh of A00/P00 and Kp
    python profiles/run_cfg.py cfg5newton [n]    3D lid-driven cavity, Newton-coupled velocity block (general
                                                 bs=1 path, no Kronecker structure), PCD BRM2

One JSON line per run: time-to-solve (CUDA events on the library's stream), iterations, PC applies/s, the
per-stage table of bench.py (bytes, GB/s, fraction of the measured HBM peak; operators under ~100 MB are
L2 resident and labelled so) and, for cfg3, the refresh / solve / host-assembly split of the time loop."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
from fenapack_b200 import capi  # noqa: E402
from oracle import petsc_algos as pa  # noqa: E402
from oracle import problems  # noqa: E402
from util import make_context  # noqa: E402

AMG = {"fieldsplit_u_pc_amg_eig_ratio": 4, "fieldsplit_p_PCD_Ap_pc_amg_eig_ratio": 4}


def timed_solves(ctx, prob, reps=5, warm=3):
    import torch
    b = torch.from_numpy(np.concatenate([prob.b_u, prob.b_p])).cuda()
    x = torch.empty_like(b)
    bu, bp, xu, xp = b.data_ptr(), b.data_ptr() + 8 * prob.n_u, x.data_ptr(), x.data_ptr() + 8 * prob.n_u
    for _ in range(warm):
        its, rn, nap = ctx.solve_device(bu, bp, xu, xp)
    ctx.synchronize()
    ctx.tic()
    for _ in range(reps):
        its, rn, nap = ctx.solve_device(bu, bp, xu, xp)
    ms = ctx.toc() / reps
    peak, _ = bench.hbm_peak()
    ctx.set_option("fnp_timers", 2)
    ctx.reset_timers()
    ctx.solve_device(bu, bp, xu, xp)
    stages = bench.stage_table(ctx, peak)
    ctx.set_option("fnp_timers", 0)
    sol = x.cpu().numpy()
    A, rhs = prob.system_matrix(), prob.rhs()
    true_rel = float(np.linalg.norm(rhs - A @ sol) / np.linalg.norm(rhs))
    return {"time_to_solve_ms": ms, "fgmres_iterations": its, "pc_applies_per_s": nap / (ms * 1e-3),
            "true_rel_residual": true_rel, "stages": stages}


def cfg2(n):
    t0 = time.perf_counter()
    prob, _ = problems.lid_driven_cavity(n, dim=2, variant="BRM2")
    t_asm = time.perf_counter() - t0
    ctx = make_context(prob, AMG)
    out = {"config": f"cfg2: 2D lid-driven cavity P2/P1, n={n}, {prob.meta['ndofs']} dofs, PCD BRM2, Chebyshev-Jacobi Mp + "
                     "SA-AMG Ap/velocity, 1 B200", "ndofs": prob.meta["ndofs"], "host_assembly_s": t_asm,
           "kronecker_block_size": ctx.block_size(capi.MAT_A00)}
    out.update(timed_solves(ctx, prob))
    ctx.close()
    return out


def cfg5newton(n):
    t0 = time.perf_counter()
    prob, _ = problems.lid_driven_cavity(n, dim=3, variant="BRM2", newton=True)
    t_asm = time.perf_counter() - t0
    ctx = make_context(prob, AMG)
    out = {"config": f"cfg5 (Newton): 3D lid-driven cavity P2/P1, n={n}, {prob.meta['ndofs']} dofs, full derivative(F, w) "
                     "coupling in the velocity block (general path, no Kronecker structure), PCD BRM2, 1 B200",
           "ndofs": prob.meta["ndofs"], "host_assembly_s": t_asm, "kronecker_block_size": ctx.block_size(capi.MAT_A00)}
    out.update(timed_solves(ctx, prob))
    ctx.close()
    return out


def cfg3(level, steps):
    """Backward Euler on the BFS, Picard linearisation (demo_unsteady-navier-stokes-pcd.py:118-138,188-208):
    every time step re-assembles the convection around the last velocity (host), refreshes the values
    of A00 / P00 / Kp on the device (same pattern) and solves  ((1/dt) M + N(u_k) + nu K) u_{k+1} + B^T p
    = (1/dt) M u_k + lifted boundary data,  B u_{k+1} = 0."""
    dt = 0.2
    t0 = time.perf_counter()
    p0, _ = problems.backward_facing_step(level, variant="BRM1", idt=1.0 / dt)
    t_asm0 = time.perf_counter() - t0
    pst, _ = problems.backward_facing_step(level, variant="BRM1", idt=0.0)
    M_dt = (p0.A00 - pst.A00).tocsr()               # (1/dt) M on the free rows (Dirichlet rows cancel)
    del pst
    ctx = make_context(p0, AMG)
    u = np.zeros(p0.n_u)
    t_host = 0.0
    its_all, refresh_ms, solve_ms = [], [], []
    for step in range(steps):
        t0 = time.perf_counter()
        prob, _ = problems.backward_facing_step(level, variant="BRM1", idt=1.0 / dt, wind=u.reshape(-1, 2))
        b_u, b_p = prob.b_u + M_dt @ u, prob.b_p
        t_host += time.perf_counter() - t0
        t0 = time.perf_counter()
        ctx.set_values(capi.MAT_A00, prob.A00.data)
        ctx.set_values(capi.MAT_KP, prob.Kp.data)
        ctx.setup()
        ctx.synchronize()
        dtr = time.perf_counter() - t0
        t0 = time.perf_counter()
        xu, xp, its, rn, nap = ctx.solve(b_u, b_p)
        dts = time.perf_counter() - t0
        if step > 0:                 # the first refresh also builds the Galerkin plans (once per pattern)
            refresh_ms.append(1e3 * dtr)
        else:
            first_refresh_ms = 1e3 * dtr
        solve_ms.append(1e3 * dts)
        its_all.append(int(its))
        u = xu
    A, rhs = prob.system_matrix(), np.concatenate([b_u, b_p])
    true_rel = float(np.linalg.norm(rhs - A @ np.concatenate([xu, xp])) / np.linalg.norm(rhs))
    out = {"config": f"cfg3: unsteady BFS level {level}, {p0.meta['ndofs']} dofs, backward Euler dt={dt}, {steps} time steps, "
                     "one Picard step each with value refresh of A00 / Kp, PCD BRM1, 1 B200",
           "ndofs": p0.meta["ndofs"], "steps": steps, "fgmres_iterations_per_step": its_all,
           "refresh_ms_median": float(np.median(refresh_ms)) if refresh_ms else None,
           "first_refresh_ms": first_refresh_ms,
           "solve_ms_median": float(np.median(solve_ms)), "host_assembly_s_per_step": t_host / steps,
           "first_assembly_s": t_asm0, "true_rel_residual_last_step": true_rel,
           "kronecker_block_size": ctx.block_size(capi.MAT_A00),
           "note": "refresh = fnp_set_values x2 from pageable host arrays + fnp_setup (device-side Galerkin refresh); solve = "
                   "fnp_solve with host vectors (copies inside); host assembly = the oracle's numpy assembler standing in for DOLFIN"}
    ctx.close()
    return out


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "cfg2":
        res = cfg2(int(sys.argv[2]) if len(sys.argv) > 2 else 471)
    elif what == "cfg3":
        res = cfg3(int(sys.argv[2]) if len(sys.argv) > 2 else 7, int(sys.argv[3]) if len(sys.argv) > 3 else 50)
    else:
        res = cfg5newton(int(sys.argv[2]) if len(sys.argv) > 2 else 32)
    print(json.dumps(res))
