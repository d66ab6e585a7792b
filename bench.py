#!/usr/bin/env python
"""bench.py -- PCD PC applies/s and FGMRES time-to-solve, 3D P2/P1 Oseen, N B200.

One "step" = one right-preconditioned FGMRES(150) solve (rtol 1e-6) of the 3D
P2/P1 Oseen system with the block-triangular PCD preconditioner (inner solvers:
the reference's "iterative" set-up, demo_navier-stokes-pcd.py:153-165, with the
SA-AMG V-cycle in BoomerAMG's slot).  `value` = PC applies per second inside
those solves (device-resident vectors) x Mdof of the system, `ms_per_step` =
time to solution, `e2e` = the same through the C-ABI call with HOST vectors
(copies inside).

Workload (config.workload): lid-driven cavity on the unit cube, n x n x n
bricks x 6 tetrahedra, P2/P1, nu = 0.02, Oseen wind = analytic recirculation,
PCD BRM2 -- BASELINE.json's metric configuration (>= 50 M dofs: n = 128 gives
53.07 M).  Default `--scaling strong`: the SAME n = 128 system on 1, 2, 4, 8
GPUs (rows of all operators partitioned in contiguous z-slabs), as configs[4]
asks.  `--scaling weak --n1 64` keeps 6.7 M dofs per GPU instead
(n = round(n1 * N^(1/3))).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--scaling strong|weak] [--bricks 128] [--impl reference]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "PCD PC applies/s x Mdof (inside FGMRES) & FGMRES time-to-solve, 3D P2/P1 Oseen"
UNIT = "Mdof*applies/s"

OPTIONS = {
    "ksp_type": "fgmres",
    "ksp_gmres_restart": 150,
    "ksp_rtol": 1e-6,
    "ksp_max_it": 400,
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
    "fieldsplit_p_PCD_Mp_ksp_chebyshev_eigenvalues": "0.5, 2.5",
    "fieldsplit_p_PCD_Mp_pc_type": "jacobi",
    # SA-AMG settings of the bench (library defaults: eig_ratio 10, coarse_size 400): Chebyshev smoother
    # on [rho/4, rho], velocity hierarchy stopped at <= 2000 rows (dense inverse).  Round-2 sweep on the
    # 64^3 / 128^3 cavities: 21 -> 17 and 19 -> 17 outer iterations (profiles/r02_tuning_sweep.md)
    "fieldsplit_u_pc_amg_eig_ratio": 4,
    "fieldsplit_p_PCD_Ap_pc_amg_eig_ratio": 4,
    "fieldsplit_u_pc_amg_coarse_size": 2000,
}


def oracle_amg_kwargs(prefix, opts=None):
    """The SA-AMG settings of OPTIONS as keyword arguments of the oracle's set-up / V-cycle."""
    opts = OPTIONS if opts is None else opts
    kw = {}
    if prefix + "pc_amg_eig_ratio" in opts:
        kw["eig_ratio"] = float(opts[prefix + "pc_amg_eig_ratio"])
    if prefix + "pc_amg_coarse_size" in opts:
        kw["coarse_size"] = int(opts[prefix + "pc_amg_coarse_size"])
    return kw


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cavity3d", choices=["cavity3d", "channel3d"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: the same n^3 system on every N (default); weak: n = round(n1 * N^(1/3))")
    ap.add_argument("--bricks", "--n", dest="n", type=int, default=128,
                    help="bricks per side of the whole system (strong scaling); use --bricks under torchrun, whose own "
                         "parser claims the abbreviation --n")
    ap.add_argument("--n1", type=int, default=64, help="bricks per side on one GPU (weak scaling base)")
    ap.add_argument("--no-refresh", action="store_true", help="skip the value-refresh timing")
    ap.add_argument("--no-dist-parity", action="store_true", help="skip the small oracle-checked problem at N > 1")
    ap.add_argument("--variant", default=None, choices=["BRM1", "BRM2"])
    ap.add_argument("--nu", type=float, default=0.02)
    ap.add_argument("--cpu-sample-its", type=int, default=12,
                    help="FGMRES iterations of the same solve timed on the host cores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-only", action="store_true", help="skip e2e/cpu legs (for ncu runs)")
    ap.add_argument("--no-clocks", action="store_true", help="do not poll nvidia-smi during the timed region")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="extra library option (PETSc-style name), e.g. fieldsplit_u_pc_amg_smooth_steps=3")
    return ap.parse_args()


def mesh_size(args, world):
    n = args.n if args.scaling == "strong" else int(round(args.n1 * world ** (1.0 / 3.0)))
    if args.workload == "cavity3d":
        return (n, n, n), "cavity", args.variant or "BRM2"
    m = max(2, int(round(n / 4.0 ** (1.0 / 3.0))))      # 4m x m x m bricks of the 4x1x1 box: ~n^3 bricks
    return (4 * m, m, m), "channel", args.variant or "BRM1"


def workload_name(args, dims, kind, variant, ndofs):
    return (f"{'lid-driven cavity' if kind == 'cavity' else 'channel'} 3D P2/P1 Oseen, {dims[0]}x{dims[1]}x{dims[2]} bricks x6 tets, "
            f"{ndofs} dofs, nu={args.nu}, PCD {variant}, FGMRES(150) rtol 1e-6, "
            "u: richardson x1 + SA-AMG V(2,2) Chebyshev-Jacobi, Ap: richardson x2 + SA-AMG, Mp: chebyshev x5 + jacobi"
            + "".join(f", {k}={v}" for k, v in sorted(OPTIONS.items()) if "pc_amg_" in k))


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock, power and throttle reasons of one GPU DURING the timed region.
    In-process NVML (nvidia_ml_py) from a background thread: spawning nvidia-smi from a
    process that holds a CUDA context, and its NVML polling, measurably perturb
    launch-bound multi-rank runs."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, gpu_index=0, period=0.1):
        self.idx, self.period = gpu_index, period
        self.samples = []
        self._stop = None
        self._thr = None
        self.err = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.idx)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self._stop.is_set():
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((sm, pw, rs))
                self._stop.wait(self.period)
        except Exception as e:      # pragma: no cover
            self.err = str(e)

    def __enter__(self):
        import threading
        self._stop = threading.Event()
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=5)

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": len(self.samples), "how": "NVML in-process"}
        if self.samples:
            out["sm_mhz"] = float(np.median([s[0] for s in self.samples]))
            out["sm_max_mhz"] = float(getattr(self, "sm_max", 0))
            out["power_w_max"] = max(s[1] for s in self.samples)
            for name, bit in self.REASONS:
                if any(s[2] & bit for s in self.samples):
                    out["reasons"].append(name)
        if self.err:
            out["error"] = self.err
        return out


class NoClocks:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass

    def summary(self):
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def spmv_bytes(nrows, ncols, nnz):
    # matrix once (8 B value + 4 B column), row pointers, y written once, x read once
    return 12.0 * nnz + 4.0 * (nrows + 1) + 8.0 * nrows + 8.0 * ncols


# ---------------------------------------------------------------------------
# CPU arm (oracle port, C/OpenMP): the reference's algorithm chain on the host cores
# ---------------------------------------------------------------------------
def cpu_port_sample(prob, hier_u, hier_p, variant, its, rtol=1e-6):
    """Time the first `its` FGMRES iterations of the same solve with the
    C/OpenMP restatement (oracle/pcd_ref.c).  Returns (applies/s, seconds, its, threads)."""
    from oracle import cref
    mats = {k: prob.scipy(k) for k in ("A00", "A01", "A10", "Ap", "Mp", "Kp")}
    pc = cref.CPCD(mats, variant, prob.bc_idx, prob.bc_val, hier_u, hier_p, prob.cheb_bounds)
    b = np.concatenate([prob.b_u, prob.b_p])
    pc.fgmres(b, rtol=rtol, restart=150, max_it=1)          # warm-up (page in, thread pool)
    t0 = time.perf_counter()
    x, n_it, hist, nap = pc.fgmres(b, rtol=rtol, restart=150, max_it=its)
    dt = time.perf_counter() - t0
    return nap / dt, dt, n_it, cref.num_threads(), hist


def oracle_hierarchies_from_library(ctx, opts=None):
    """The hierarchies the GPU library built, as oracle objects (same inner operators
    on both sides)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fenapack_b200 import capi
    from util import oracle_hierarchy_from_device
    eu = oracle_amg_kwargs("fieldsplit_u_", opts).get("eig_ratio", 10.0)
    ep = oracle_amg_kwargs("fieldsplit_p_PCD_Ap_", opts).get("eig_ratio", 10.0)
    return (oracle_hierarchy_from_device(ctx, capi.MAT_A00, eig_ratio=eu),
            oracle_hierarchy_from_device(ctx, capi.MAT_AP, eig_ratio=ep))


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
STAGE_ORDER = ("FENaPack: PCDKSP solve", "FENaPack: PCD fieldsplit apply", "FENaPack: PCDPC_BRM1 apply",
               "FENaPack: PCDPC_BRM2 apply", "FENaPack: PCD_Ap solve", "FENaPack: PCD_Mp solve",
               "FENaPack: fieldsplit_u solve", "FENaPack: A01 mult", "FENaPack: system MatMult",
               "FENaPack: GMRES orthogonalization")


def stage_table(ctx, peak):
    """Every timer of the instrumented solve.  Per-kernel timers (spmv <operator>, multidot,
    maxpy+norm, gemv coarse) carry their algorithmic bytes: GB/s and the fraction
    of the measured HBM peak follow.  Operators of at most ~100 MB are L2 resident between
    their launches; their fraction is labelled as such, not an HBM figure."""
    out = {}
    names = ctx.timer_names()
    for nm in list(STAGE_ORDER) + sorted(n for n in names if n not in STAGE_ORDER):
        t_ms, calls = ctx.timer(nm)
        if not calls:
            continue
        e = {"ms": t_ms, "calls": calls}
        by = ctx.timer_bytes(nm)
        if by > 0 and t_ms > 0:
            e["bytes_per_launch"] = by / calls
            e["gbps"] = by / (t_ms * 1e-3) / 1e9
            e["frac_of_hbm_peak"] = e["gbps"] / peak
            if by / calls < 100e6:
                e["l2_resident"] = True
        out[nm] = e
    return out


def upload(ctx, capi, prob):
    ctx.set_layout(prob.n_u, prob.n_p, prob.u_begin, prob.n_u_global, prob.p_begin, prob.n_p_global)
    for name, which in (("A00", capi.MAT_A00), ("A01", capi.MAT_A01), ("A10", capi.MAT_A10),
                        ("Ap", capi.MAT_AP), ("Mp", capi.MAT_MP), ("Kp", capi.MAT_KP)):
        rp, ci, va = getattr(prob, name)
        ctx.set_pattern(which, rp, ci)
        ctx.set_values(which, va)
    ctx.set_bc(prob.bc_idx, prob.bc_val)


def b200_arm(args):
    import torch
    import torch.distributed as dist
    from fenapack_b200 import capi
    import bench_inputs as bi

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- N > 1: a small oracle-checked problem first, so that the multi-rank path carries its
    # own parity evidence in the bench line (the GPU test box has one GPU) --------------------
    dist_parity = None
    if world > 1 and not args.no_dist_parity and not args.profile_only:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from dist_worker import parity_check
        try:
            dist_parity = parity_check(rank, world, local, "BRM2")
            dist_parity["ok"] = True
        except AssertionError as e:
            dist_parity = {"ok": False, "failed": str(e)[:200]}

    dims, kind, variant = mesh_size(args, world)
    t0 = time.perf_counter()
    prob = bi.generate(*dims, kind=kind, nu=args.nu, variant=variant, rank=rank, nranks=world, device=f"cuda:{local}")
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    t_gen = time.perf_counter() - t0

    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ctx = capi.Context(local, nccl_id=bytes(idt.cpu().numpy().tobytes()), rank=rank, nranks=world)
    else:
        ctx = capi.Context(local)
    opts = dict(OPTIONS)
    opts["fieldsplit_p_pc_python_type"] = "fenapack.PCDPC_" + variant
    for kv in args.opt:
        k, _, v = kv.partition("=")
        opts[k] = v
    ctx.set_options(opts)
    if "FNP_OVERLAP" in os.environ:
        ctx.set_option("fnp_halo_overlap", os.environ["FNP_OVERLAP"])
    if "FNP_P2P" in os.environ:
        ctx.set_option("fnp_halo_p2p", os.environ["FNP_P2P"])
    if "FNP_GRAPH" in os.environ:
        ctx.set_option("fnp_cuda_graph", os.environ["FNP_GRAPH"])
    t0 = time.perf_counter()
    upload(ctx, capi, prob)
    t_upload = time.perf_counter() - t0
    t0 = time.perf_counter()
    ctx.setup()
    t_setup = time.perf_counter() - t0

    n_loc = prob.n_u + prob.n_p
    # device-resident right-hand side / solution (torch is only the allocator here)
    b_dev = torch.from_numpy(np.concatenate([prob.b_u, prob.b_p])).cuda()
    x_dev = torch.empty_like(b_dev)
    bu_p, bp_p = b_dev.data_ptr(), b_dev.data_ptr() + 8 * prob.n_u
    xu_p, xp_p = x_dev.data_ptr(), x_dev.data_ptr() + 8 * prob.n_u
    torch.cuda.synchronize()

    def sync_all():
        barrier()
        ctx.synchronize()

    # ---- warm-up -------------------------------------------------------------
    nwarm = args.warmup if args.profile_only else max(args.warmup, 3)
    for _ in range(nwarm):
        its, rn, nap = ctx.solve_device(bu_p, bp_p, xu_p, xp_p)
    # ---- timed region: K solves, CUDA events on the library's stream ----------
    # one sampler per job (rank 0's GPU): N copies of nvidia-smi polling NVML perturb the
    # launch-bound parts of a multi-rank run
    with (ClockSampler(local) if (not args.no_clocks and rank == 0) else NoClocks()) as clk:
        sync_all()
        l0 = ctx.kernel_launches()
        ctx.tic()
        total_applies = 0
        for _ in range(args.steps):
            its, rn, nap = ctx.solve_device(bu_p, bp_p, xu_p, xp_p)
            total_applies += nap
        ms = ctx.toc()
        sync_all()
        launches = ctx.kernel_launches() - l0
    clocks = clk.summary()
    ms = max_over_ranks(ms)
    applies_per_s = total_applies / (ms * 1e-3)
    mdof = prob.ndofs_global / 1e6
    # whole-job throughput: (PC applies per second) x (Mdof of the system they act on), so that
    # runs on systems of different size (weak scaling) compare; pc_applies_per_s is BASELINE's
    # bare metric for this system
    value = applies_per_s * mdof
    hist = ctx.residual_history()
    final_rel = float(hist[-1] / hist[0])

    # true residual ||b - A x|| / ||b|| through distributed SpMVs of the stored operators
    yu = torch.empty(prob.n_u, dtype=torch.float64, device="cuda")
    yu2 = torch.empty_like(yu)
    yp = torch.empty(prob.n_p, dtype=torch.float64, device="cuda")
    ctx.spmv_device(capi.MAT_A00, xu_p, yu.data_ptr())
    ctx.spmv_device(capi.MAT_A01, xp_p, yu2.data_ptr())
    ctx.spmv_device(capi.MAT_A10, xu_p, yp.data_ptr())
    ctx.synchronize()
    ru = b_dev[:prob.n_u] - yu - yu2
    rp_ = b_dev[prob.n_u:] - yp
    rr = sum_over_ranks(float((ru * ru).sum() + (rp_ * rp_).sum()))
    bb = sum_over_ranks(float((b_dev * b_dev).sum()))
    true_rel = float(np.sqrt(rr / bb))
    del yu, yu2, yp, ru, rp_

    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "pc_applies_per_s": applies_per_s,
        "steps": args.steps, "warmup": nwarm, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, dims, kind, variant, prob.ndofs_global),
                   "ndofs": prob.ndofs_global, "n_u": prob.n_u_global, "n_p": prob.n_p_global,
                   "partition": f"row slabs x{world}", "l2": "inputs larger than L2 (A00 alone >> 126 MB), no flush"},
        "time_to_solve_ms": ms / args.steps, "fgmres_iterations": its, "pc_applies_per_solve": nap,
        "final_rel_residual": final_rel, "true_rel_residual": true_rel, "clocks": clocks, "gpu_launches": int(launches),
        "setup": {"generate_s": t_gen, "upload_s": t_upload, "fnp_setup_s": t_setup},
    }
    if dist_parity is not None:
        result["dist_parity"] = dist_parity
    if args.opt:
        result["config"]["extra_options"] = list(args.opt)

    if not args.profile_only:
        # ---- e2e: the C-ABI call with HOST (pinned) vectors, copies inside the timed region ----
        bu_h = torch.from_numpy(prob.b_u.copy()).pin_memory().numpy()
        bp_h = torch.from_numpy(prob.b_p.copy()).pin_memory().numpy()
        xu_h = torch.empty(prob.n_u, dtype=torch.float64).pin_memory().numpy()
        xp_h = torch.empty(prob.n_p, dtype=torch.float64).pin_memory().numpy()
        ctx.solve(bu_h, bp_h, out=(xu_h, xp_h))
        sync_all()
        t0 = time.perf_counter()
        applies = 0
        for _ in range(args.steps):
            xu, xp, its_h, rn_h, nap_h = ctx.solve(bu_h, bp_h, out=(xu_h, xp_h))
            applies += nap_h
        sync_all()
        dt = max_over_ranks(time.perf_counter() - t0)
        hess = sum(8 * (j + 2) for j in range(its_h)) + 16
        result["e2e"] = {"value": applies / dt * mdof, "unit": UNIT, "pc_applies_per_s": applies / dt,
                         "ms_per_solve": 1e3 * dt / args.steps,
                         "h2d_bytes_per_step": 8 * n_loc, "d2h_bytes_per_step": 8 * n_loc + hess,
                         "api": "fnp_solve(host pointers; b and x in pinned host memory)"}
        del bu_h, bp_h, xu_h, xp_h
        # standalone PC applies on device-resident vectors
        z_dev = torch.empty_like(b_dev)
        zu_p, zp_p = z_dev.data_ptr(), z_dev.data_ptr() + 8 * prob.n_u
        for _ in range(3):
            ctx.pc_apply_device(bu_p, bp_p, zu_p, zp_p)
        sync_all()
        ctx.tic()
        reps = 20
        for _ in range(reps):
            ctx.pc_apply_device(bu_p, bp_p, zu_p, zp_p)
        pms = max_over_ranks(ctx.toc())
        result["pc_apply_only"] = {"applies_per_s": reps / (pms * 1e-3), "ms_per_apply": pms / reps,
                                   "mdof_applies_per_s": reps / (pms * 1e-3) * mdof}
        del z_dev

    # ---- roofline of the dominant kernel and of every stage, measured in situ (instrumented solve) ----
    peak, peak_src = hbm_peak()
    ctx.set_option("fnp_timers", 2)
    ctx.reset_timers()
    ctx.solve_device(bu_p, bp_p, xu_p, xp_p)
    stage = stage_table(ctx, peak)
    ctx.set_option("fnp_timers", 0)
    bs = ctx.block_size(capi.MAT_A00)
    roof = {"bound": "hbm", "kernel": "spmv_sell_kernel<Epi> on A00 (P2 velocity block: AMG level-0 smoother/residual + outer MatMult)",
            "peak": peak, "peak_source": peak_src, "unit": "GB/s",
            "kronecker_block_size": bs,
            "bytes_note": "12 B per STORED entry + row pointers + y written once + x read once (owned + ghost entries, this "
                          "rank's rows); with kronecker_block_size 3 the stored operator is the scalar S of A00 = S (x) I_3"}
    if "spmv A00" in stage:
        e = stage["spmv A00"]
        roof["bytes_per_launch"] = e["bytes_per_launch"]
        roof["avg_launch_ms"] = e["ms"] / e["calls"]
        roof["achieved"] = e["gbps"]
        roof["frac"] = e["frac_of_hbm_peak"]
        tot = stage.get("FENaPack: PCDKSP solve", {}).get("ms")
        if tot:
            roof["share_of_step"] = e["ms"] / tot
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    roof["traffic"] = None
    if os.path.exists(tpath) and world == 1:
        # DRAM bytes per launch of this kernel from the ncu --set full capture at the same size (one rank)
        try:
            tj = json.load(open(tpath))
            roof["traffic"] = tj.get("spmv_A00_dram_bytes_per_launch", {}).get(str(prob.ndofs_global))
            roof["traffic_source"] = tj.get("source")
            roof["traffic_note"] = "captured on the store-epilogue launch (no extra epilogue vectors: 12 B per stored entry + row " \
                                   "pointers + x + y); bytes_per_launch is the mean over the five launch kinds of an iteration"
        except Exception:
            pass
    result["roofline"] = roof
    result["stages"] = stage

    # ---- value refresh: the per-Newton-step / per-time-step path (new A00 and Kp values in the
    # caller's layout, pinned host memory -> fnp_set_values x2 + fnp_setup), then one solve ------
    if not args.no_refresh and not args.profile_only:
        try:
            va = torch.from_numpy(prob.A00[2]).pin_memory().numpy()
            vk = torch.from_numpy(prob.Kp[2]).pin_memory().numpy()

            def refresh():
                ctx.set_values(capi.MAT_A00, va)
                ctx.set_values(capi.MAT_KP, vk)
                ctx.setup()
                ctx.synchronize()
            sync_all()
            t0 = time.perf_counter()
            refresh()                      # the first refresh also builds the Galerkin plans (once per pattern)
            t_first = max_over_ranks(time.perf_counter() - t0)
            reps = 3
            sync_all()
            t0 = time.perf_counter()
            for _ in range(reps):
                refresh()
            sync_all()
            t_ref = max_over_ranks(time.perf_counter() - t0) / reps
            its_r, _, _ = ctx.solve_device(bu_p, bp_p, xu_p, xp_p)
            result["refresh"] = {"ms": 1e3 * t_ref, "first_ms": 1e3 * t_first, "h2d_bytes": int(va.nbytes + vk.nbytes),
                                 "its_after": int(its_r),
                                 "what": "fnp_set_values(A00) + fnp_set_values(Kp) from pinned host arrays + fnp_setup; "
                                         + "device-side: value scatter, Jacobi diagonals, Galerkin coarse operators with frozen "
                                           "prolongators (rank local); the coarsest level is gathered and inverted on the host"}
            del va, vk
        except Exception as e:
            result["refresh"] = {"ms": None, "failed": str(e)[:300]}

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on the host cores, bounded sample ----
    if world == 1 and not args.no_cpu_baseline and not args.profile_only:
        try:
            Hu, Hp = oracle_hierarchies_from_library(ctx, opts)
            v, dt, n_it, thr, chist = cpu_port_sample(prob, Hu, Hp, variant, args.cpu_sample_its)
            k = min(len(chist), len(hist)) - 1
            result["cpu_baseline"] = {
                "value": v * mdof, "unit": UNIT, "pc_applies_per_s": v, "cores": thr, "kind": "port",
                "sample": f"first {n_it} FGMRES iterations of the same solve ({dt:.1f} s), C/OpenMP restatement "
                          "oracle/pcd_ref.c on the AMG hierarchy the library built",
                "host_cpus": os.cpu_count(),
                "residual_after_sample_rel_diff_vs_gpu": float(abs(chist[k] - hist[k]) / hist[k]),
            }
        except Exception as e:  # keep the bench line even if the host leg fails
            result["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                      "sample": f"failed: {e}"}
    if rank == 0:
        print(json.dumps(result))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------
# reference arm: the reference's algorithm chain on the host cores (oracle port)
# ---------------------------------------------------------------------------
def reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    import torch
    import bench_inputs as bi
    from oracle import amg as oamg
    from oracle import cref
    # the same workload as the B200 arm at this N (strong scaling: the same system for every N),
    # all host cores of the box; every step is a bounded sample of its solve
    dims, kind, variant = mesh_size(args, world)
    dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    prob = bi.generate(*dims, kind=kind, nu=args.nu, variant=variant, device=dev)
    if dev != "cpu":
        torch.cuda.empty_cache()
    t0 = time.perf_counter()
    # same hierarchy arrangement as the library (velocity block = S (x) I_3: S is coarsened);
    # the CPU arm itself works on the general CSR format, as PETSc's AIJ would
    Hu = oamg.build_hierarchy_kron(prob.scipy("A00"), bs=3, S=prob.scipy("S00"), **oracle_amg_kwargs("fieldsplit_u_"))
    Hp = oamg.build_hierarchy(prob.scipy("Ap"), **oracle_amg_kwargs("fieldsplit_p_PCD_Ap_"))
    t_setup = time.perf_counter() - t0
    mats = {k: prob.scipy(k) for k in ("A00", "A01", "A10", "Ap", "Mp", "Kp")}
    pc = cref.CPCD(mats, variant, prob.bc_idx, prob.bc_val, Hu, Hp, prob.cheb_bounds)
    b = np.concatenate([prob.b_u, prob.b_p])
    mdof = prob.ndofs_global / 1e6
    # iterations per step: about 1.5 s of CPU work per step (measured ~25 ms per Mdof and iteration on 16 cores)
    its = int(max(2, min(args.cpu_sample_its, round(60.0 / mdof))))
    for _ in range(max(1, min(args.warmup, 1))):
        pc.fgmres(b, rtol=1e-6, restart=150, max_it=1)
    t0 = time.perf_counter()
    applies = 0
    for _ in range(args.steps):
        _, n_it, hist, nap = pc.fgmres(b, rtol=1e-6, restart=150, max_it=its)
        applies += nap
    dt = time.perf_counter() - t0
    v = applies / dt * mdof
    thr = cref.num_threads()
    sample = (f"each step = the first {its} FGMRES iterations (the cheapest ones: short Gram-Schmidt) of the solve of the "
              f"same {prob.ndofs_global}-dof system; C/OpenMP restatement (oracle/pcd_ref.c) of the PETSc algorithm "
              f"chain, {thr} threads = all host cores, whatever N")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "pc_applies_per_s": applies / dt,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, dims, kind, variant, prob.ndofs_global),
                   "ndofs": prob.ndofs_global},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": thr, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "setup": {"oracle_amg_setup_s": t_setup},
        "note": "the reference itself (petsc4py/DOLFIN/hypre) is not installable in this image; see DESIGN.md",
    }))


if __name__ == "__main__":
    _w = int(os.environ.get("WORLD_SIZE", 1))
    a = parse_args()
    if a.impl == "reference":
        # rank 0 alone works: it gets every host core
        if int(os.environ.get("RANK", 0)) == 0 and (_w > 1 or os.environ.get("OMP_NUM_THREADS", "") == "1"):
            os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        reference_arm(a)
    else:
        if _w > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
            # host-side set-up (AMG hierarchy) is OpenMP code: share the cores between the ranks
            # (torchrun pins OMP_NUM_THREADS=1 by default, which would serialise it)
            os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // _w))
        b200_arm(a)
