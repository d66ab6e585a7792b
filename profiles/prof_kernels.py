"""Profiling driver (run under ncu on the GPU box): sets up the bench workload and
launches each SpMV-class kernel of the hot path a few times in isolation, so that
`ncu -k regex:<name> --set full` captures land on a known operator.
    python profiles/prof_kernels.py [n1] [reps]
Prints CUDA-event timings per operator as well (never under the profiler for numbers)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bench_inputs as bi  # noqa: E402
from fenapack_b200 import capi  # noqa: E402

n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
prob = bi.generate(n1, n1, n1, kind="cavity", variant="BRM2", device="cuda:0")
torch.cuda.empty_cache()
ctx = capi.Context(0)
opts = dict(bench.OPTIONS)
opts["fieldsplit_p_pc_python_type"] = "fenapack.PCDPC_BRM2"
ctx.set_options(opts)
if len(sys.argv) > 3:
    ctx.set_option("fnp_spmv_kernel", sys.argv[3])
ctx.set_layout(prob.n_u, prob.n_p)
ops = (("A00", capi.MAT_A00), ("A01", capi.MAT_A01), ("A10", capi.MAT_A10), ("Ap", capi.MAT_AP),
       ("Mp", capi.MAT_MP), ("Kp", capi.MAT_KP))
for name, which in ops:
    rp, ci, va = getattr(prob, name)
    ctx.set_pattern(which, rp, ci)
    ctx.set_values(which, va)
ctx.set_bc(prob.bc_idx, prob.bc_val)
ctx.setup()
xu = torch.randn(prob.n_u, dtype=torch.float64, device="cuda")
xp = torch.randn(prob.n_p, dtype=torch.float64, device="cuda")
yu = torch.empty_like(xu)
yp = torch.empty_like(xp)
torch.cuda.synchronize()
peak, _ = bench.hbm_peak()
for name, which in ops:
    rp, ci, va = getattr(prob, name)
    nrows, nnz = rp.size - 1, int(rp[-1])
    x = xu if name in ("A00", "A10") else xp
    y = yu if name in ("A00", "A01") else yp
    ncols = x.numel()
    for _ in range(2):
        ctx.spmv_device(which, x.data_ptr(), y.data_ptr())
    ctx.synchronize()
    ctx.tic()
    for _ in range(reps):
        ctx.spmv_device(which, x.data_ptr(), y.data_ptr())
    ms = ctx.toc() / reps
    byt = bench.spmv_bytes(nrows, ncols, nnz)
    print(f"spmv {name}: rows {nrows} nnz {nnz} mean_row {nnz / nrows:.2f} {ms:.4f} ms  "
          f"{byt / ms / 1e6:.0f} GB/s  frac {byt / ms / 1e6 / peak:.3f}")
# one block-triangular PC apply and one V-cycle, for the launch list
zu = torch.empty_like(xu)
zp = torch.empty_like(xp)
ctx.pc_apply_device(xu.data_ptr(), xp.data_ptr(), zu.data_ptr(), zp.data_ptr())
ctx.synchronize()
ctx.tic()
for _ in range(reps):
    ctx.pc_apply_device(xu.data_ptr(), xp.data_ptr(), zu.data_ptr(), zp.data_ptr())
print(f"pc_apply: {ctx.toc() / reps:.3f} ms")
for which, nm in ((capi.MAT_A00, "A00"), (capi.MAT_AP, "Ap")):
    nl = capi.C.c_int32()
    ctx._lib.fnp_amg_num_levels(ctx._h, which, capi.C.byref(nl))
    info = []
    for l in range(nl.value):
        nr, nc, nnz, rho = capi.C.c_int64(), capi.C.c_int64(), capi.C.c_int64(), capi.C.c_double()
        ctx._lib.fnp_amg_level_info(ctx._h, which, l, 0, capi.C.byref(nr), capi.C.byref(nc), capi.C.byref(nnz), capi.C.byref(rho))
        info.append((nr.value, nnz.value, round(rho.value, 3)))
    print("AMG", nm, "levels (rows, nnz, rho):", info)
ctx.close()
