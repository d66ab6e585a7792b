"""world_size-2 gloo worker (CPU): the N>1 host logic -- ownership ranges, ghost lists,
request exchange, halo exchange, distributed Chebyshev-Jacobi and a global dot --
against the serial oracle.  Launched by tests/test_dist_cpu.py."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dist as odist  # noqa: E402
from oracle import petsc_algos as pa  # noqa: E402
from oracle import problems  # noqa: E402
import bench_inputs as bi  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    prob, _ = problems.channel(8, 3, 4, variant="BRM1")
    ub, pb = odist.split_rows(prob.n_u, world, 3), odist.split_rows(prob.n_p, world)
    rng = np.random.default_rng(0)
    spaces = {"u": (ub, prob.n_u), "p": (pb, prob.n_p)}
    ops = {"A00": (prob.A00, "u", "u"), "A01": (prob.A01, "u", "p"), "A10": (prob.A10, "p", "u"),
           "Ap": (prob.Ap, "p", "p"), "Mp": (prob.Mp, "p", "p"), "Kp": (prob.Kp, "p", "p")}
    plans = {}
    for name, (A, rs, cs) in ops.items():
        rb, cb = spaces[rs][0], spaces[cs][0]
        plan = odist.HaloPlan(A[rb[rank]:rb[rank + 1], :], cb, rank)
        plans[name] = plan
        x = rng.standard_normal(A.shape[1])
        y = plan.spmv(x[cb[rank]:cb[rank + 1]])
        ref = (A @ x)[rb[rank]:rb[rank + 1]]
        assert np.linalg.norm(y - ref) <= 1e-13 * max(np.linalg.norm(ref), 1e-300), name
        # every ghost is owned by another rank and every request is in the owner's range
        assert np.all((plan.ghosts < cb[rank]) | (plan.ghosts >= cb[rank + 1]))
        assert sum(plan.send_count) == sum(len(s) for s in plan.send_idx)
    # distributed Chebyshev-Jacobi (SpMV + local vector work) == serial
    p0, p1 = pb[rank], pb[rank + 1]
    b = rng.standard_normal(prob.n_p)
    dinv = 1.0 / prob.Mp.diagonal()

    class DistMp:
        def __matmul__(self, v):
            return plans["Mp"].spmv(v)
    x_loc = pa.chebyshev_jacobi(DistMp(), dinv[p0:p1], b[p0:p1], 0.5, 2.5, 5)
    x_ser = pa.chebyshev_jacobi(prob.Mp, dinv, b, 0.5, 2.5, 5)
    assert np.linalg.norm(x_loc - x_ser[p0:p1]) <= 1e-13 * np.linalg.norm(x_ser)
    assert abs(odist.global_dot(b[p0:p1], x_ser[p0:p1]) - float(b @ x_ser)) <= 1e-12 * abs(float(b @ x_ser))
    # the bench generator's slab partition: rank slabs are the row slices of the global problem
    full = bi.OseenBoxProblem(3, 2, 4, kind="channel", variant="BRM1", device="cpu")
    mine = bi.OseenBoxProblem(3, 2, 4, kind="channel", variant="BRM1", device="cpu", rank=rank, nranks=world)
    for name in ("A00", "A10", "Ap"):
        F, M = full.scipy(name), mine.scipy(name)
        r0 = mine.u_begin if name == "A00" else mine.p_begin
        d = abs(F[r0:r0 + M.shape[0], :] - M)
        assert (d.max() if d.nnz else 0.0) <= 1e-13
        cbeg = [0, full.n_u_global // 2, full.n_u_global] if name != "Ap" else [0, full.n_p_global // 2, full.n_p_global]
    dist.barrier()
    if rank == 0:
        print("DIST CPU OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
