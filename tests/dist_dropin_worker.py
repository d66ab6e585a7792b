"""Multi-rank worker of the drop-in API (launched with torchrun, one rank per GPU -- or on the CPU
with --host-only for the host logic): ``PCDKSP(comm)`` / ``PCDKrylovSolver`` / ``PCDNewtonSolver``
on row-partitioned operators, as the reference runs them under mpirun
(fenapack/field_split.py:46,71-77; SubfieldBC.h:138-140), checked against the serial run."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fenapack_b200 as fp  # noqa: E402
from fenapack_b200.field_split import dofmap_dofs_is  # noqa: E402
from fenapack_b200.field_split_backend import PCDInterface  # noqa: E402
from fenapack_b200.petsc_shim import IS, Mat, Options, TorchDistComm, Vec  # noqa: E402
from fem_forms import BFSModel, PartitionedModel  # noqa: E402


def host_logic(comm, rank, world):
    """No device: distributed sub-matrix extraction, split offsets, BC index mapping, assembler."""
    m = BFSModel(level=1, variant="BRM2")
    pm = PartitionedModel(m, rank, world)
    r0, r1 = pm.rows
    asm = fp.PCDAssembler(pm.a, pm.L, [], ap=pm.ap, kp=pm.kp, mp=pm.mp, bcs_pcd=pm.bc_pcd, function_space=pm.W)
    A = Mat(comm=comm)
    asm.system_matrix(A)
    assert A.getOwnershipRange() == (r0, r1) and A.getSize() == (m.N, m.N)
    is_u = dofmap_dofs_is(pm.W.sub(0).dofmap(), comm)
    is_p = dofmap_dofs_is(pm.W.sub(1).dofmap(), comm)
    assert is_u.getSize() == m.is_u.size and is_p.getSize() == m.is_p.size
    # block extraction == rows of the serial extraction
    Aser = Mat(m.a())
    su, sp_ = IS(m.is_u), IS(m.is_p)
    ub, pb = comm.exscan(is_u.getLocalSize()), comm.exscan(is_p.getLocalSize())
    for (ir, ic, sr, sc, beg) in ((is_u, is_u, su, su, ub), (is_u, is_p, su, sp_, ub), (is_p, is_u, sp_, su, pb)):
        loc = A.createSubMatrix(ir, ic)
        ref = Aser.createSubMatrix(sr, sc).csr[beg:beg + ir.getLocalSize(), :]
        assert loc.getOwnershipRange() == (beg, beg + ir.getLocalSize())
        d = abs(loc.csr - ref)
        assert loc.csr.shape == ref.shape and (d.max() if d.nnz else 0.0) == 0.0
    # PCD operators: Ap with the symmetric Dirichlet rows/columns, partitioned == serial rows
    aser = fp.PCDAssembler(m.a, m.L, [], ap=m.ap, kp=m.kp, mp=m.mp, bcs_pcd=m.bc_pcd, function_space=m.W)
    Ap, Aps = Mat(comm=comm), Mat()
    asm.ap(Ap)
    aser.ap(Aps)
    d = abs(Ap.csr - Aps.csr[r0:r1, :])
    assert (d.max() if d.nnz else 0.0) == 0.0
    # BC indices: local positions, and global = local + exscan (SubfieldBC.h:138-140)
    itf = PCDInterface(asm, A, is_u, is_p)
    idx, vals = itf.pcd_bc_indices()
    gidx, _ = itf.pcd_bc_indices_global()
    assert np.array_equal(is_p.getIndices()[idx], np.intersect1d(m.bc_pcd.dofs(), is_p.getIndices()))
    all_g = np.concatenate(comm.allgather(gidx))
    ser_idx, _ = PCDInterface(aser, Aser, su, sp_).pcd_bc_indices()
    assert np.array_equal(np.sort(all_g), np.sort(ser_idx))
    assert Vec(np.ones(r1 - r0), comm).norm() == np.sqrt(m.N)


def device_run(comm, rank, world, variant):
    """PCDNewtonSolver (Picard steps) on 2+ GPUs against the serial direct solve."""
    Options.clear()
    m = BFSModel(level=3, variant=variant)
    pm = PartitionedModel(m, rank, world)
    r0, r1 = pm.rows
    o = Options("")
    extra = [tuple(kv.split("=")) for kv in os.environ.get("FNP_TEST_OPTS", "").split(",") if kv]      # bisecting aid
    for k, v in extra + [("ksp_gmres_restart", 150), ("fieldsplit_p_pc_python_type", "fenapack.PCDPC_" + variant),
                 ("fieldsplit_u_ksp_type", "richardson"), ("fieldsplit_u_ksp_max_it", 1), ("fieldsplit_u_pc_type", "hypre"),
                 ("fieldsplit_p_PCD_Ap_ksp_type", "richardson"), ("fieldsplit_p_PCD_Ap_ksp_max_it", 2),
                 ("fieldsplit_p_PCD_Ap_pc_type", "hypre"), ("fieldsplit_p_PCD_Mp_ksp_type", "chebyshev"),
                 ("fieldsplit_p_PCD_Mp_ksp_max_it", 5), ("fieldsplit_p_PCD_Mp_ksp_chebyshev_eigenvalues", "0.5, 2.0"),
                 ("fieldsplit_p_PCD_Mp_pc_type", "jacobi")]:
        o.setValue(k, v)
    asm = fp.PCDAssembler(pm.a, pm.L, [], pm.a_pc, ap=pm.ap, kp=pm.kp, mp=pm.mp, bcs_pcd=pm.bc_pcd, function_space=pm.W)
    linear_solver = fp.PCDKrylovSolver(comm=comm)
    linear_solver.parameters["relative_tolerance"] = 1e-8
    linear_solver.parameters["maximum_iterations"] = int(os.environ.get("FNP_TEST_MAXIT", "600"))   # a stagnating solve fails fast
    if "FNP_TEST_MAXIT" in os.environ:
        linear_solver.parameters["error_on_nonconvergence"] = False
    linear_solver.set_from_options()
    solver = fp.PCDNewtonSolver(linear_solver)
    solver.parameters["relative_tolerance"] = 1e-6
    solver.parameters["maximum_iterations"] = 3
    solver.parameters["error_on_nonconvergence"] = False
    problem = fp.PCDNonlinearProblem(asm)
    # the iterate all forms are evaluated at is replicated (host assembly is the test harness);
    # the solver sees this rank's part
    x_loc = Vec(np.zeros(r1 - r0), comm)

    class Problem(fp.PCDNonlinearProblem):
        def F(self, b, x):
            m.w.array[:] = np.concatenate(comm.allgather(x.array))
            super().F(b, x)
            nb = b.norm()                      # collective: every rank
            if rank == 0:
                print(f"[dropin] F assembled, |b| = {nb:.3e}, krylov so far {solver.krylov_iterations()}", flush=True)
                k = linear_solver.ksp()
                if k.device_context() is not None:
                    h = k.getConvergenceHistory()
                    print(f"[dropin] last solve: its {k.getIterationNumber()} reason {k.getConvergedReason()} "
                          f"history {np.array2string(h[:4], precision=3)} ... {np.array2string(h[-4:], precision=3)}", flush=True)
    its, _ = solver.solve(Problem(asm), x_loc)
    # reference: the same Picard steps with direct solves
    ref = BFSModel(level=3, variant=variant)
    for _ in range(its):
        J, b = ref._system()
        dx = spla.spsolve(J.tocsc(), b)
        ws = np.concatenate([ref.w.array[ref.is_u], ref.w.array[ref.is_p]]) - dx
        ref.w.array[ref.is_u] = ws[:ref.n_u]
        ref.w.array[ref.is_p] = ws[ref.n_u:]
    err = np.linalg.norm(x_loc.array - ref.w.array[r0:r1]) ** 2
    tot = np.linalg.norm(ref.w.array[r0:r1]) ** 2
    rel = np.sqrt(comm.allreduce(err) / comm.allreduce(tot))
    assert rel <= 1e-5, rel
    ksp = linear_solver.ksp()
    assert ksp.getConvergedReason() > 0 and ksp.device_context() is not None
    return its, solver.krylov_iterations(), rel


def main():
    import faulthandler
    # a hang (a collective one rank never reaches) reports where it is stuck and ends the run
    faulthandler.dump_traceback_later(int(os.environ.get("FNP_TEST_HANG_S", "240")), exit=True)
    host_only = "--host-only" in sys.argv
    variant = next((a for a in sys.argv[1:] if a.startswith("BRM")), "BRM1")
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if host_only:
        dist.init_process_group("gloo")
    else:
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = TorchDistComm()
    host_logic(comm, rank, world)
    msg = "DROPIN HOST OK"
    if not host_only:
        its, kits, rel = device_run(comm, rank, world, variant)
        msg = f"DROPIN DIST OK world={world} variant={variant} picard={its} krylov={kits} relerr={rel:.2e}"
    dist.barrier()
    if rank == 0:
        print(msg)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
