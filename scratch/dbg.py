import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import fenapack_b200 as fp
from fenapack_b200.petsc_shim import Options, Mat, Vec
from fem_forms import BFSModel
from test_api_dropin import make_assembler, set_iterative_options
Options.clear()
m = BFSModel(level=2, variant="BRM1")
set_iterative_options("", "BRM1")
ls = fp.PCDKrylovSolver(); ls.parameters["relative_tolerance"] = 1e-6; ls.parameters["maximum_iterations"]=300; ls.parameters["error_on_nonconvergence"]=False
ls.set_from_options()
problem = fp.PCDNonlinearProblem(make_assembler(m))
A = Mat(); b = Vec(np.zeros(m.N)); dx = Vec(np.zeros(m.N))
problem.F(b, m.w); problem.J(A, m.w)
ls.set_operators(A, A); ls.init_pcd(problem.pcd_assembler)
its = ls.solve(dx, b)
ksp = ls.ksp()
print("its", its, "hist", ksp.getConvergenceHistory()[:10], ksp.getConvergenceHistory()[-3:])
import scipy.sparse.linalg as spla
ref = spla.spsolve(A.csr.tocsc(), b.array)
print("err vs direct", np.linalg.norm(dx.array-ref)/np.linalg.norm(ref), "bnorm", np.linalg.norm(b.array))
ctx = ksp.device_context()
# compare blocks with oracle
from oracle import problems, petsc_algos as pa
p0,_ = problems.backward_facing_step(2, variant="BRM1")
print("rhs diff", np.linalg.norm(b.array[m.is_u] + p0.b_u), np.linalg.norm(b.array[m.is_p]+p0.b_p))
x = np.random.default_rng(0).standard_normal(p0.n_p)
print("schur", np.linalg.norm(ctx.schur_apply(x)))
print("Kp nnz", ksp._pcd_pc.mat_Kp.csr.nnz, "Ap diag min", ksp._pcd_pc.mat_Ap.csr.diagonal().min(), "Mp", ksp._pcd_pc.mat_Mp.csr.diagonal().min())
