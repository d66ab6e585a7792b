"""Generates the golden fixtures of tests/golden/ from the CPU oracle.

The reference (blechta/fenapack) holds no golden vectors for this path and cannot be
imported here (no DOLFIN / petsc4py), so these fixtures pin the ORACLE, not the
reference: they guard the restatement against accidental change and give the GPU
tests a committed, size-independent target.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import amg, petsc_algos as pa, problems  # noqa: E402

out = {}
for variant in ("BRM1", "BRM2"):
    p0, _ = problems.backward_facing_step(2, variant=variant)
    x = pa.direct_solver(p0.system_matrix())(p0.rhs())
    prob, _ = problems.backward_facing_step(2, variant=variant, wind=x[:p0.n_u].reshape(-1, 2), stabilise=True)
    rng = np.random.default_rng(2024)
    xu, xp = rng.standard_normal(prob.n_u), rng.standard_normal(prob.n_p)
    # direct inner solves: depends on nothing but the reference's own formulas
    pcd = pa.PCDPreconditioner(prob, "direct")
    yu, yp = pcd.apply_split(xu, xp)
    xs, its, hist, _ = pa.fgmres(prob.system_matrix(), pcd, prob.rhs(), rtol=1e-6, restart=150)
    out[f"{variant}_direct_yu"], out[f"{variant}_direct_yp"] = yu, yp
    out[f"{variant}_direct_its"] = np.array([its])
    out[f"{variant}_direct_hist"] = np.array(hist)
    # iterative inner solves (Chebyshev-Jacobi, SA-AMG V-cycles)
    P00 = prob.P00 if prob.P00 is not None else prob.A00
    pci = pa.PCDPreconditioner(prob, "iterative", amg_u=amg.build_hierarchy(P00), amg_p=amg.build_hierarchy(prob.Ap))
    yu, yp = pci.apply_split(xu, xp)
    xs, its, hist, _ = pa.fgmres(prob.system_matrix(), pci, prob.rhs(), rtol=1e-6, restart=150)
    out[f"{variant}_iter_yu"], out[f"{variant}_iter_yp"] = yu, yp
    out[f"{variant}_iter_its"] = np.array([its])
    out[f"{variant}_cheb"] = pa.chebyshev_jacobi(prob.Mp, 1.0 / prob.Mp.diagonal(), xp, 0.5, 2.0, 5)
    out[f"{variant}_xu"], out[f"{variant}_xp"] = xu, xp
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bfs_l2_pcd.npz"), **out)
print("written", {k: v.shape for k, v in out.items()})
