"""Newton/Picard driver around the PCD-preconditioned linear solver -- drop-in for
fenapack/nonlinear_solvers.py.  The reference subclasses DOLFIN's C++
``NewtonSolver``; the loop below restates that solver's behaviour for the calls
the reference overrides: ``F``, ``J``, ``J_pc`` per iteration, ``solver_setup``
whose body runs only for the first iteration (nonlinear_solvers.py:63-78), then
one linear solve and the update ``x -= relaxation * dx``.  It is the *caller* of
the hot path and decides when the value refresh happens."""
from __future__ import annotations

import numpy as np

from ._backend import PETSc


class PCDNonlinearProblem(object):
    """Nonlinear problem fed from a ``PCDAssembler`` (reference :88-112)."""

    def __init__(self, pcd_assembler):
        self.pcd_assembler = pcd_assembler

    def F(self, b, x):
        self.pcd_assembler.rhs_vector(b, x)

    def J(self, A, x):
        self.pcd_assembler.system_matrix(A)

    def J_pc(self, P, x):
        return self.pcd_assembler.pc_matrix(P)


class PCDNewtonSolver(object):
    def __init__(self, solver, pcd_pc_class=None):
        """``solver``: a ``PCDKrylovSolver``; ``pcd_pc_class``: optional PCD PC class
        handed to ``init_pcd`` (reference :33-50)."""
        self._solver = solver
        self._pcd_pc_class = pcd_pc_class
        self.parameters = {"relative_tolerance": 1e-9, "absolute_tolerance": 1e-10, "maximum_iterations": 50,
                           "relaxation_parameter": 1.0, "error_on_nonconvergence": True}
        self._krylov_iterations = 0
        self._A = PETSc.Mat()
        self._P = PETSc.Mat()

    def linear_solver(self):
        return self._solver

    def krylov_iterations(self):
        """Accumulated Krylov iterations (reference fenapack/__init__.py:44-56)."""
        return self._krylov_iterations

    def solver_setup(self, A, P, nonlinear_problem, iteration):
        # Only do the setup once
        if iteration > 0 or getattr(self, "_initialized", False):
            return
        self._initialized = True
        self._solver.set_operators(A, P if (P is not None and P.isAssembled()) else A)
        self._solver.init_pcd(nonlinear_problem.pcd_assembler, self._pcd_pc_class)

    def solve(self, problem, x):
        """Solve F(x) = 0.  Returns (iterations, converged)."""
        self._problem = problem
        prm = self.parameters
        n = x.getLocalSize()
        comm = getattr(x, "comm", None)
        b = PETSc.Vec(np.zeros(n), comm)
        dx = PETSc.Vec(np.zeros(n), comm)
        self._A.comm = self._P.comm = comm if comm is not None else self._A.comm
        problem.F(b, x)
        r0 = b.norm()
        converged = r0 <= prm["absolute_tolerance"]
        it = 0
        while not converged and it < prm["maximum_iterations"]:
            problem.J(self._A, x)
            P = problem.J_pc(self._P, x)
            self.solver_setup(self._A, P, problem, it)
            self._krylov_iterations += self._solver.solve(dx, b)
            x.axpy(-prm["relaxation_parameter"], dx)
            it += 1
            problem.F(b, x)
            r = b.norm()
            converged = r <= prm["absolute_tolerance"] or r <= prm["relative_tolerance"] * r0
        if not converged and prm["error_on_nonconvergence"]:
            raise RuntimeError("PCDNewtonSolver did not converge in %d iterations" % it)
        return it, converged
