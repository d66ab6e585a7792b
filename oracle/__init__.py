"""CPU oracle for the PCD-preconditioned FGMRES hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the shipped product (``fenapack_b200/``) imports this package.  It may
be imported only by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and there only as
the checker or as the CPU arm that is timed next to the GPU, never as the thing
shipped.

PARITY: PINNED FOR THE REFERENCE-OWNED PART, UNPINNED FOR THE PETSc-OWNED PART.
The reference (blechta/fenapack, /root/reference) *wires* third-party solvers
together (PETSc KSP/PC/Mat/Vec, hypre BoomerAMG, MUMPS, DOLFIN assembly); none of
them is vendored, none is installed in this image, no version is pinned
(README.rst:26 names only "FEniCS 2019.2.0.dev0"), and the reference's own tests
assert nothing on this path except "the solve converged"
(test/bench/test_pcd_scaling.py:223): it holds no golden vectors.
  * pinned: what the reference's own Python computes -- the four Schur-complement
    applies PCDPC_BRM1/2 and PCDRPC_BRM1/2 (order of operations, signs, where the
    subfield BC is applied), PCDInterface's operator extraction / refresh protocol
    and its Rp = B diag(Mu)^-1 B^T.  tests/golden/make_reference_golden.py executes
    fenapack/preconditioners.py and field_split_backend.py UNMODIFIED in this
    container (numpy/scipy stand-ins for the petsc4py objects they call, Cholesky
    inner solves = the reference's default) and commits the outputs as
    tests/golden/ref_pcd_apply.npz; tests/test_reference_golden.py holds the oracle
    to them at 1e-11 and the CUDA library at 1e-8.
  * unpinned: the algorithms that live inside PETSc / hypre (KSPGMRES, PCFIELDSPLIT
    Schur/upper, KSPCHEBYSHEV, KSPRICHARDSON; BoomerAMG is replaced by SA-AMG on both
    sides).  They are restated from the published algorithms ("recalled" details:
    SURVEY.md appendix B) and held by the self-consistency properties of SURVEY.md
    section 8c (exact-Schur two-iteration convergence, Chebyshev polynomial
    optimality, convergence of all 16 scenario combinations of the reference bench,
    iteration counts in the neighbourhood of the un-asserted table in
    demo/unsteady-navier-stokes-pcd/documentation.rst:137).

Modules
-------
fem         own P2/P1 simplex assembler (stands in for DOLFIN assembly, which
            stays on the host in the product as well)
problems    the scenario builders: BFS L-shape (demo/data/mesh_lshape.xml),
            lid-driven cavities, 3D channel
petsc_algos restatement of the PETSc algorithm chain the reference selects
            (KSPGMRES right-preconditioned, PCFIELDSPLIT Schur/upper,
            KSPCHEBYSHEV+PCJACOBI, KSPRICHARDSON, KSPCG) and of
            fenapack/preconditioners.py BRM1/BRM2
amg         smoothed-aggregation AMG (stands in for hypre BoomerAMG, which
            cannot be reproduced bit-wise; the product builds the *same*
            hierarchy so V-cycles can be compared to rounding)
"""
