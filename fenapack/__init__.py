"""Drop-in alias: PETSc resolves the Schur-complement PC by dotted name
(``-fieldsplit_p_pc_python_type fenapack.PCDPC_BRM1``,
demo_navier-stokes-pcd.py:151; looked up at fenapack/field_split.py:109-114), so
the public names of the reference package (fenapack/__init__.py:35-40) must be
importable as ``fenapack.<name>``.  Everything lives in ``fenapack_b200``."""
from fenapack_b200 import (PCDKSP, PCDAssembler, PCDForm, PCDKrylovSolver,  # noqa: F401
                           PCDNewtonSolver, PCDNonlinearProblem, PCDPC_BRM1, PCDPC_BRM2,
                           PCDRPC_BRM1, PCDRPC_BRM2, StabilizationParameterSD, __version__)
