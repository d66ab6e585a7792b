"""fenapack_b200 -- B200-native implementation of FENaPack's hot path: the PCD
block-triangular preconditioner inside right-preconditioned (F)GMRES for the
P2/P1 Oseen / Navier-Stokes system.

``capi``            ctypes binding of libfenapack_cuda.so (the C ABI, include/fenapack_cuda.h)
``preconditioners`` PCDPC_BRM1 / PCDPC_BRM2 python-PC contexts (reference: fenapack/preconditioners.py)
``field_split``     PCDKSP / PCDKrylovSolver / PCDKSPPython   (reference: fenapack/field_split.py; KSPPYTHON
                    context = the entry point for a SNES that owns its KSP, demo/defcon/navier-stokes.py:252-280)
``field_split_backend`` PCDInterface                          (reference: fenapack/field_split_backend.py)
``assembling``      PCDAssembler / PCDForm                    (reference: fenapack/assembling.py)
``nonlinear_solvers`` PCDNewtonSolver / PCDNonlinearProblem   (reference: fenapack/nonlinear_solvers.py)
``stabilization``   StabilizationParameterSD (host helper)    (reference: fenapack/stabilization.py)

The CUDA library is mandatory; nothing here falls back to the CPU.
"""
__version__ = "0.1.0"

from .assembling import PCDAssembler, PCDForm  # noqa: E402,F401
from .field_split import PCDKSP, PCDKSPPython, PCDKrylovSolver  # noqa: E402,F401
from .nonlinear_solvers import PCDNewtonSolver, PCDNonlinearProblem  # noqa: E402,F401
from .preconditioners import PCDPC_BRM1, PCDPC_BRM2, PCDRPC_BRM1, PCDRPC_BRM2  # noqa: E402,F401
from .stabilization import StabilizationParameterSD  # noqa: E402,F401
