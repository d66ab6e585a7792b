"""Selects the PETSc binding: real petsc4py when importable, otherwise the
duck-typed stand-in (fenapack_b200/petsc_shim.py).  Only host-side glue objects
(Vec/Mat/IS/Options) come from here -- never the arithmetic of the hot path."""
try:  # pragma: no cover - petsc4py is not installed in the build image
    from petsc4py import PETSc  # type: ignore
    HAVE_PETSC4PY = True
except Exception:
    from .petsc_shim import PETSc
    HAVE_PETSC4PY = False
