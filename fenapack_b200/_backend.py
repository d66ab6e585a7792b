"""Host-side glue objects (Vec / Mat / IS / Options / Comm) of the drop-in classes.

The drop-in layer is written against the duck-typed stand-in of ``petsc_shim`` (the slice of
petsc4py the reference's hot path touches) -- never the arithmetic of the hot
path, which always goes through libfenapack_cuda.  petsc4py is not installable in the build image,
so the real binding has never been exercised: it is NOT selected automatically (the container
constructors and ``Mat.stateGet`` differ from the stand-in's).  ``FENAPACK_B200_PETSC4PY=1`` opts in
for a maintainer who validates that path; INTEGRATION.md lists what to check."""
import os

HAVE_PETSC4PY = False
if os.environ.get("FENAPACK_B200_PETSC4PY") == "1":  # pragma: no cover - petsc4py is not in the build image
    try:
        from petsc4py import PETSc  # type: ignore
        HAVE_PETSC4PY = True
    except Exception:
        from .petsc_shim import PETSc
else:
    from .petsc_shim import PETSc


def mat_state(mat):
    """Object state of a Mat (PetscObjectStateGet): bumps whenever the matrix is re-assembled in
    place.  Raises instead of returning None, so that a binding without the probe cannot silently
    freeze the first Jacobian."""
    if hasattr(mat, "stateGet"):
        return mat.stateGet()
    st = getattr(mat, "state", None)
    if st is None:
        raise RuntimeError("cannot read the object state of %r: value refreshes would go unnoticed" % (mat,))
    return st
