"""Alias of fenapack_b200.nonlinear_solvers (same module path as the reference's fenapack/nonlinear_solvers.py)."""
from fenapack_b200.nonlinear_solvers import *  # noqa: F401,F403
import fenapack_b200.nonlinear_solvers as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
