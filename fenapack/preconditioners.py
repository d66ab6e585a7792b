"""Alias of fenapack_b200.preconditioners (same module path as the reference's fenapack/preconditioners.py)."""
from fenapack_b200.preconditioners import *  # noqa: F401,F403
import fenapack_b200.preconditioners as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
