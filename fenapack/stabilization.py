"""Alias of fenapack_b200.stabilization (same module path as the reference's fenapack/stabilization.py)."""
from fenapack_b200.stabilization import *  # noqa: F401,F403
import fenapack_b200.stabilization as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
