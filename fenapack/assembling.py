"""Alias of fenapack_b200.assembling (same module path as the reference's fenapack/assembling.py)."""
from fenapack_b200.assembling import *  # noqa: F401,F403
import fenapack_b200.assembling as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
