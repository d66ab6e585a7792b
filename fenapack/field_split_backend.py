"""Alias of fenapack_b200.field_split_backend (same module path as the reference's fenapack/field_split_backend.py)."""
from fenapack_b200.field_split_backend import *  # noqa: F401,F403
import fenapack_b200.field_split_backend as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
