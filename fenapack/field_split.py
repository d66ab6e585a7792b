"""Alias of fenapack_b200.field_split (same module path as the reference's fenapack/field_split.py)."""
from fenapack_b200.field_split import *  # noqa: F401,F403
import fenapack_b200.field_split as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
