"""Alias of fenapack_b200.utils (the reference exposes these helpers as ``fenapack.utils``)."""
from fenapack_b200.utils import *  # noqa: F401,F403
from fenapack_b200.utils import allow_only_one_call  # noqa: F401
