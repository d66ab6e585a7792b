"""Small helpers kept from the reference's utility module (fenapack/utils.py)."""
import functools


def allow_only_one_call(func):
    """Decorator: a second call of the decorated method on the same object raises
    RuntimeError.  Same contract as the reference helper (fenapack/utils.py:37-60),
    which guards ``PCDKSP.init_pcd`` (fenapack/field_split.py:60)."""
    flag = "_called_once_" + func.__name__

    @functools.wraps(func)
    def wrapper(self, *args, **kwargs):
        if getattr(self, flag, False):
            raise RuntimeError("Multiple calls to '%s' not allowed" % func.__name__)
        setattr(self, flag, True)
        return func(self, *args, **kwargs)
    return wrapper
