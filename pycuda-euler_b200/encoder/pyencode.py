"""src/encoder/pyencode.py layout: same module as eulercuda.pyencode."""
from eulercuda.pyencode import *  # noqa: F401,F403
from eulercuda import pyencode as _m
__all__ = [n for n in dir(_m) if not n.startswith("__")]
globals().update({n: getattr(_m, n) for n in __all__})
