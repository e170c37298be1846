"""src/debruijn/pydebruijn.py layout: same module as eulercuda.pydebruijn."""
from eulercuda.pydebruijn import *  # noqa: F401,F403
from eulercuda import pydebruijn as _m
__all__ = [n for n in dir(_m) if not n.startswith("__")]
globals().update({n: getattr(_m, n) for n in __all__})
