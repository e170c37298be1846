"""src/gpuhash/pygpuhash.py layout: same module as eulercuda.pygpuhash."""
from eulercuda.pygpuhash import *  # noqa: F401,F403
from eulercuda import pygpuhash as _m
__all__ = [n for n in dir(_m) if not n.startswith("__")]
globals().update({n: getattr(_m, n) for n in __all__})
