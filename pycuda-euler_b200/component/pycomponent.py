"""src/component/pycomponent.py layout: same module as eulercuda.pycomponent."""
from eulercuda.pycomponent import *  # noqa: F401,F403
from eulercuda import pycomponent as _m
__all__ = [n for n in dir(_m) if not n.startswith("__")]
globals().update({n: getattr(_m, n) for n in __all__})
