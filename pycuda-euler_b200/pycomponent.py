"""flat src/pycomponent.py layout: same module as eulercuda.pycomponent."""
from eulercuda import pycomponent as _m
globals().update({n: getattr(_m, n) for n in dir(_m) if not n.startswith("__")})
