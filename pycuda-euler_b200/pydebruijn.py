"""flat src/pydebruijn.py layout: same module as eulercuda.pydebruijn."""
from eulercuda import pydebruijn as _m
globals().update({n: getattr(_m, n) for n in dir(_m) if not n.startswith("__")})
