"""flat src/pyencode.py layout: same module as eulercuda.pyencode."""
from eulercuda import pyencode as _m
globals().update({n: getattr(_m, n) for n in dir(_m) if not n.startswith("__")})
