"""flat src/pyeulertour.py layout: same module as eulercuda.pyeulertour."""
from eulercuda import pyeulertour as _m
globals().update({n: getattr(_m, n) for n in dir(_m) if not n.startswith("__")})
