"""flat src/pygpuhash.py layout: same module as eulercuda.pygpuhash."""
from eulercuda import pygpuhash as _m
globals().update({n: getattr(_m, n) for n in dir(_m) if not n.startswith("__")})
