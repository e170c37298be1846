"""TEST / BENCH INFRASTRUCTURE ONLY.  Load the UNMODIFIED reference CPU assembler
(src/referenceassembler/referenceAssembler.py): from /root/reference when it exists (the authoring container),
else from the bytecode oracle/build_ref.py left in oracle/_ref/ (the GPU box).  Its one third-party import
(`from dask import delayed`, referenceAssembler.py:5) is unused and satisfied by a stub."""
import importlib.machinery
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/referenceassembler/referenceAssembler.py"
PYC = os.path.join(HERE, "_ref", "referenceAssembler.bytecode")
_mod = None


def load():
    """the reference module, or None when it is not available in either form"""
    global _mod
    if _mod is not None:
        return _mod
    sys.modules.setdefault("dask", types.SimpleNamespace(delayed=lambda f=None, **kw: f))
    if os.path.exists(SRC):
        loader = importlib.machinery.SourceFileLoader("reference_referenceAssembler", SRC)
    elif os.path.exists(PYC):
        loader = importlib.machinery.SourcelessFileLoader("reference_referenceAssembler", PYC)
    else:
        return None
    spec = importlib.util.spec_from_loader(loader.name, loader)
    mod = importlib.util.module_from_spec(spec)
    try:
        loader.exec_module(mod)
    except Exception:
        return None
    _mod = mod
    return mod


def kind():
    return "source" if os.path.exists(SRC) else ("bytecode" if os.path.exists(PYC) else None)
