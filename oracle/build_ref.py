#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE ONLY.  Byte-compile the UNMODIFIED reference CPU assembler
(/root/reference/src/referenceassembler/referenceAssembler.py) into oracle/_ref/referenceAssembler.bytecode.

The reference path is pure Python: there is nothing for gcc to build.  Its compiled form is CPython bytecode,
so that is what goes into oracle/_ref/ (git-ignored, but shipped to the GPU box like the built .so files): the
box has no /root/reference, and `bench.py --impl reference` / `cpu_baseline` must time the reference's own
`build()` there, not a port.  No reference SOURCE is copied into the repository.
Run from the repo root:  python oracle/build_ref.py   (also called by __graft_entry__.build()).
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/referenceassembler/referenceAssembler.py"
OUT = os.path.join(HERE, "_ref", "referenceAssembler.bytecode")


def build(force=False):
    """returns the path of the bytecode file, or None when neither the reference nor a previous build is present"""
    if os.path.exists(SRC):
        if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
            os.makedirs(os.path.dirname(OUT), exist_ok=True)
            py_compile.compile(SRC, cfile=OUT, dfile="referenceAssembler.py", doraise=True, optimize=0)
    return OUT if os.path.exists(OUT) else None


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p or "reference not available")
