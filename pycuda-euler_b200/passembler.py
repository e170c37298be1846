"""passembler.py -- the reference's parallel-assembler entry point (a dask stub there, src/passembler.py):
here it is the partition driver; see assembler.py."""
from assembler import main, split_records, detect_format  # noqa: F401

if __name__ == "__main__":
    import sys
    sys.exit(main())
