"""eulercuda package -- B200-native drop-in for pycuda-euler's ``src/eulercuda/`` package.

The reference package has an empty ``__init__`` (callers import ``eulercuda.eulercuda as ec``,
cli_spark_gpu.py:21); the entry points are re-exported here so ``import eulercuda as ec`` works
too (SURVEY §2.1 shadowing note)."""
from .eulercuda import (assemble, assemble2, constructDebruijnGraph, readLmersKmersCuda, findEulerTour,  # noqa: F401
                        generatePartialContig, findSpanningTree, getString, read_fasta, read_fastq, parse_fastq,
                        dna_translate, doErrorCorrection, verify_kmers, check_kmers)
