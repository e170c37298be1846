"""Drop-in for the reference's ``src/referenceassembler`` package: same functions, the
counting / unitig / link-graph work done on the GPU through libeuler_b200 (no CPU fallback)."""
from .referenceAssembler import (twin, kmers, fw, bw, build, contig_to_string, get_contig, get_contig_forward,  # noqa: F401
                                 all_contigs, print_GFA, print_dbg, write_gfa, write_fasta, runAssembler)
