"""eulercuda -- drop-in for pycuda-euler's orchestrator (src/eulercuda/eulercuda.py).

Same entry points (``assemble2`` / ``assemble``, ``constructDebruijnGraph``, ``readLmersKmersCuda``,
``findEulerTour``, ``generatePartialContig``, ``findSpanningTree``, the readers and ``getString``),
same argument order and return order.  All compute runs in libeuler_b200.so; the two host loops
the reference left on the CPU (the dict fill :141-178 and the contig walk :351-402) and its
graph_tool spanning tree (:267-306) are device kernels here.
"""
import argparse
import logging
import os

import numpy as np

import _native
from . import pyencode as enc          # noqa: F401
from . import pygpuhash as gh          # noqa: F401
from . import pydebruijn as db         # noqa: F401
from . import pyeulertour as et

ULONGLONG = 8
UINTC = 4


def parse_fastq(filename):
    """eulercuda.py:23 -- {read name: read} of a 4-line-record FASTQ file."""
    result = {}
    current_name = None
    with open(filename) as handle:
        for i, line in enumerate(handle):
            if i % 4 == 0:
                current_name = line.rstrip('\n')
            elif i % 4 == 1:
                result[current_name] = line.rstrip('\n')
    return result


def read_fastq(filename):
    """eulercuda.py:44 -- list of reads of a 4-line-record FASTQ file."""
    with open(filename, "r") as handle:
        return [line.rstrip('\n') for i, line in enumerate(handle) if i % 4 == 1]


def read_fasta(infilename):
    """eulercuda.py:439 -- every non-'>' line, stripped."""
    with open(infilename, 'r') as handle:
        return [line.strip() for line in handle if line and line[0] != '>']


def doErrorCorrection(readBuffer, readCount, ec_tuple_size, max_ec_pos):
    """eulercuda.py:59 (stub in the reference too)."""
    return readCount


def dna_translate(i):
    """eulercuda.py:308"""
    return 'ACGT'[i] if 0 <= i < 4 else '.'


def getString(length, value):
    """eulercuda.py:315 -- u64 -> ACGT string, MSB-first."""
    v = int(value)
    return ''.join('ACGT'[(v >> (2 * (length - 1 - i))) & 3] for i in range(length))


def verify_kmers(buffer, encoded_list, length):
    """eulercuda.py:62"""
    hits, misses = [], []
    for val in encoded_list:
        s = getString(length, val)
        (hits if s in buffer else misses).append(s)
    return hits, misses


def check_kmers(outfile, length, kmers):
    """eulercuda.py:323 -- tab-separated dump (only on request; no CWD side effects by default)."""
    with open(outfile, 'w') as ofile:
        for kmer in kmers:
            ofile.write(getString(length, kmer) + '\t')


def _flat(readBuffer, readLength, numReads, read_offsets=None):
    buf = enc._as_bytes(readBuffer)
    off = enc._offsets(buf.size, readLength, read_offsets)
    return buf, off


def readLmersKmersCuda(readBuffer, readLength, partitionReadCount, lmerLength, lmerKeys, lmerValues, lmerCount,
                       kmerKeys, kmerValues, kmerCount, numReads, read_offsets=None):
    """eulercuda.py:74 -- [lmerCount, kmerCount, lmerKeys, lmerValues, kmerKeys, kmerValues] over both
    strands: distinct l-mers with multiplicities, distinct k-mers with ids (rank in ascending key
    order, B14).  Poly-A is counted like any other l-mer (B3)."""
    ctx = _native.default_context()
    buf, off = _flat(readBuffer, readLength, numReads, read_offsets)
    ctx.run_host(buf, off, lmerLength, _native.RUN_CANONICAL_IDS)
    lk = ctx.download(_native.ART_LMER_KEYS)
    lv = ctx.download(_native.ART_LMER_VALUES)
    kk = ctx.download(_native.ART_KMER_KEYS)
    kv = np.arange(kk.size, dtype=np.uint32)
    if int(lmerLength) > 32:
        # two-word keys (csrc/wide.cu): hand them back as Python integers, like the reference's lists
        lk = (ctx.download(_native.ART_LMER_KEYS_HI).astype(object) << 64) | lk.astype(object)
        kk = (ctx.download(_native.ART_KMER_KEYS_HI).astype(object) << 64) | kk.astype(object)
    return [int(lk.size), int(kk.size), lk, lv, kk, kv]


def constructDebruijnGraph(readBuffer, partitionReadCount, readLength, lmerLength, evList, eeList, levEdgeList,
                           entEdgeList, numReads, read_offsets=None):
    """eulercuda.py:183 -- returns (ee, ev, l, e, vertexCount, edgeCount), the order assemble2
    unpacks (B13).  One fused device-resident pass instead of 4 wrappers x PCIe round trips."""
    ctx = _native.default_context()
    buf, off = _flat(readBuffer, readLength, numReads, read_offsets)
    st = ctx.run_host(buf, off, lmerLength, _native.RUN_CANONICAL_IDS | _native.RUN_EXPAND_EDGES)
    ee = ctx.download(_native.ART_EE)
    ev = ctx.download(_native.ART_EV)
    lev = ctx.download(_native.ART_LEV)
    ent = ctx.download(_native.ART_ENT)
    return ee, ev, lev, ent, int(st.distinct_kmers), int(st.edge_count)


def findSpanningTree(cg_edge, cg_edgecount, cg_vertexcount):
    """eulercuda.py:267 -- spanning forest of the circuit graph as a flat list of circuit-edge
    indices (B10); Boruvka on device, equal to Kruskal in index order."""
    cg = np.asarray(cg_edge, dtype=_native.CE_DTYPE)[:int(cg_edgecount)]
    return _native.default_context().spanning_forest(cg, int(cg_vertexcount))


def generatePartialContig(outfile, d_ev, vcount, d_ee, ecount, l):
    """eulercuda.py:329 -- contigs of the successor chains, written as '>%u\\n<seq>\\n'; returns the
    list of contig strings (first k-mer + one base per edge, B12)."""
    contigs = _native.default_context().emit_contigs(np.asarray(d_ev, dtype=_native.EV_DTYPE)[:int(vcount)],
                                                     np.asarray(d_ee, dtype=_native.EE_DTYPE)[:int(ecount)], int(l))
    if outfile:
        with open(outfile, 'w') as ofile:
            for i, c in enumerate(contigs):
                ofile.write('>%u\n' % i)
                ofile.write(c + '\n')
    return contigs


def findEulerTour(d_ev, d_ee, d_levEdge, d_entEdge, edgeCountList, vertexCount, lmerLength, outfile):
    """eulercuda.py:407 -- successor assignment, circuits, spanning forest, swipe, contigs."""
    ecount, vcount = int(edgeCountList), int(vertexCount)
    if ecount == 0:
        return []
    _, ee = et.assign_successor_device(d_ev, d_levEdge, d_entEdge, vcount, d_ee, ecount)
    cg_edge, cg_edgeCount, cg_vertexCount = et.findEulerDevice(d_ev, d_levEdge, d_entEdge, vcount, d_ee, ecount,
                                                               {}, 0, 0)
    if cg_edgeCount > 0:
        tree = findSpanningTree(cg_edge, cg_edgeCount, cg_vertexCount)
        ee = et.executeSwipeDevice(d_ev, d_entEdge, vcount, ee, ecount, cg_edge, cg_edgeCount, tree, len(tree))
    return generatePartialContig(outfile, d_ev, vcount, ee, ecount, lmerLength)


def assemble2(lmerLength, buffer='', readLength=0, readCount=0, infile='', outfile='', mode='unitig', limit=1):
    """eulercuda.py:450 -- assemble reads into contigs.

    DEVIATION from the reference, stated up front: the reference's ``assemble2`` is the (unfinished) GPU-Euler
    path; here the DEFAULT ``mode='unitig'`` gives the contigs of the reference's own runnable CPU assembler
    (BASELINE configs[0] checks exactly that), and ``mode='euler'`` selects the GPU-Euler path.  ``buffer`` is an
    iterable of reads, or one flat string / bytes object cut every ``readLength`` bases; ``readCount`` > 0 keeps
    the first reads only.

    ``lmerLength`` is the CLI's ``-k`` value (:554).  ``mode='unitig'`` (default) reproduces the
    reference's CPU assembler on the GPU: nodes are ``lmerLength``-mers with both-strand count >
    ``limit`` and the result equals ``referenceAssembler.all_contigs`` as an orientation-free set.
    ``mode='euler'`` is the GPU-Euler path of the port: de Bruijn graph with l-mer edges ->
    Euler tour -> partial contigs."""
    logger = logging.getLogger(__name__)
    if infile != '':
        # the file is parsed on the device (csrc/ingest.cu): the reads never exist on the host
        extension = infile.split('.')[-1]
        if extension in ['fa', 'fasta', 'fsa']:
            fmt = 1
        elif extension in ['fq', 'fastq']:
            fmt = 2
        else:
            raise ValueError("unknown read file extension: %s" % infile)
        with open(infile, 'rb') as handle:
            data = handle.read()
        ctx = _native.default_context()
        nreads, nbases = ctx.ingest(data, fmt)
        logger.info("Got %s reads." % nreads)
        contigs = []
        if nbases > 0:
            if mode == 'unitig':
                contigs = ctx.unitigs_ingested(int(lmerLength), int(limit))
            elif mode == 'euler':
                ctx.run_ingested(int(lmerLength), _native.RUN_CANONICAL_IDS | _native.RUN_EXPAND_EDGES)
                contigs = ctx.pipeline_contigs()
            else:
                raise ValueError("mode must be 'unitig' or 'euler'")
        if outfile:
            with open(outfile, 'w') as ofile:
                for i, c in enumerate(contigs):
                    ofile.write('>%u\n%s\n' % (i, c))
        return contigs
    if isinstance(buffer, (str, bytes, bytearray)):
        # the reference's other calling convention: one flat string of concatenated reads plus readLength
        # (eulercuda.py:485-486 joins the reads before encoding); iterating it would yield 1-base "reads"
        if len(buffer) == 0:
            buffer = []
        elif readLength and int(readLength) > 0:
            rl = int(readLength)
            buffer = [buffer[i:i + rl] for i in range(0, len(buffer), rl)]
        else:
            raise ValueError("assemble2: a flat read buffer needs readLength > 0 (or pass an iterable of reads)")
    reads = [r.decode('ascii') if isinstance(r, (bytes, bytearray)) else str(r) for r in buffer]
    if readCount and int(readCount) > 0:
        reads = reads[:int(readCount)]
    logger.info("Got %s reads." % len(reads))
    data = b''.join(r.encode('ascii') for r in reads)
    buf = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(0, np.uint8)
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    if reads:
        off[1:] = np.cumsum([len(r) for r in reads], dtype=np.uint64)
    contigs = []
    if buf.size > 0:
        ctx = _native.default_context()
        if mode == 'unitig':
            contigs = ctx.unitigs(buf, off, int(lmerLength), int(limit))
        elif mode == 'euler':
            ctx.run_host(buf, off, int(lmerLength), _native.RUN_CANONICAL_IDS | _native.RUN_EXPAND_EDGES)
            contigs = ctx.pipeline_contigs()
        else:
            raise ValueError("mode must be 'unitig' or 'euler'")
    if outfile:
        with open(outfile, 'w') as ofile:
            for i, c in enumerate(contigs):
                ofile.write('>%u\n%s\n' % (i, c))
    return contigs


assemble = assemble2


def main(argv=None):
    """eulercuda.py:508 -- ``python eulercuda.py -i in.fa -o out -k L [-d]``."""
    parser = argparse.ArgumentParser()
    parser.add_argument('-i', action='store', dest='input_filename', help='Input File Name')
    parser.add_argument('-o', action='store', dest='output_filename', default='', help='Output File Name')
    parser.add_argument('-k', action='store', dest='k', type=int, default=0, help='kmer size')
    parser.add_argument('-d', action='store_true', dest='debug', default=False)
    parser.add_argument('--mode', choices=['unitig', 'euler'], default='unitig')
    results = parser.parse_args(argv)
    if results.debug:
        logging.basicConfig(level=logging.DEBUG)
    k = results.k if results.k and results.k > 0 else 21     # :533-536
    contigs = assemble2(k, infile=results.input_filename or '', outfile=results.output_filename or '', mode=results.mode)
    if not results.output_filename:
        for i, c in enumerate(contigs):
            print('>%u\n%s' % (i, c))
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
