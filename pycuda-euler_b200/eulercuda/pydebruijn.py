"""debruijn -- drop-in for pycuda-euler's ``pydebruijn`` (src/eulercuda/pydebruijn.py)."""
import logging

import numpy as np

import _native
from .pyencode import getOptimalLaunchConfiguration as _launch_cfg

module_logger = logging.getLogger('eulercuda.pydebruijn')


def _l_from_mask(valid_bitmask):
    return bin(int(valid_bitmask)).count("1") // 2 + 1


def debruijn_count_device(d_lmerKeys, d_lmerValues, lmerCount, d_TK, d_TV, d_bucketSize, bucketCount,
                          d_lcount, d_ecount, valid_bitmask, readLength):
    """pydebruijn.py:16 -- leaving / entering multiplicity slots (4 per vertex)."""
    n = int(lmerCount)
    nv = len(d_lcount) // 4
    return _native.default_context().debruijn_count(np.asarray(d_lmerKeys, dtype=np.uint64)[:n],
                                                    np.asarray(d_lmerValues, dtype=np.uint32)[:n],
                                                    d_TK, d_TV, _l_from_mask(valid_bitmask), nv)


def setup_vertices_device(d_kmerKeys, kmerCount, d_TK, d_TV, d_bucketSeed, bucketCount, d_ev, d_lcount, d_lstart,
                          d_ecount, d_estart):
    """pydebruijn.py:182 -- EulerVertex records in vertex-id order."""
    n = int(kmerCount)
    return _native.default_context().setup_vertices(np.asarray(d_kmerKeys, dtype=np.uint64)[:n], d_TK, d_TV,
                                                    d_lcount, d_lstart, d_ecount, d_estart)


def setup_edges_device(d_lmerKeys, d_lmerValues, d_lmerOffsets, lmerCount, d_TK, d_TV, d_bucketSeed, bucketCount,
                       d_l, d_e, d_ee, d_lstart, d_estart, validBitMask):
    """pydebruijn.py:327 -- EulerEdge records and the per-vertex leaving / entering edge lists."""
    n = int(lmerCount)
    vals = np.asarray(d_lmerValues, dtype=np.uint32)[:n]
    ecount = int(vals.sum(dtype=np.uint64))
    return _native.default_context().setup_edges(np.asarray(d_lmerKeys, dtype=np.uint64)[:n], vals,
                                                 np.asarray(d_lmerOffsets, dtype=np.uint32)[:n], d_TK, d_TV,
                                                 _l_from_mask(validBitMask), d_lstart, d_estart, ecount)


def construct_debruijn_graph_device(d_lmerKeys, d_lmerValues, lmerCount, d_kmerKeys, kmerCount, l, d_TK, d_TV,
                                    d_bucketSize, bucketCount, d_ev, d_l, d_e, d_ee, readLength):
    """pydebruijn.py:516 -- returns (d_ee, d_ev, d_l, d_e, kmerCount, edgeCount) with
    edgeCount = sum of multiplicities (B6)."""
    module_logger.info("started construct_debruijn_graph_device.")
    ctx = _native.default_context()
    nl, nk = int(lmerCount), int(kmerCount)
    lkeys = np.asarray(d_lmerKeys, dtype=np.uint64)[:nl]
    lvals = np.asarray(d_lmerValues, dtype=np.uint32)[:nl]
    kkeys = np.asarray(d_kmerKeys, dtype=np.uint64)[:nk]
    valid_bitmask = (1 << (2 * (int(l) - 1))) - 1
    lcount, ecount = ctx.debruijn_count(lkeys, lvals, d_TK, d_TV, int(l), nk)
    lstart = ctx.exclusive_scan(lcount)
    estart = ctx.exclusive_scan(ecount)
    loffs = ctx.exclusive_scan(lvals)
    edge_count = int(lvals.sum(dtype=np.uint64))
    ev = ctx.setup_vertices(kkeys, d_TK, d_TV, lcount, lstart, ecount, estart)
    ee, lev, ent = ctx.setup_edges(lkeys, lvals, loffs, d_TK, d_TV, int(l), lstart, estart, edge_count)
    module_logger.info('Finished construct_debruijn_graph_device.')
    return ee, ev, lev, ent, nk, edge_count


def getOptimalLaunchConfiguration(threadCount, threadPerBlock=32):
    """pydebruijn.py:623"""
    return _launch_cfg(threadCount, threadPerBlock)
