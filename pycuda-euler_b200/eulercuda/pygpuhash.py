"""gpuhash -- drop-in for pycuda-euler's ``pygpuhash`` (src/eulercuda/pygpuhash.py).

The reference builds a static bucketed table (histogram -> scan -> scatter -> per-bucket rank
sort, fixed 520-slot buckets).  The B200 table is open addressing with linear probing (atomicCAS
insert); ``d_TK``/``d_TV`` are its slot arrays (empty = 0xFFFF.../0xFFFFFFFF), ``tableLength`` its
capacity.  The bucket layout is implementation specific in both; key -> value
content is what is kept.
"""
import logging

import numpy as np

import _native

module_logger = logging.getLogger('eulercuda.pygpuhash')

MAX_BUCKET_ITEM = 520
ULONGLONG = 8
UINTC = 4
EMPTY_KEY = np.uint64(0xFFFFFFFFFFFFFFFF)


def hash_h(key, bucketCount):
    """pygpuhash.py:28-36 (u64 wrap-around arithmetic)."""
    return int(((0x01010101 + 0x12345678 * int(key)) & 0xFFFFFFFFFFFFFFFF) % 1900813 % int(bucketCount))


def phase1_device(d_keys, d_offset, d_length, count, bucketCount):
    """pygpuhash.py:19 -- bucket histogram and per-key arrival offset."""
    return _native.default_context().compat_phase1(np.asarray(d_keys, dtype=np.uint64)[:int(d_length)], int(bucketCount))


def copy_to_bucket_device(d_keys, d_values, d_offset, d_length, d_start, bucketCount, d_bufferK, d_bufferV):
    """pygpuhash.py:77 -- scatter (key, value) to start[bucket] + offset."""
    n = int(d_length)
    return _native.default_context().compat_copy_to_bucket(
        np.asarray(d_keys, dtype=np.uint64)[:n], np.asarray(d_values, dtype=np.uint32)[:n],
        np.asarray(d_offset, dtype=np.uint32)[:n], np.asarray(d_start, dtype=np.uint32), int(bucketCount))


def bucket_sort_device(d_bufferK, d_bufferV, d_start, d_bucketSize, bucketCount, d_TK, d_TV):
    """pygpuhash.py:174 -- per-bucket ascending sort into 520-slot buckets."""
    return _native.default_context().compat_bucket_sort(
        np.asarray(d_bufferK, dtype=np.uint64), np.asarray(d_bufferV, dtype=np.uint32),
        np.asarray(d_start, dtype=np.uint32), np.asarray(d_bucketSize, dtype=np.uint32), int(bucketCount))


def create_hash_table_device(d_keys, d_values, d_length, d_TK, d_TV, tableLength, d_bucketSize, bucketCount):
    """pygpuhash.py:262 -- returns [tableLength, d_bucketSize, bucketCount, d_TK, d_TV]."""
    module_logger.info("started.")
    n = int(d_length)
    keys = np.asarray(d_keys, dtype=np.uint64)[:n]
    values = np.asarray(d_values, dtype=np.uint32)[:n]
    ctx = _native.default_context()
    TK, TV = ctx.hash_build(keys, values)
    tableLength = TK.size
    d_bucketSize = np.array([n], dtype=np.uint32)   # vestigial: one open-addressed "bucket"
    bucketCount = 1
    module_logger.info("Finished. Leaving.")
    return [tableLength, d_bucketSize, bucketCount, TK, TV]


def get_hash_value(d_keys, d_TK, d_TV, d_bucketSize=None, bucketCount=None):
    """getHashValue (pydebruijn.py:57-88) as a bulk lookup: value or 0xffffffff."""
    return _native.default_context().hash_lookup(d_TK, d_TV, np.asarray(d_keys, dtype=np.uint64))
