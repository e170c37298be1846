"""encoder -- drop-in for pycuda-euler's ``pyencode`` (src/eulercuda/pyencode.py).

Same function names, argument order and return values; the CUDA work happens in
libeuler_b200.so (hand-written sm_100a kernels) through the ctypes shim ``_native``.  No PyCUDA
SourceModule, no CPU fallback: without the library or a GPU these functions raise.
"""
import logging

import numpy as np

import _native

module_logger = logging.getLogger('eulercuda.pyencode')


def _as_bytes(buffer):
    """The reference passes ``np.array(readBuffer).astype('S')`` (eulercuda.py:93): a 0-d bytes
    array holding all reads concatenated.  Accept that, plain bytes/str, or a uint8 array."""
    if isinstance(buffer, np.ndarray):
        if buffer.dtype.kind == 'S':
            return np.frombuffer(b"".join(buffer.ravel().tolist()), dtype=np.uint8)
        return np.ascontiguousarray(buffer, dtype=np.uint8).ravel()
    if isinstance(buffer, str):
        buffer = buffer.encode('ascii')
    return np.frombuffer(bytes(buffer), dtype=np.uint8)


def _offsets(nbytes, readLength, read_offsets=None):
    """Read boundaries.  The reference layout is fixed-length reads back to back (readLength from
    the first read, eulercuda.py:478); windows are per read."""
    if read_offsets is not None:
        return np.ascontiguousarray(read_offsets, dtype=np.uint64)
    readLength = int(readLength)
    if readLength <= 0 or readLength >= nbytes:
        return np.array([0, nbytes], dtype=np.uint64)
    off = np.arange(0, nbytes + readLength, readLength, dtype=np.uint64)
    off[-1] = nbytes
    return off[:int(np.searchsorted(off, nbytes)) + 1]


def _fit(out, template):
    """Honour a caller-supplied output array's length (the reference sizes d_lmers itself)."""
    if isinstance(template, np.ndarray) and template.size and template.size != out.size:
        res = np.zeros(template.size, dtype=out.dtype)
        n = min(out.size, template.size)
        res[:n] = out[:n]
        return res
    return out


def encode_lmer_device(buffer, readCount, d_lmers, readLength, lmerLength, read_offsets=None):
    """pyencode.py:16 -- l-mer starting at every byte of the flat read buffer (0 where no
    window starts).  Returns ``d_lmers`` (u64)."""
    module_logger.info("started encode_lmer_device.")
    buf = _as_bytes(buffer)
    off = _offsets(buf.size, readLength, read_offsets)
    fwd, _, _ = _native.default_context().encode_lmers(buf, off, lmerLength, want_rc=False, want_valid=False)
    module_logger.info("finished encode_lmer_device.")
    return _fit(fwd, d_lmers)


def compute_kmer_device(lmers, pkmers, skmers, kmerBitMask, readLength, readCount):
    """pyencode.py:103 -- prefix / suffix (l-1)-mers of every l-mer."""
    module_logger.info("started compute_kmer_device.")
    pk, sk = _native.default_context().compute_kmers(np.asarray(lmers, dtype=np.uint64), int(kmerBitMask))
    module_logger.info("leaving compute_kmer_device.")
    return pk, sk


def compute_lmer_complement_device(buffer, readCount, d_lmers, readLength, lmerLength, read_offsets=None):
    """pyencode.py:164 -- reverse-complement l-mer of the window starting at every byte
    (the reference kernel's intent; its body is defective)."""
    module_logger.info("started compute_lmer_complement_device.")
    buf = _as_bytes(buffer)
    off = _offsets(buf.size, readLength, read_offsets)
    _, rc, _ = _native.default_context().encode_lmers(buf, off, lmerLength, want_rc=True, want_valid=False)
    module_logger.info("Finished compute_lmer_complement_device.")
    return _fit(rc, d_lmers)


def valid_window_mask(buffer, readLength, lmerLength, read_offsets=None):
    """Extension: 1 where a whole l-mer window starts (disambiguates poly-A from 'no window')."""
    buf = _as_bytes(buffer)
    off = _offsets(buf.size, readLength, read_offsets)
    return _native.default_context().encode_lmers(buf, off, lmerLength, want_rc=False, want_valid=True)[2]


def getOptimalLaunchConfiguration(threadCount, threadPerBlock):
    """pyencode.py:239 -- (block_dim, grid_dim) 3-tuples of the reference's 2-D grid rule.  Kept
    for callers; the sm_100a kernels size their own persistent grids."""
    bx = int(threadPerBlock)
    gx, gy = 1, 1
    if threadCount > bx:
        gy = -(-int(threadCount) // bx)
        gx = gy // 65535 + 1
        gy = min(gy, 65535)
    return (bx, 1, 1), (gx, gy, 1)
