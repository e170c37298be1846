"""CPU oracle for the pycuda-euler hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker / reported CPU
baseline.  Nothing under ``pycuda-euler_b200/`` imports it.

Two layers:

* ``libeuler_oracle.so`` (``euler_oracle.c`` + ``oracle_impl.h``): plain-C restatement of the
  encode -> l-mer multiset -> de Bruijn graph -> Euler tour -> contig walk path.
* this module: ctypes bindings plus a pure-Python restatement of the reference CPU assembler
  (``/root/reference/src/referenceassembler/referenceAssembler.py``) for small cases.

Parity pinning: see ``tests/golden/make_golden.py`` (runs the unmodified reference here and
commits its outputs) and ``tests/test_oracle_golden.py``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libeuler_oracle.so")

EV_DTYPE = np.dtype([("vid", np.uint64), ("ep", np.uint32), ("ecount", np.uint32),
                     ("lp", np.uint32), ("lcount", np.uint32)])
EE_DTYPE = np.dtype([("eid", np.uint64), ("v1", np.uint32), ("v2", np.uint32),
                     ("s", np.uint32), ("pad", np.uint32)])
SV_DTYPE = np.dtype([("vid", np.uint32), ("n1", np.uint32), ("n2", np.uint32)])
CE_DTYPE = np.dtype([("ceid", np.uint32), ("e1", np.uint32), ("e2", np.uint32),
                     ("c1", np.uint32), ("c2", np.uint32)])


def build_library(force=False):
    """Compile the C restatement (gcc).  Building the checker is not using it."""
    src = [os.path.join(_HERE, f) for f in ("euler_oracle.c", "oracle_impl.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libeuler_oracle.so"],
                          stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_library()
        L = C.CDLL(_LIB_PATH)
        vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
        L.orc_encode_positions.argtypes = [vp, vp, u64, u32, vp, vp, vp]
        L.orc_encode_positions.restype = C.c_int
        L.orc_compute_kmers.argtypes = [vp, u64, u64, vp, vp]
        L.orc_compute_kmers.restype = None
        L.orc_hash_h.argtypes = [u64, u32]
        L.orc_hash_h.restype = u32
        L.orc_revcomp64.argtypes = [u64, u32]
        L.orc_revcomp64.restype = u64
        L.orc_count.argtypes = [vp, vp, u64, u32]
        L.orc_count.restype = vp
        L.orc_counts_n.argtypes = [vp]
        L.orc_counts_n.restype = u64
        L.orc_counts_copy.argtypes = [vp, vp, vp, vp]
        L.orc_counts_copy.restype = None
        L.orc_counts_free.argtypes = [vp]
        L.orc_counts_free.restype = None
        L.orc_graph_build.argtypes = [vp, vp, u64, u32, C.c_int]
        L.orc_graph_build.restype = vp
        L.orc_graph_counts.argtypes = [vp, vp, vp, vp, vp]
        L.orc_graph_counts.restype = None
        L.orc_graph_copy.argtypes = [vp, C.c_int, vp]
        L.orc_graph_copy.restype = C.c_int
        L.orc_graph_free.argtypes = [vp]
        L.orc_graph_free.restype = None
        L.orc_assign_successor.argtypes = [vp, vp, vp, u32, vp, u32]
        L.orc_assign_successor.restype = None
        L.orc_successor_graph.argtypes = [vp, vp, u32]
        L.orc_successor_graph.restype = None
        L.orc_components.argtypes = [vp, vp, u32]
        L.orc_components.restype = None
        L.orc_circuit_vertices.argtypes = [vp, u32, vp, vp, vp]
        L.orc_circuit_vertices.restype = u32
        L.orc_circuit_edges.argtypes = [vp, vp, u32, vp, vp, u32, vp, u64]
        L.orc_circuit_edges.restype = u64
        L.orc_spanning_forest.argtypes = [vp, u64, u32, vp]
        L.orc_spanning_forest.restype = u32
        L.orc_mark_spanning.argtypes = [vp, vp, u32, vp]
        L.orc_mark_spanning.restype = None
        L.orc_swipe.argtypes = [vp, vp, u32, vp, vp, u32]
        L.orc_swipe.restype = None
        L.orc_contig_starts.argtypes = [vp, u32, vp]
        L.orc_contig_starts.restype = None
        L.orc_walk_contigs.argtypes = [vp, vp, vp, u32, u32, vp, u64, vp]
        L.orc_walk_contigs.restype = u64
        L.orc_synth_reads.argtypes = [u64, u32, u32, u64, u64, vp]
        L.orc_synth_reads.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# --------------------------------------------------------------------------- inputs
def pack_reads(reads):
    """list[str|bytes] -> (flat uint8 buffer, uint64 offsets[R+1]) -- the encoder input contract
    (one long string, eulercuda.py:485-486) plus explicit read boundaries (B1 fix)."""
    bs = [r.encode("ascii") if isinstance(r, str) else bytes(r) for r in reads]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    if buf.size == 0:
        buf = np.zeros(1, np.uint8)[:0]
    return buf, off


def read_fasta(path):
    """non-'>' lines, stripped (eulercuda.py:439-447)."""
    out = []
    with open(path) as f:
        for line in f:
            if line[0] != ">":
                out.append(line.strip())
    return out


def synth_reads(G, L, cov=None, err_ppm=0, first=0, count=None):
    """Deterministic synthetic reads; returns uint8[count*L]."""
    if count is None:
        count = -(-int(G * cov) // L)
    out = np.empty(int(count) * int(L), dtype=np.uint8)
    lib().orc_synth_reads(int(G), int(L), int(err_ppm), int(first), int(count), _p(out))
    return out


def fixed_offsets(nreads, L):
    return (np.arange(nreads + 1, dtype=np.uint64) * np.uint64(L)).astype(np.uint64)


# --------------------------------------------------------------------------- encoder
def encode_positions(buf, off, length):
    B = int(off[-1])
    fwd = np.zeros(B, np.uint64)
    rc = np.zeros(B, np.uint64)
    valid = np.zeros(B, np.uint8)
    if B:
        r = lib().orc_encode_positions(_p(buf), _p(off), len(off) - 1, length, _p(fwd), _p(rc), _p(valid))
        assert r == 0
    return fwd, rc, valid


def compute_kmers(lmers, mask):
    lmers = np.ascontiguousarray(lmers, dtype=np.uint64)
    pk = np.zeros_like(lmers)
    sk = np.zeros_like(lmers)
    lib().orc_compute_kmers(_p(lmers), lmers.size, int(mask), _p(pk), _p(sk))
    return pk, sk


def hash_h(key, bucket_count):
    return int(lib().orc_hash_h(int(key), int(bucket_count)))


def revcomp(x, length):
    return int(lib().orc_revcomp64(int(x), int(length)))


def encode_str(s):
    v = 0
    for ch in s:
        v = (v << 2) | "ACGT".index(ch)
    return v


def decode_key(lo, hi, length):
    x = (int(hi) << 64) | int(lo)
    out = []
    for i in range(length):
        out.append("ACGT"[(x >> (2 * (length - 1 - i))) & 3])
    return "".join(out)


# --------------------------------------------------------------------------- counting
def count_mers(buf, off, length):
    """Both-strand multiset of `length`-mers: (lo u64[n], hi u64[n], counts u32[n]), ascending."""
    h = lib().orc_count(_p(buf), _p(off), len(off) - 1, length)
    if not h:
        raise ValueError("orc_count failed")
    n = lib().orc_counts_n(h)
    lo = np.zeros(n, np.uint64)
    hi = np.zeros(n, np.uint64)
    vals = np.zeros(n, np.uint32)
    lib().orc_counts_copy(h, _p(lo), _p(hi), _p(vals))
    lib().orc_counts_free(h)
    return lo, hi, vals


# --------------------------------------------------------------------------- graph
class Graph:
    """Outputs of D1-D6 with ids = rank in ascending key order."""
    _FIELDS = [("lk_lo", 0, np.uint64, "nl"), ("lk_hi", 1, np.uint64, "nl"), ("lvals", 2, np.uint32, "nl"),
               ("loffs", 3, np.uint64, "nl"), ("vk_lo", 4, np.uint64, "nv"), ("vk_hi", 5, np.uint64, "nv"),
               ("lcount", 6, np.uint32, "nv4"), ("ecount", 7, np.uint32, "nv4"),
               ("lstart", 8, np.uint64, "nv4"), ("estart", 9, np.uint64, "nv4"),
               ("ev1", 10, np.uint32, "nl"), ("ev2", 11, np.uint32, "nl"), ("ev", 12, EV_DTYPE, "nv"),
               ("ee", 13, EE_DTYPE, "ne"), ("lev", 14, np.uint32, "ne"), ("ent", 15, np.uint32, "ne")]


def graph_build(buf, off, l, expand=True):
    h = lib().orc_graph_build(_p(buf), _p(off), len(off) - 1, l, 1 if expand else 0)
    if not h:
        raise ValueError("orc_graph_build failed")
    nl, nv, ne = C.c_uint64(), C.c_uint64(), C.c_uint64()
    ex = C.c_int()
    lib().orc_graph_counts(h, C.byref(nl), C.byref(nv), C.byref(ne), C.byref(ex))
    g = Graph()
    g.l = l
    g.nl, g.nv, g.ne, g.expanded = nl.value, nv.value, ne.value, bool(ex.value)
    sizes = {"nl": g.nl, "nv": g.nv, "nv4": 4 * g.nv, "ne": g.ne}
    for name, which, dt, sz in Graph._FIELDS:
        if which >= 13 and not g.expanded:
            setattr(g, name, None)
            continue
        a = np.zeros(sizes[sz], dtype=dt)
        if a.size:
            r = lib().orc_graph_copy(h, which, _p(a))
            assert r == 0
        setattr(g, name, a)
    lib().orc_graph_free(h)
    return g


# --------------------------------------------------------------------------- tour
def assign_successor(ev, l, e, ee):
    ee = ee.copy()
    lib().orc_assign_successor(_p(ev), _p(l), _p(e), len(ev), _p(ee), len(ee))
    return ee


def successor_graph(ee):
    v = np.zeros(len(ee), SV_DTYPE)
    lib().orc_successor_graph(_p(ee), _p(v), len(ee))
    return v


def components(v):
    D = np.zeros(len(v), np.uint32)
    lib().orc_components(_p(v), _p(D), len(v))
    return D


def circuit_vertices(D):
    n = len(D)
    Cm = np.zeros(n, np.uint32)
    offset = np.zeros(n, np.uint32)
    cv = np.zeros(n, np.uint32)
    cnt = lib().orc_circuit_vertices(_p(D), n, _p(Cm), _p(offset), _p(cv))
    return Cm, offset, cv[:cnt].copy(), int(cnt)


def circuit_edges(ev, e, D, cmap):
    ecount = len(D)
    need = lib().orc_circuit_edges(_p(ev), _p(e), len(ev), _p(D), _p(cmap), ecount, None, 0)
    out = np.zeros(need, CE_DTYPE)
    if need:
        got = lib().orc_circuit_edges(_p(ev), _p(e), len(ev), _p(D), _p(cmap), ecount, _p(out), need)
        assert got == need
    return out


def spanning_forest(cg, cg_vcount):
    tree = np.zeros(max(len(cg), 1), np.uint32)
    nt = lib().orc_spanning_forest(_p(cg), len(cg), cg_vcount, _p(tree))
    return tree[:nt].copy()


def mark_spanning(cg, tree, ecount):
    mark = np.zeros(ecount, np.uint32)
    lib().orc_mark_spanning(_p(cg), _p(tree), len(tree), _p(mark))
    return mark


def swipe(ev, e, ee, mark):
    ee = ee.copy()
    lib().orc_swipe(_p(ev), _p(e), len(ev), _p(ee), _p(mark), len(ee))
    return ee


def contig_starts(ee):
    st = np.zeros(len(ee), np.uint32)
    lib().orc_contig_starts(_p(ee), len(ee), _p(st))
    return st


def walk_contigs(vk_lo, vk_hi, ee, l):
    nc = C.c_uint64()
    need = lib().orc_walk_contigs(_p(vk_lo), _p(vk_hi), _p(ee), len(ee), l, None, 0, C.byref(nc))
    out = np.zeros(max(need, 1), np.uint8)
    lib().orc_walk_contigs(_p(vk_lo), _p(vk_hi), _p(ee), len(ee), l, _p(out), need, C.byref(nc))
    txt = out[:need].tobytes().decode("ascii")
    return txt.split("\n")[:-1] if need else []


def euler_contigs(buf, off, l):
    """Whole Euler-mode pipeline on the CPU: graph -> T1..T12 -> contigs (list[str])."""
    g = graph_build(buf, off, l, expand=True)
    if g.ne == 0:
        return [], g
    ee = assign_successor(g.ev, g.lev, g.ent, g.ee)
    v = successor_graph(ee)
    D = components(v)
    Cm, offset, cv, ncirc = circuit_vertices(D)
    if ncirc > 1:
        cg = circuit_edges(g.ev, g.ent, D, offset)
        if len(cg):
            tree = spanning_forest(cg, ncirc)
            mark = mark_spanning(cg, tree, len(ee))
            ee = swipe(g.ev, g.ent, ee, mark)
    return walk_contigs(g.vk_lo, g.vk_hi, ee, l), g


# --------------------------------------------------------------------------- unitig mode
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def twin(km):
    """reverse complement (referenceAssembler.py:7-10)."""
    return "".join(_COMP.get(b, b) for b in reversed(km))


def py_build(reads, k=31, limit=1):
    """Both-strand k-mer counts, reads split at 'N', keep count > limit (referenceAssembler.py:25-42)."""
    d = {}
    for read in reads:
        for seg in read.split("N"):
            for s in (seg, twin(seg)):
                for i in range(len(s) - k + 1):
                    km = s[i:i + k]
                    d[km] = d.get(km, 0) + 1
    return {km: c for km, c in d.items() if c > limit}


def _fw(km):
    return [km[1:] + x for x in "ACGT"]


def _bw(km):
    return [x + km[:-1] for x in "ACGT"]


def _walk_forward(d, km):
    """referenceAssembler.py:59-77 (unique-successor / unique-predecessor walk with the cycle,
    mobius and hairpin stops)."""
    path = [km]
    while True:
        last = path[-1]
        nxt = [x for x in _fw(last) if x in d]
        if len(nxt) != 1:
            break
        cand = nxt[0]
        if cand == km or cand == twin(km):
            break
        if cand == twin(last):
            break
        if sum(1 for x in _bw(cand) if x in d) != 1:
            break
        path.append(cand)
    return path


def py_all_contigs(d, k):
    """Unitigs in dict iteration order (referenceAssembler.py:47-56,79-88). Returns list[str]."""
    done = set()
    out = []
    for x in d:
        if x in done:
            continue
        fwd = _walk_forward(d, x)
        bwd = _walk_forward(d, twin(x))
        if x in _fw(fwd[-1]):
            path = fwd
        else:
            path = [twin(y) for y in bwd[-1:0:-1]] + fwd
        for y in path:
            done.add(y)
            done.add(twin(y))
        out.append(path[0] + "".join(y[-1] for y in path[1:]))
    return out


def canonical_contigs(contigs):
    """Orientation-free sorted list: min(c, twin(c)) per contig."""
    return sorted(min(c, twin(c)) for c in contigs)
