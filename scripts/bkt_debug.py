"""Development check of the bucketed path on a GPU box: small parity cases against the oracle (with diagnostics on
mismatch), then timings of the bench workload.  Not part of the product or the tests."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pycuda-euler_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import oracle
import _native as N
from util import random_reads

ctx = N.Context(0)


def diff(name, a, b):
    if a.shape != b.shape:
        print("   %s: shape %s vs %s" % (name, a.shape, b.shape))
        return False
    if a.dtype.names:
        ok = True
        for f in a.dtype.names:
            ok &= diff(name + "." + f, a[f], b[f])
        return ok
    bad = np.nonzero(a != b)[0]
    if len(bad):
        print("   %s: %d / %d differ, first at %d: got %s want %s" % (name, len(bad), a.size, bad[0], a[bad[:4]], b[bad[:4]]))
        return False
    return True


def check(reads, l, tag):
    buf, off = oracle.pack_reads(reads)
    g = oracle.graph_build(buf, off, l, expand=True)
    try:
        st = ctx.run_host(buf, off, l, N.RUN_EXPAND_EDGES | N.RUN_CANONICAL_IDS)
    except Exception as e:
        print("FAIL %s l=%d: %s" % (tag, l, e))
        return False
    ok = (st.distinct_lmers, st.distinct_kmers, st.edge_count) == (g.nl, g.nv, g.ne)
    print("%s l=%d path=%d nb=%d retries=%d U=%d/%d V=%d/%d E=%d/%d Nl=%d" % (tag, l, st.path, st.n_buckets, st.retries, st.distinct_lmers, g.nl,
                                                                          st.distinct_kmers, g.nv, st.edge_count, g.ne, st.n_lmer_windows))
    if ok:
        for name, art, want in (("lkeys", N.ART_LMER_KEYS, g.lk_lo), ("lvals", N.ART_LMER_VALUES, g.lvals),
                                ("loffs", N.ART_LMER_OFFSETS, g.loffs.astype(np.uint32)), ("vkeys", N.ART_KMER_KEYS, g.vk_lo),
                                ("lcount", N.ART_LCOUNT, g.lcount), ("ecount", N.ART_ECOUNT, g.ecount),
                                ("lstart", N.ART_LSTART, g.lstart.astype(np.uint32)), ("estart", N.ART_ESTART, g.estart.astype(np.uint32)),
                                ("ev1", N.ART_EDGE_V1, g.ev1), ("ev2", N.ART_EDGE_V2, g.ev2), ("ev", N.ART_EV, g.ev), ("ee", N.ART_EE, g.ee),
                                ("lev", N.ART_LEV, g.lev), ("ent", N.ART_ENT, g.ent)):
            ok &= diff(name, ctx.download(art), want)
    print("   ->", "ok" if ok else "MISMATCH")
    return ok


allok = True
reads = random_reads(4, 1500, genome_len=20000) + ["A" * 90, "ACGT" * 30] + ["C" * 100] * 40
for l in (32, 22, 27, 12, 10, 5, 2, 31, 16):
    allok &= check(reads, l, "mixed")
allok &= check(random_reads(7, 20000, genome_len=300000, lens=(100,)), 32, "20k")
allok &= check(["ACGTACGTTGCAACGTTGCATGCAAACCGGTT" * 4], 32, "one")
allok &= check([], 32, "empty")
allok &= check(["ACGT"], 32, "short")
os.environ["EULER_B200_BKT_NB"] = "1"
allok &= check(reads[:300], 32, "nb1")
os.environ["EULER_B200_BKT_NB"] = "37"
allok &= check(reads, 32, "nb37")
del os.environ["EULER_B200_BKT_NB"]
print("PARITY", "OK" if allok else "FAILED")

# ---- timings on the bench workload
G, L, cov, l = 4_600_000, 100, 30, 32
R = G * cov // L
d_reads = torch.empty(R * L, dtype=torch.uint8, device="cuda")
ctx.synth_reads_dev(d_reads.data_ptr(), G, L, 0, 0, R)
d_off = torch.arange(R + 1, dtype=torch.int64, device="cuda") * L
ctx.sync()
for mode in ("1", "0"):
    os.environ["EULER_B200_BUCKETED"] = mode
    for logcap, load in (((1536, 0.3), (2048, 0.3)) if mode == "1" else ((0, 0),)):
        if logcap:
            os.environ["EULER_B200_BKT_CAP"] = str(logcap)
            os.environ["EULER_B200_BKT_LOAD"] = str(load)
        for hint in (G,):
            ts = []
            for it in range(6):
                t0 = time.perf_counter()
                st = ctx.run_dev(d_reads.data_ptr(), d_off.data_ptr(), R, R * L, l, 0, hint)
                ts.append(1e3 * (time.perf_counter() - t0))
            print("bucketed=%s cap=%d load=%.2f hint=%d: path=%d nb=%d retries=%d maxrec=%d U=%d V=%d E=%d | ms total %.3f part %.3f build %.3f count %.3f graph %.3f | wall %s"
                  % (mode, logcap, load, hint, st.path, st.n_buckets, st.retries, st.bucket_records, st.distinct_lmers, st.distinct_kmers, st.edge_count,
                     st.ms_total, st.ms_count_kernel, st.ms_build_kernel, st.ms_count, st.ms_graph, " ".join("%.2f" % t for t in ts)))
