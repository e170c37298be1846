"""GPU parity for 128-bit keys: l in 33..64 (k up to 63, BASELINE.json configs[4]).  The reference
stops at l = 32 (KEY_T is 64-bit, pyencode.py:22-33); the oracle's 128-bit instantiation, itself
checked against referenceAssembler.build at K = 64 (tests/test_cpu_oracle.py::test_wide_keys_k63),
is the checker.  Bit-exact."""
from collections import Counter

import numpy as np
import pytest

import oracle
from util import random_reads, lower_some

pytestmark = pytest.mark.gpu


def _inputs():
    return {
        "rand": random_reads(11, 300, genome_len=3000, lens=(64, 80, 100, 100, 150), n_frac=0.1),
        "lower": lower_some(random_reads(12, 200, genome_len=2500, lens=(70, 100, 150), n_frac=0.05), 3),
        "repeat": ["ACGT" * 40, "A" * 150, "T" * 150, "AC" * 75, "ACGTTGCA" * 20, "TGCAACGT" * 12] * 2,
        "short": ["ACGT" * 8, "ACGTACGTAC", "", "A" * 33, "ACGGTCA" * 9 + "A"],
        "one_long": ["".join("ACGT"[(i * 7 + i // 3 + i // 11) % 4] for i in range(4000))],
    }


WIDE_CASES = [("rand", 33), ("rand", 40), ("rand", 64), ("lower", 48), ("lower", 63), ("repeat", 33), ("repeat", 64),
              ("short", 33), ("short", 34), ("one_long", 64), ("one_long", 57)]


@pytest.mark.parametrize("name,l", WIDE_CASES)
def test_wide_canonical_ids_bit_exact(ctx, name, l):
    import _native as N
    buf, off = oracle.pack_reads(_inputs()[name])
    g = oracle.graph_build(buf, off, l, expand=True)
    st = ctx.run_host(buf, off, l, N.RUN_EXPAND_EDGES | N.RUN_CANONICAL_IDS)
    assert (st.distinct_lmers, st.distinct_kmers, st.edge_count) == (g.nl, g.nv, g.ne)
    assert st.n_lmer_windows * 2 == g.ne
    assert np.array_equal(ctx.download(N.ART_LMER_KEYS), g.lk_lo)
    assert np.array_equal(ctx.download(N.ART_LMER_KEYS_HI), g.lk_hi)
    assert np.array_equal(ctx.download(N.ART_LMER_VALUES), g.lvals)
    assert np.array_equal(ctx.download(N.ART_LMER_OFFSETS), g.loffs.astype(np.uint32))
    assert np.array_equal(ctx.download(N.ART_KMER_KEYS), g.vk_lo)
    assert np.array_equal(ctx.download(N.ART_KMER_KEYS_HI), g.vk_hi)
    for which, exp in ((N.ART_LCOUNT, g.lcount), (N.ART_ECOUNT, g.ecount), (N.ART_LSTART, g.lstart.astype(np.uint32)),
                       (N.ART_ESTART, g.estart.astype(np.uint32)), (N.ART_EDGE_V1, g.ev1), (N.ART_EDGE_V2, g.ev2),
                       (N.ART_EV, g.ev), (N.ART_EE, g.ee), (N.ART_LEV, g.lev), (N.ART_ENT, g.ent)):
        assert np.array_equal(ctx.download(which), exp), which


@pytest.mark.parametrize("name,l", WIDE_CASES)
def test_wide_contigs(ctx, name, l):
    import _native as N
    buf, off = oracle.pack_reads(_inputs()[name])
    c0, g = oracle.euler_contigs(buf, off, l)
    ctx.run_host(buf, off, l, N.RUN_EXPAND_EDGES | N.RUN_CANONICAL_IDS)
    c1 = ctx.pipeline_contigs()
    assert c0 == c1


@pytest.mark.parametrize("name,l", [("rand", 64), ("lower", 40), ("repeat", 64)])
def test_wide_slot_order_spells_every_edge(ctx, name, l):
    """Slot-order ids (no sort): same key/multiplicity multiset, and the contigs spell every l-mer
    exactly `multiplicity` times -- the size-independent Euler property."""
    import _native as N
    buf, off = oracle.pack_reads(_inputs()[name])
    g = oracle.graph_build(buf, off, l, expand=False)
    st = ctx.run_host(buf, off, l, N.RUN_EXPAND_EDGES)
    assert (st.distinct_lmers, st.distinct_kmers, st.edge_count) == (g.nl, g.nv, g.ne)
    lo, hi, lv = ctx.download(N.ART_LMER_KEYS), ctx.download(N.ART_LMER_KEYS_HI), ctx.download(N.ART_LMER_VALUES)
    got = {oracle.decode_key(a, b, l): int(v) for a, b, v in zip(lo, hi, lv)}
    exp = {oracle.decode_key(a, b, l): int(v) for a, b, v in zip(g.lk_lo, g.lk_hi, g.lvals)}
    assert got == exp
    cnt = Counter()
    for c in ctx.pipeline_contigs():
        for i in range(len(c) - l + 1):
            cnt[c[i:i + l]] += 1
    assert cnt == exp


def test_wide_rejects_beyond_64(ctx):
    import _native as N
    buf, off = oracle.pack_reads(["ACGT" * 30])
    with pytest.raises(N.EulerError):
        ctx.run_host(buf, off, 65, 0)


def test_wide_k63_synthetic_full_reads(ctx):
    """configs[4] shape at reduced size: 150 bp reads, k = 63 (l = 64), 1 % errors, device-resident input."""
    import torch
    import _native as N
    G, L, cov = 200_000, 150, 20
    nreads = G * cov // L
    reads = oracle.synth_reads(G, L, err_ppm=10000, first=0, count=nreads)
    off = oracle.fixed_offsets(nreads, L)
    d_buf = torch.from_numpy(reads.copy()).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    st = ctx.run_dev(d_buf.data_ptr(), d_off.data_ptr(), nreads, reads.size, 64, 0)
    lo, hi, vals = oracle.count_mers(reads, off, 64)
    assert st.distinct_lmers == lo.size
    assert st.n_lmer_windows == nreads * (L - 64 + 1)
    assert st.n_kmer_windows == nreads * (L - 63 + 1)
    glo, ghi, gv = ctx.download(N.ART_LMER_KEYS), ctx.download(N.ART_LMER_KEYS_HI), ctx.download(N.ART_LMER_VALUES)
    o = np.lexsort((glo, ghi))
    assert np.array_equal(glo[o], lo) and np.array_equal(ghi[o], hi) and np.array_equal(gv[o], vals)


def test_wide_module_surface(ctx):
    """readLmersKmersCuda / assemble2(mode='euler') with lmerLength = 64 (fixed-length reads, reference layout)."""
    import eulercuda.eulercuda as ec
    reads = [r for r in random_reads(21, 120, genome_len=2000, lens=(100,), n_frac=0.0) if len(r) == 100]
    readBuffer = "".join(reads).encode("ascii")
    buf, off = oracle.pack_reads(reads)
    g = oracle.graph_build(buf, off, 64, expand=True)
    vals = ec.readLmersKmersCuda(readBuffer, 100, len(readBuffer), 64, [], [], 0, [], [], 0, len(reads))
    assert vals[0] == g.nl and vals[1] == g.nv
    assert [int(x) for x in vals[2]] == [(int(h) << 64) | int(lo) for lo, h in zip(g.lk_lo, g.lk_hi)]
    assert [int(x) for x in vals[4]] == [(int(h) << 64) | int(lo) for lo, h in zip(g.vk_lo, g.vk_hi)]
    assert np.array_equal(vals[3], g.lvals)
    ref, _ = oracle.euler_contigs(buf, off, 64)
    assert ec.assemble2(64, buffer=reads, mode="euler") == ref
