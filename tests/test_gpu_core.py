"""GPU parity tests: every stage of the CUDA path, called through the C ABI, against the CPU
oracle on the same inputs.  Bit-exact (integer / byte work)."""
import numpy as np
import pytest

import oracle
from util import random_reads, lower_some

pytestmark = pytest.mark.gpu


def _inputs(g200_reads):
    return {
        "g200": g200_reads,
        "rand_ragged_N": random_reads(1, 400),
        "rand_lower": lower_some(random_reads(2, 300), 7),
        "polyA": ["A" * 80, "T" * 80, "ACGT" * 20, "A" * 31, "AC" * 40] * 3,
        "tiny": ["ACG", "", "A", "ACGTACGTACGT"],
        "one_long": ["".join("ACGT"[(i * 7 + i // 3) % 4] for i in range(5000))],
    }


@pytest.mark.parametrize("name", ["g200", "rand_ragged_N", "rand_lower", "polyA", "tiny", "one_long"])
@pytest.mark.parametrize("l", [1, 2, 10, 17, 31, 32])
def test_encode_positions(ctx, g200_reads, name, l):
    reads = _inputs(g200_reads)[name]
    buf, off = oracle.pack_reads(reads)
    f0, r0, v0 = oracle.encode_positions(buf, off, l)
    f1, r1, v1 = ctx.encode_lmers(buf, off, l)
    assert np.array_equal(v0, v1)
    assert np.array_equal(f0, f1)
    assert np.array_equal(r0, r1)


def test_encode_known_answers(ctx):
    # SURVEY §8c kernel-spec known answers
    buf, off = oracle.pack_reads(["TGGGATAATATGGTACGATC", "T" * 32, "A" * 32, "ACGT"])
    f, r, v = ctx.encode_lmers(buf, off, 10)
    assert f[0] == 959244 and r[0] == 848724
    f32, r32, v32 = ctx.encode_lmers(buf, off, 32)
    assert f32[20] == 2 ** 64 - 1 and r32[20] == 0 and v32[20] == 1
    assert f32[52] == 0 and r32[52] == 2 ** 64 - 1 and v32[52] == 1
    f4, _, _ = ctx.encode_lmers(buf, off, 4)
    assert f4[84] == 27
    pk, sk = ctx.compute_kmers(np.array([959244], np.uint64), (1 << 18) - 1)
    assert pk[0] == 239811 and sk[0] == 172812


@pytest.mark.parametrize("name", ["g200", "rand_ragged_N", "rand_lower", "polyA", "tiny"])
@pytest.mark.parametrize("length,limit", [(1, 0), (9, 0), (9, 1), (10, 0), (21, 0), (31, 1), (32, 0)])
def test_count_mers(ctx, g200_reads, name, length, limit):
    reads = _inputs(g200_reads)[name]
    buf, off = oracle.pack_reads(reads)
    lo, hi, vals = oracle.count_mers(buf, off, length)
    keep = vals > limit
    k1, v1 = ctx.count_mers(buf, off, length, limit)
    assert np.array_equal(lo[keep], k1)
    assert np.array_equal(vals[keep], v1)


def test_scan(ctx):
    rng = np.random.default_rng(0)
    for n in (0, 1, 31, 4096, 4097, 100000, 1 << 20):
        a = rng.integers(0, 50, n, dtype=np.uint32)
        out = ctx.exclusive_scan(a)
        ref = np.concatenate([[0], np.cumsum(a, dtype=np.uint64)[:-1]]).astype(np.uint32) if n else a
        assert np.array_equal(out, ref), n


def test_hash_build_lookup(ctx):
    rng = np.random.default_rng(1)
    keys = np.unique(rng.integers(0, 1 << 62, 50000, dtype=np.uint64))
    vals = np.arange(keys.size, dtype=np.uint32)
    TK, TV = ctx.hash_build(keys, vals)
    assert TK.size == ctx.hash_capacity(keys.size)
    q = np.concatenate([keys[::3], rng.integers(0, 1 << 62, 1000, dtype=np.uint64)])
    out = ctx.hash_lookup(TK, TV, q)
    kd = {int(k): int(v) for k, v in zip(keys, vals)}
    exp = np.array([kd.get(int(x), 0xFFFFFFFF) for x in q], np.uint32)
    assert np.array_equal(out, exp)
    assert (TK != np.uint64(0xFFFFFFFFFFFFFFFF)).sum() == keys.size


GRAPH_CASES = [("g200", 10), ("g200", 18), ("rand_ragged_N", 16), ("rand_ragged_N", 32), ("rand_lower", 21),
               ("polyA", 12), ("polyA", 32), ("tiny", 4), ("one_long", 32), ("one_long", 8)]


@pytest.mark.parametrize("name,l", GRAPH_CASES)
def test_pipeline_canonical_ids_bit_exact(ctx, g200_reads, name, l):
    import _native as N
    reads = _inputs(g200_reads)[name]
    buf, off = oracle.pack_reads(reads)
    g = oracle.graph_build(buf, off, l, expand=True)
    st = ctx.run_host(buf, off, l, N.RUN_EXPAND_EDGES | N.RUN_CANONICAL_IDS)
    assert (st.distinct_lmers, st.distinct_kmers, st.edge_count) == (g.nl, g.nv, g.ne)
    assert st.n_lmer_windows * 2 == g.ne
    assert np.array_equal(ctx.download(N.ART_LMER_KEYS), g.lk_lo)
    assert np.array_equal(ctx.download(N.ART_LMER_VALUES), g.lvals)
    assert np.array_equal(ctx.download(N.ART_LMER_OFFSETS), g.loffs.astype(np.uint32))
    assert np.array_equal(ctx.download(N.ART_KMER_KEYS), g.vk_lo)
    assert np.array_equal(ctx.download(N.ART_LCOUNT), g.lcount)
    assert np.array_equal(ctx.download(N.ART_ECOUNT), g.ecount)
    assert np.array_equal(ctx.download(N.ART_LSTART), g.lstart.astype(np.uint32))
    assert np.array_equal(ctx.download(N.ART_ESTART), g.estart.astype(np.uint32))
    assert np.array_equal(ctx.download(N.ART_EDGE_V1), g.ev1)
    assert np.array_equal(ctx.download(N.ART_EDGE_V2), g.ev2)
    assert np.array_equal(ctx.download(N.ART_EV), g.ev)
    assert np.array_equal(ctx.download(N.ART_EE), g.ee)
    assert np.array_equal(ctx.download(N.ART_LEV), g.lev)
    assert np.array_equal(ctx.download(N.ART_ENT), g.ent)


@pytest.mark.parametrize("name,l", GRAPH_CASES)
def test_pipeline_slot_order_label_free(ctx, g200_reads, name, l):
    """Fast path (ids in table-slot order): same graph up to the id bijection."""
    import _native as N
    reads = _inputs(g200_reads)[name]
    buf, off = oracle.pack_reads(reads)
    g = oracle.graph_build(buf, off, l, expand=False)
    st = ctx.run_host(buf, off, l, 0)
    assert (st.distinct_lmers, st.distinct_kmers, st.edge_count) == (g.nl, g.nv, g.ne)
    lk, lv = ctx.download(N.ART_LMER_KEYS), ctx.download(N.ART_LMER_VALUES)
    o = np.argsort(lk, kind="stable")
    assert np.array_equal(lk[o], g.lk_lo) and np.array_equal(lv[o], g.lvals)
    vk = ctx.download(N.ART_KMER_KEYS)
    ov = np.argsort(vk, kind="stable")
    assert np.array_equal(vk[ov], g.vk_lo)
    rank = np.empty(vk.size, np.int64)
    rank[ov] = np.arange(vk.size)           # my id -> oracle id
    lc = ctx.download(N.ART_LCOUNT).reshape(-1, 4)
    ec = ctx.download(N.ART_ECOUNT).reshape(-1, 4)
    assert np.array_equal(lc[ov], g.lcount.reshape(-1, 4))
    assert np.array_equal(ec[ov], g.ecount.reshape(-1, 4))
    v1, v2 = ctx.download(N.ART_EDGE_V1), ctx.download(N.ART_EDGE_V2)
    assert np.array_equal(rank[v1[o]], g.ev1) and np.array_equal(rank[v2[o]], g.ev2)
    ev = ctx.download(N.ART_EV)
    assert np.array_equal(ev["vid"], vk)
    assert np.array_equal(ev["lcount"], lc.sum(1)) and np.array_equal(ev["ecount"], ec.sum(1))
    ls, es = ctx.download(N.ART_LSTART), ctx.download(N.ART_ESTART)
    assert np.array_equal(ls, np.concatenate([[0], np.cumsum(lc.ravel())[:-1]]).astype(np.uint32))
    assert np.array_equal(es, np.concatenate([[0], np.cumsum(ec.ravel())[:-1]]).astype(np.uint32))
    assert np.array_equal(ev["lp"], ls[::4]) and np.array_equal(ev["ep"], es[::4])
    lo = ctx.download(N.ART_LMER_OFFSETS)
    assert np.array_equal(lo, np.concatenate([[0], np.cumsum(lv)[:-1]]).astype(np.uint32))


@pytest.mark.parametrize("name,l", GRAPH_CASES)
def test_tour_stagewise(ctx, g200_reads, name, l):
    """T1..T12 + components + spanning forest + swipe + emission, each stage against the oracle."""
    reads = _inputs(g200_reads)[name]
    buf, off = oracle.pack_reads(reads)
    g = oracle.graph_build(buf, off, l, expand=True)
    if g.ne == 0:
        pytest.skip("empty graph")
    ee0 = oracle.assign_successor(g.ev, g.lev, g.ent, g.ee)
    ee1 = ctx.assign_successor(g.ev, g.lev, g.ent, g.ee)
    assert np.array_equal(ee0, ee1)
    v0 = oracle.successor_graph(ee0)
    v1 = ctx.successor_graph(ee0)
    assert np.array_equal(v0, v1)
    D0 = oracle.components(v0)
    D1 = ctx.find_components(v0)
    assert np.array_equal(D0, D1)
    C0, off0, cv0, n0 = oracle.circuit_vertices(D0)
    C1, off1, cv1, n1 = ctx.circuit_vertices(D0)
    assert n0 == n1 and np.array_equal(C0, C1) and np.array_equal(off0, off1) and np.array_equal(cv0, cv1)
    cg0 = oracle.circuit_edges(g.ev, g.ent, D0, off0)
    cg1 = ctx.circuit_edges(g.ev, g.ent, D0, off0)
    assert np.array_equal(cg0, cg1)
    ee_s0 = ee0
    if len(cg0):
        t0 = oracle.spanning_forest(cg0, n0)
        t1 = ctx.spanning_forest(cg0, n0)
        assert np.array_equal(t0, t1)
        m0 = oracle.mark_spanning(cg0, t0, len(ee0))
        m1 = ctx.mark_spanning(cg0, t0, len(ee0))
        assert np.array_equal(m0, m1)
        ee_s0 = oracle.swipe(g.ev, g.ent, ee0, m0)
        ee_s1 = ctx.swipe(g.ev, g.ent, ee0, m0)
        assert np.array_equal(ee_s0, ee_s1)
    assert np.array_equal(oracle.contig_starts(ee_s0), ctx.contig_starts(ee_s0))
    c0 = oracle.walk_contigs(g.vk_lo, g.vk_hi, ee_s0, l)
    c1 = ctx.emit_contigs(g.ev, ee_s0, l)
    assert c0 == c1


@pytest.mark.parametrize("name,l", GRAPH_CASES)
def test_pipeline_contigs(ctx, g200_reads, name, l):
    import _native as N
    reads = _inputs(g200_reads)[name]
    buf, off = oracle.pack_reads(reads)
    c0, g = oracle.euler_contigs(buf, off, l)
    ctx.run_host(buf, off, l, N.RUN_EXPAND_EDGES | N.RUN_CANONICAL_IDS)
    c1 = ctx.pipeline_contigs()
    assert c0 == c1
    assert ctx.pipeline_contigs() == c1   # idempotent


def test_synth_reads_match_oracle(ctx):
    import torch
    G, L, n = 100000, 100, 5000
    for err in (0, 10000):
        d = torch.empty(n * L, dtype=torch.uint8, device="cuda")
        ctx.synth_reads_dev(d.data_ptr(), G, L, err, 17, n)
        ctx.sync()
        ref = oracle.synth_reads(G, L, err_ppm=err, first=17, count=n)
        assert np.array_equal(d.cpu().numpy(), ref)


def test_pipeline_device_resident_synthetic(ctx):
    """Reads generated on device, pipeline on device pointers, vs oracle on the same reads."""
    import torch
    import _native as N
    G, L, cov, l = 50000, 100, 8, 32
    n = G * cov // L
    d = torch.empty(n * L, dtype=torch.uint8, device="cuda")
    ctx.synth_reads_dev(d.data_ptr(), G, L, 5000, 0, n)
    offs = torch.arange(n + 1, dtype=torch.int64, device="cuda") * L
    ctx.sync()
    torch.cuda.synchronize()
    st = ctx.run_dev(d.data_ptr(), offs.data_ptr(), n, n * L, l, N.RUN_CANONICAL_IDS, 0)
    buf = d.cpu().numpy()
    g = oracle.graph_build(buf, oracle.fixed_offsets(n, L), l, expand=False)
    assert (st.distinct_lmers, st.distinct_kmers, st.edge_count) == (g.nl, g.nv, g.ne)
    assert st.n_kmer_windows == n * (L - l + 2)
    assert np.array_equal(ctx.download(N.ART_LMER_KEYS), g.lk_lo)
    assert np.array_equal(ctx.download(N.ART_LMER_VALUES), g.lvals)
    assert np.array_equal(ctx.download(N.ART_EV), g.ev)
    # second run with the learned capacity, then with an explicit hint: same result
    for hint in (0, G):
        st2 = ctx.run_dev(d.data_ptr(), offs.data_ptr(), n, n * L, l, N.RUN_CANONICAL_IDS, hint)
        assert st2.distinct_lmers == g.nl
        assert np.array_equal(ctx.download(N.ART_EV), g.ev)
    # deliberately tiny hint: the table must regrow and still be exact
    st3 = ctx.run_dev(d.data_ptr(), offs.data_ptr(), n, n * L, l, N.RUN_CANONICAL_IDS, 64)
    assert st3.retries > 0 and st3.distinct_lmers == g.nl
    assert np.array_equal(ctx.download(N.ART_LMER_VALUES), g.lvals)


@pytest.mark.parametrize("knob", ["EULER_B200_GLOBAL_TABLE", "EULER_B200_MINHASH", "EULER_B200_PACKED", "EULER_B200_COHASH", "EULER_B200_MERGED"])
def test_opt_in_table_variants_give_the_same_graph(knob):
    """The round-1 global-table path (EULER_B200_BUCKETED=0, also the fallback of the bucketed path) and its variants:
    EULER_B200_MINHASH=1 (minimizer-ordered homes; rolling-minimum kernels for l = 32 / 22, the
    brute-force one otherwise), EULER_B200_PACKED=1 (packed quotient count table, count-wrap side
    table) and EULER_B200_COHASH=1 (l-mer table hashed by the canonical prefix k-mer) must not change
    any artefact.  The knobs are read once per process, so this runs in a
    child process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)
import oracle, _native as N
from util import random_reads
ctx = N.Context(0)
reads = random_reads(4, 1500, genome_len=20000) + ["A" * 90, "ACGT" * 30] + ["C" * 100] * 40   # poly-C: wraps the packed count field
buf, off = oracle.pack_reads(reads)
for l in (32, 22, 27, 12):
    g = oracle.graph_build(buf, off, l, expand=True)
    st = ctx.run_host(buf, off, l, N.RUN_EXPAND_EDGES | N.RUN_CANONICAL_IDS)
    assert (st.distinct_lmers, st.distinct_kmers, st.edge_count) == (g.nl, g.nv, g.ne), l
    assert np.array_equal(ctx.download(N.ART_LMER_KEYS), g.lk_lo) and np.array_equal(ctx.download(N.ART_LMER_VALUES), g.lvals)
    assert np.array_equal(ctx.download(N.ART_EV), g.ev) and np.array_equal(ctx.download(N.ART_EE), g.ee)
    st = ctx.run_host(buf, off, l, 0)
    lk = ctx.download(N.ART_LMER_KEYS); o = np.argsort(lk)
    assert np.array_equal(lk[o], g.lk_lo) and np.array_equal(ctx.download(N.ART_LMER_VALUES)[o], g.lvals)
print("ok")
''' % (root, os.path.join(root, "pycuda-euler_b200"), os.path.join(root, "tests"))
    env = dict(os.environ, **{knob: "1", "EULER_B200_BUCKETED": "0"})
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_second_pass_rebuilds_the_buckets_that_overflow():
    """Small tables at a high load make some buckets overflow the first pass of the per-bucket build; the second pass
    rebuilds them with the largest tables and appends their artefacts: same graph, no repartition."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)
import oracle, _native as N
from util import random_reads
ctx = N.Context(0)
reads = random_reads(4, 3000, genome_len=40000) + ["A" * 90, "ACGT" * 30]
buf, off = oracle.pack_reads(reads)
redone = 0
for l in (32, 22, 27, 13):
    g = oracle.graph_build(buf, off, l, expand=True)
    st = ctx.run_host(buf, off, l, N.RUN_EXPAND_EDGES | N.RUN_CANONICAL_IDS)
    assert st.path == 1
    assert (st.distinct_lmers, st.distinct_kmers, st.edge_count) == (g.nl, g.nv, g.ne), l
    assert np.array_equal(ctx.download(N.ART_LMER_KEYS), g.lk_lo) and np.array_equal(ctx.download(N.ART_LMER_VALUES), g.lvals)
    assert np.array_equal(ctx.download(N.ART_EV), g.ev) and np.array_equal(ctx.download(N.ART_EE), g.ee)
    redone += st.redo_buckets
    print(l, st.n_buckets, st.redo_buckets, st.retries)
assert redone > 0
print("ok")
''' % (root, os.path.join(root, "pycuda-euler_b200"), os.path.join(root, "tests"))
    env = dict(os.environ, EULER_B200_BUCKETED="2", EULER_B200_BKT_CAP="256", EULER_B200_BKT_LOAD="0.6")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout[-500:] + out.stderr[-2000:]
