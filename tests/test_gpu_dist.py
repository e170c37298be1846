"""GPU test of the k-mer-space partition: `world` ranks emulated in one process on one
GPU (partition -> exchange -> per-rank build), checked against the single-graph oracle."""
import numpy as np
import pytest

import oracle
from util import random_reads

pytestmark = pytest.mark.gpu
NO_ID = 0xFFFFFFFF


def owner_kmer(v, k, world):
    """numpy-free restatement of the device rule: owner = low 16 bits of the min scrambled canonical m-mer score, scaled"""
    m = min(12, k)
    mask = (1 << (2 * m)) - 1
    best = 0xffffffff
    for j in range(k - m + 1):
        w = (int(v) >> (2 * (k - m - j))) & mask
        c = min(w, oracle.revcomp(w, m))
        s = (c * 2654435761) & 0xffffffff
        s ^= s >> 15
        best = min(best, s)
    return ((best & 0xffff) * world) >> 16


def owner_np(kmers, world, k):
    return np.array([owner_kmer(v, k, world) for v in kmers], dtype=np.int64)


def canon_np(x, k):
    return np.array([min(int(v), oracle.revcomp(int(v), k)) for v in x], dtype=np.uint64)


@pytest.mark.parametrize("impl", ["records", "keys"])   # records: the bucketed form (csrc/bucket.cuh); keys: the round-1 key exchange
@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("l", [12, 22, 32])   # 22 and 32 use the rolling-minimizer kernels, 12 the generic one
def test_partitioned_graph_equals_oracle(ctx, world, l, impl):
    from eulercuda.dist import emulate_partitioned, emulate_partitioned_bucketed
    reads = random_reads(21, 600, genome_len=5000) + ["A" * 70, "ACGT" * 12]
    k = l - 1
    shards = []
    for r in range(world):
        shards.append(oracle.pack_reads(reads[r::world] if world < 8 else reads[r * 97 % len(reads)::world]))
    if world == 8:   # uneven shards (one rank has nothing): region counts of an idle source must read as zero
        shards = [oracle.pack_reads(reads[r::7]) for r in range(7)] + [oracle.pack_reads([])]
    if impl == "records":
        parts, windows = emulate_partitioned_bucketed(ctx, shards, l, world, nb_per_rank=5 if world != 3 else None)
    else:
        parts, windows = emulate_partitioned(ctx, shards, l, world)
    buf, off = oracle.pack_reads(reads)
    g = oracle.graph_build(buf, off, l, expand=False)
    assert sum(w[0] for w in windows) * 2 == g.ne

    # vertices: every vertex on exactly one rank, the one its canonical k-mer hashes to
    all_v = np.concatenate([p["KMER_KEYS"] for p in parts])
    assert np.array_equal(np.sort(all_v), g.vk_lo)
    okey = {int(kk): i for i, kk in enumerate(g.vk_lo)}
    for r, p in enumerate(parts):
        vk = p["KMER_KEYS"]
        if vk.size:
            assert (owner_np(vk, world, k) == r).all()
        ids = np.array([okey[int(x)] for x in vk], dtype=np.int64)
        lc, ec = p["LCOUNT"].reshape(-1, 4), p["ECOUNT"].reshape(-1, 4)
        assert np.array_equal(lc, g.lcount.reshape(-1, 4)[ids]) and np.array_equal(ec, g.ecount.reshape(-1, 4)[ids])
        ev = p["EV"]
        assert np.array_equal(ev["vid"], vk)
        assert np.array_equal(ev["lcount"], lc.sum(1)) and np.array_equal(ev["ecount"], ec.sum(1))
        ls = np.concatenate([[0], np.cumsum(lc.ravel())[:-1]]).astype(np.uint32) if vk.size else p["LSTART"]
        es = np.concatenate([[0], np.cumsum(ec.ravel())[:-1]]).astype(np.uint32) if vk.size else p["ESTART"]
        assert np.array_equal(p["LSTART"], ls) and np.array_equal(p["ESTART"], es)
        assert np.array_equal(ev["lp"], ls[::4]) and np.array_equal(ev["ep"], es[::4])

    # edges: every both-strand l-mer homed exactly once, on the owner of its prefix
    all_l = np.concatenate([p["LMER_KEYS"] for p in parts])
    all_m = np.concatenate([p["LMER_VALUES"] for p in parts])
    o = np.argsort(all_l, kind="stable")
    assert np.array_equal(all_l[o], g.lk_lo) and np.array_equal(all_m[o], g.lvals)
    kmask = (1 << (2 * k)) - 1
    total_e = 0
    for r, p in enumerate(parts):
        lk, lv, lo = p["LMER_KEYS"], p["LMER_VALUES"], p["LMER_OFFSETS"]
        total_e += int(lv.sum())
        assert p["stats"]["edge_count"] == int(lv.sum())
        if lk.size == 0:
            continue
        assert np.array_equal(lo, np.concatenate([[0], np.cumsum(lv)[:-1]]).astype(np.uint32))
        vk = p["KMER_KEYS"]
        pre = lk >> np.uint64(2)
        suf = lk & np.uint64(kmask)
        assert np.array_equal(vk[p["EDGE_V1"]], pre)          # v1 always local
        v2 = p["EDGE_V2"]
        local = v2 != NO_ID
        assert np.array_equal(vk[v2[local]], suf[local])
        assert (owner_np(suf[~local], world, k) != r).all()
        assert (owner_np(suf[local], world, k) == r).all()
    assert total_e == g.ne


def test_partition_counts_are_exact(ctx):
    """per-destination counts from the count pass equal what the scatter pass writes"""
    import torch
    reads = random_reads(3, 400, genome_len=3000)
    buf, off = oracle.pack_reads(reads)
    d_buf = torch.from_numpy(buf).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    l, world = 20, 4
    counts = ctx.dist_count(d_buf.data_ptr(), d_off.data_ptr(), len(reads), buf.size, l, world)
    sc = counts[:world].astype(np.int64)
    f, r, v = oracle.encode_positions(buf, off, l)
    fw = f[v == 1]
    assert counts[world] == fw.size
    k = l - 1
    pre, suf = fw >> np.uint64(2), fw & np.uint64((1 << (2 * k)) - 1)
    o1, o2 = owner_np(pre, world, k), owner_np(suf, world, k)
    exp = np.bincount(o1, minlength=world) + np.bincount(o2[o2 != o1], minlength=world)
    assert np.array_equal(sc, exp)
    send_off = np.zeros(world, np.uint64)
    send_off[1:] = np.cumsum(sc[:-1])
    send = torch.full((int(sc.sum()),), -1, dtype=torch.int64, device="cuda")
    ctx.dist_scatter(d_buf.data_ptr(), d_off.data_ptr(), len(reads), buf.size, l, world, send.data_ptr(), send_off)
    ctx.sync()
    keys = send.cpu().numpy().view(np.uint64)
    canon = canon_np(fw, l)
    for d in range(world):
        seg = np.sort(keys[int(send_off[d]):int(send_off[d]) + int(sc[d])])
        want = np.sort(np.concatenate([canon[o1 == d], canon[(o2 == d) & (o2 != o1)]]))
        assert np.array_equal(seg, want)


@pytest.mark.parametrize("l,world", [(32, 8), (22, 3), (20, 4)])
def test_segment_scatter_matches_exact_scatter(ctx, l, world):
    """The single-pass scatter into fixed-capacity segments (what the peer-store exchange uses) delivers the
    same key multiset per destination as the exact two-pass scatter, and reports a segment that overflows."""
    import torch
    reads = random_reads(5, 500, genome_len=4000)
    buf, off = oracle.pack_reads(reads)
    d_buf = torch.from_numpy(buf).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    exact = ctx.dist_count(d_buf.data_ptr(), d_off.data_ptr(), len(reads), buf.size, l, world)
    sc = exact[:world].astype(np.int64)
    send_off = np.zeros(world, np.uint64)
    send_off[1:] = np.cumsum(sc[:-1])
    send = torch.zeros(max(int(sc.sum()), 1), dtype=torch.int64, device="cuda")
    ctx.dist_scatter(d_buf.data_ptr(), d_off.data_ptr(), len(reads), buf.size, l, world, send.data_ptr(), send_off)
    ctx.sync()
    ref = send.cpu().numpy().view(np.uint64)
    cap = int(sc.max()) + 7
    seg = torch.zeros(cap * world, dtype=torch.int64, device="cuda")
    counts = ctx.dist_scatter_segments(d_buf.data_ptr(), d_off.data_ptr(), len(reads), buf.size, l, world, seg.data_ptr(), cap)
    assert np.array_equal(counts, exact)
    got = seg.cpu().numpy().view(np.uint64)
    for d in range(world):
        a = np.sort(got[d * cap:d * cap + int(sc[d])])
        b = np.sort(ref[int(send_off[d]):int(send_off[d]) + int(sc[d])])
        assert np.array_equal(a, b)
    small = max(int(sc.max()) // 2, 1)
    seg2 = torch.full((small * world + 8,), -1, dtype=torch.int64, device="cuda")
    counts2 = ctx.dist_scatter_segments(d_buf.data_ptr(), d_off.data_ptr(), len(reads), buf.size, l, world, seg2.data_ptr(), small)
    assert np.array_equal(counts2, exact)                 # the cursors keep counting past the capacity ...
    assert (seg2.cpu().numpy()[small * world:] == -1).all()   # ... but nothing is written outside the segments


def _canon_part(p):
    # slot-order ids depend on the table capacity and on insertion races: compare through the keys
    o = np.argsort(p["LMER_KEYS"], kind="stable")
    ov = np.argsort(p["KMER_KEYS"], kind="stable")
    vk = p["KMER_KEYS"]
    v2 = p["EDGE_V2"][o]
    v2k = np.where(v2 == NO_ID, np.uint64(0xFFFFFFFFFFFFFFFF), vk[np.minimum(v2, vk.size - 1)])
    return (p["LMER_KEYS"][o], p["LMER_VALUES"][o], vk[p["EDGE_V1"][o]], v2k, vk[ov],
            p["LCOUNT"].reshape(-1, 4)[ov], p["ECOUNT"].reshape(-1, 4)[ov])



def test_big_edge_total_path_matches_the_fused_one(ctx):
    """Ranks whose edge total may pass 2^32 (BASELINE configs[3], 1 Gbp) scan lcount / ecount separately and
    sum the multiplicities in 64 bits; on an input without overflow both paths must give the same artefacts."""
    import os
    from eulercuda.dist import emulate_partitioned
    reads = random_reads(33, 500, genome_len=4000)
    shards = [oracle.pack_reads(reads[r::2]) for r in range(2)]
    a, _ = emulate_partitioned(ctx, shards, 32, 2)
    os.environ["EULER_B200_FORCE_BIG"] = "1"
    try:
        b, _ = emulate_partitioned(ctx, shards, 32, 2)
    finally:
        del os.environ["EULER_B200_FORCE_BIG"]
    for pa, pb in zip(a, b):
        for x, y in zip(_canon_part(pa), _canon_part(pb)):
            assert np.array_equal(x, y)
        assert pa["stats"]["edge_count"] == pb["stats"]["edge_count"] == int(pb["LMER_VALUES"].sum())
        for p in (pa, pb):   # offsets are exclusive scans in id order (modulo 2^32 on the big path)
            lc, ec = p["LCOUNT"].astype(np.uint64), p["ECOUNT"].astype(np.uint64)
            assert np.array_equal(p["LSTART"], (np.concatenate([np.zeros(1, np.uint64), np.cumsum(lc)[:-1]]) & np.uint64(0xFFFFFFFF)).astype(np.uint32))
            assert np.array_equal(p["ESTART"], (np.concatenate([np.zeros(1, np.uint64), np.cumsum(ec)[:-1]]) & np.uint64(0xFFFFFFFF)).astype(np.uint32))
            assert np.array_equal(p["EV"]["vid"], p["KMER_KEYS"])
            assert np.array_equal(p["EV"]["lp"], p["LSTART"][::4]) and np.array_equal(p["EV"]["ep"], p["ESTART"][::4])
            assert np.array_equal(p["EV"]["lcount"], p["LCOUNT"].reshape(-1, 4).sum(1))
            assert np.array_equal(p["EV"]["ecount"], p["ECOUNT"].reshape(-1, 4).sum(1))


def test_l2_blocked_count_matches_arrival_order_count(ctx):
    """Tables far larger than L2 (1 Gbp over 8 GPUs: 3 GB) regroup the received keys by table region before
    counting (EULER_B200_BLOCK_MB, default 512); forced here on a 4 MB table: same graph either way."""
    import os
    from eulercuda.dist import emulate_partitioned
    G, L = 150_000, 100
    nreads = G * 20 // L
    reads = oracle.synth_reads(G, L, err_ppm=5000, first=0, count=nreads)
    half = nreads // 2
    shards = [(reads[:half * L], oracle.fixed_offsets(half, L)), (reads[half * L:], oracle.fixed_offsets(nreads - half, L))]
    a, _ = emulate_partitioned(ctx, shards, 32, 2)
    os.environ["EULER_B200_BLOCK_MB"] = "1"
    try:
        b, _ = emulate_partitioned(ctx, shards, 32, 2)
    finally:
        del os.environ["EULER_B200_BLOCK_MB"]
    for pa, pb in zip(a, b):
        assert pb["stats"]["kernel_launches"] > pa["stats"]["kernel_launches"]   # the blocked path really ran
        for x, y in zip(_canon_part(pa), _canon_part(pb)):
            assert np.array_equal(x, y)
        assert pa["stats"]["edge_count"] == pb["stats"]["edge_count"]


def _rc(x, n):
    """reverse complement of an n-mer held in a Python int (any width)"""
    r = 0
    for _ in range(n):
        r = (r << 2) | (3 - (x & 3))
        x >>= 2
    return r


def _wide(lo, hi):
    return [(int(h) << 64) | int(x) for x, h in zip(lo, hi)]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("l", [33, 48, 64])
def test_partitioned_wide_graph_equals_oracle(ctx, world, l):
    """128-bit keys (k up to 63) through the partition: union of the per-rank graphs == the oracle's graph."""
    from eulercuda.dist import emulate_partitioned
    reads = random_reads(23, 400, genome_len=4000, lens=(64, 80, 100, 100, 150), n_frac=0.1) + ["A" * 150, "ACGT" * 30]
    k = l - 1
    shards = [oracle.pack_reads(reads[r::world]) for r in range(world)]
    parts, windows = emulate_partitioned(ctx, shards, l, world)
    buf, off = oracle.pack_reads(reads)
    g = oracle.graph_build(buf, off, l, expand=False)
    assert sum(w[0] for w in windows) * 2 == g.ne
    gv = _wide(g.vk_lo, g.vk_hi)
    gl = _wide(g.lk_lo, g.lk_hi)
    okey = {v: i for i, v in enumerate(gv)}
    kmask = (1 << (2 * k)) - 1
    seen_v, seen_l, total_e = [], {}, 0
    for r, p in enumerate(parts):
        vk = _wide(p["KMER_KEYS"], p["KMER_KEYS_HI"])
        seen_v += vk
        assert all(owner_kmer(min(v, _rc(v, k)), k, world) == r for v in vk)
        ids = np.array([okey[v] for v in vk], dtype=np.int64)
        lc, ec = p["LCOUNT"].reshape(-1, 4), p["ECOUNT"].reshape(-1, 4)
        assert np.array_equal(lc, g.lcount.reshape(-1, 4)[ids]) and np.array_equal(ec, g.ecount.reshape(-1, 4)[ids])
        ev = p["EV"]
        assert np.array_equal(ev["vid"], p["KMER_KEYS"])
        assert np.array_equal(ev["lcount"], lc.sum(1)) and np.array_equal(ev["ecount"], ec.sum(1))
        z = np.zeros(1, np.uint64)
        if len(vk):
            assert np.array_equal(p["LSTART"], np.concatenate([z, np.cumsum(lc.ravel().astype(np.uint64))[:-1]]).astype(np.uint32))
            assert np.array_equal(p["ESTART"], np.concatenate([z, np.cumsum(ec.ravel().astype(np.uint64))[:-1]]).astype(np.uint32))
            assert np.array_equal(ev["lp"], p["LSTART"][::4]) and np.array_equal(ev["ep"], p["ESTART"][::4])
        lk = _wide(p["LMER_KEYS"], p["LMER_KEYS_HI"])
        lv = p["LMER_VALUES"]
        total_e += int(lv.sum())
        assert p["stats"]["edge_count"] == int(lv.sum())
        if len(lk):
            assert np.array_equal(p["LMER_OFFSETS"], np.concatenate([z, np.cumsum(lv.astype(np.uint64))[:-1]]).astype(np.uint32))
        for x, m, v1, v2 in zip(lk, lv, p["EDGE_V1"], p["EDGE_V2"]):
            assert x not in seen_l          # every both-strand l-mer homed exactly once
            seen_l[x] = int(m)
            assert vk[v1] == x >> 2         # v1 always local
            suf = x & kmask
            own_s = owner_kmer(min(suf, _rc(suf, k)), k, world)
            if v2 == NO_ID:
                assert own_s != r
            else:
                assert own_s == r and vk[v2] == suf
    assert sorted(seen_v) == gv
    assert seen_l == {x: int(m) for x, m in zip(gl, g.lvals)}
    assert total_e == g.ne
