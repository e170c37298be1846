"""CPU test of the N > 1 host logic with a real 2-rank gloo group: exchange planning, count
exchange and the ownership rule.  The device kernels are replaced by numpy here (the oracle supplies
the encodings), so this covers the plumbing around libeuler_b200's dist entry points."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _owner_kmer(v, k, world, revcomp):
    m = min(12, k)
    mask = (1 << (2 * m)) - 1
    best = 0xffffffff
    for j in range(k - m + 1):
        w = (int(v) >> (2 * (k - m - j))) & mask
        c = min(w, revcomp(w, m))
        s = (c * 2654435761) & 0xffffffff
        s ^= s >> 15
        best = min(best, s)
    return ((best & 0xffff) * world) >> 16


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "pycuda-euler_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle
    from util import random_reads
    from eulercuda.dist import plan_exchange, torch_count_exchange, global_id_bases
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        l, k = 14, 13
        reads = random_reads(9, 200, genome_len=2000)
        buf, off = oracle.pack_reads(reads[rank::world])
        f, r, v = oracle.encode_positions(buf, off, l)
        fw, rc = f[v == 1], r[v == 1]
        canon = np.minimum(fw, rc)
        kmask = np.uint64((1 << (2 * k)) - 1)
        pre, suf = fw >> np.uint64(2), fw & kmask
        o1 = np.array([_owner_kmer(x, k, world, oracle.revcomp) for x in pre], np.int64)
        o2 = np.array([_owner_kmer(x, k, world, oracle.revcomp) for x in suf], np.int64)
        buckets = [np.concatenate([canon[o1 == d], canon[(o2 == d) & (o2 != o1)]]) for d in range(world)]
        send_counts = [int(b.size) for b in buckets]
        send_off, recv_counts = plan_exchange(send_counts, torch_count_exchange(device="cpu"))
        assert send_off.tolist() == np.concatenate([[0], np.cumsum(send_counts)[:-1]]).tolist()
        # the exchange itself (gloo: gather everything, keep our column)
        gathered = [None] * world
        dist.all_gather_object(gathered, [b.tolist() for b in buckets])
        mine = np.array(sum((gathered[src][rank] for src in range(world)), []), dtype=np.uint64)
        assert [len(gathered[src][rank]) for src in range(world)] == recv_counts
        keys, cnt = np.unique(mine, return_counts=True)
        ids = global_id_bases(10 + rank, 100 + 7 * rank, 1000 * (rank + 1), device="cpu")
        assert ids["vertices"] == sum(10 + r for r in range(world)) and ids["edges"] == sum(1000 * (r + 1) for r in range(world))
        assert ids["vertex_base"] == sum(10 + r for r in range(rank)) and ids["lmer_base"] == sum(100 + 7 * r for r in range(rank))
        assert ids["edge_base"] == sum(1000 * (r + 1) for r in range(rank))
        q.put((rank, keys.tolist(), cnt.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_partition_over_gloo():
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import oracle
    from util import random_reads
    world, port = 2, 29611
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every canonical l-mer ends up, with its full multiplicity, on the owner(s) of its end k-mers
    l, k = 14, 13
    reads = random_reads(9, 200, genome_len=2000)
    buf, off = oracle.pack_reads(reads)
    lo, hi, vals = oracle.count_mers(buf, off, l)
    canon = {}
    for key, c in zip(lo.tolist(), vals.tolist()):
        ck = min(key, oracle.revcomp(key, l))
        canon[ck] = c if key != oracle.revcomp(key, l) else c // 2
    seen = {}
    for rank, keys, cnt in results:
        for kk, c in zip(keys, cnt):
            assert canon[kk] == c
            seen.setdefault(kk, set()).add(rank)
    kmask = (1 << (2 * k)) - 1
    for ck in canon:
        p, s = ck >> 2, ck & kmask
        owners = {_owner_kmer(p, k, world, oracle.revcomp), _owner_kmer(s, k, world, oracle.revcomp)}
        assert seen[ck] == owners


def test_plan_exchange_pure():
    sys.path.insert(0, os.path.join(ROOT, "pycuda-euler_b200"))
    from eulercuda.dist import plan_exchange
    off, recv = plan_exchange([3, 0, 5], lambda s: [7, 8, 9])
    assert off.tolist() == [0, 3, 3] and recv == [7, 8, 9]


def test_assembler_driver_splits_files_at_record_boundaries():
    """partition driver: byte ranges of a FASTA / FASTQ file per rank, cut at record boundaries"""
    sys.path.insert(0, os.path.join(ROOT, "pycuda-euler_b200"))
    from assembler import split_records, detect_format
    fq = b"".join(b"@r%d\nACGT%s\n+\nIIII%s\n" % (i, b"A" * (i % 7), b"I" * (i % 7)) for i in range(101))
    for world in (1, 2, 3, 8):
        rs = split_records(fq, world, 2)
        assert rs[0][0] == 0 and rs[-1][1] == len(fq) and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        for a, b in rs:
            chunk = fq[a:b]
            assert chunk == b"" or (chunk[:1] == b"@" and chunk.count(b"\n") % 4 == 0)
        assert b"".join(fq[a:b] for a, b in rs) == fq
    fa = b"".join(b">r%d\nACGTACGT\n" % i for i in range(50))
    rs = split_records(fa, 4, 1)
    assert b"".join(fa[a:b] for a, b in rs) == fa and all(fa[a:a + 1] in (b">", b"A", b"") for a, _ in rs)
    assert split_records(b"", 3, 1) == [(0, 0)] * 3
    assert detect_format("x.fq", b"") == 2 and detect_format("x.fasta", b"") == 1 and detect_format("x.txt", b"@r") == 2


class _FakeCtx:
    """stands in for _native.Context in the transport handshakes (no GPU): allocation / IPC opens succeed or fail on demand"""

    def __init__(self, rank, fail_alloc=False, fail_open=False):
        self.rank, self.fail_alloc, self.fail_open, self.closed = rank, fail_alloc, fail_open, []

    def bkt_area_alloc(self, which, world, scap):
        if self.fail_alloc:
            raise RuntimeError("out of memory (injected)")
        return 0x1000 * (self.rank + 1) + which, bytes([self.rank, which]) * 32

    def dist_recv_alloc(self, nkeys):
        if self.fail_alloc:
            raise RuntimeError("out of memory (injected)")
        return 0x2000 * (self.rank + 1), bytes([self.rank]) * 64

    def dist_peer_open(self, handle):
        if self.fail_open:
            raise RuntimeError("no peer access (injected)")
        return 0x9000 + handle[0] * 16 + (handle[1] if handle[1] < 2 else 0)

    def dist_peer_close(self, ptr):
        self.closed.append(ptr)


def _handshake_worker(rank, world, port, q, mode):
    sys.path.insert(0, os.path.join(ROOT, "pycuda-euler_b200"))
    import torch.distributed as dist
    from eulercuda.dist import BucketExchange, _peer_exchange
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = {}
        # one rank fails (allocation or IPC open): EVERY rank must see the exchange as unusable and nobody may hang
        ctx = _FakeCtx(rank, fail_alloc=(mode == "alloc" and rank == 1), fail_open=(mode == "open" and rank == 0))
        bx = BucketExchange(ctx, rank, world, 10 + rank, 100 - rank)
        out["bucket_ok"] = bx.ok
        out["geometry"] = (bx.nb_per_rank, bx.scap)
        px = _peer_exchange(ctx, rank, world, 64 + 2 * rank, None)
        out["peer"] = px is not None
        out["peer_again"] = _peer_exchange(ctx, rank, world, 64, None) is not None   # remembered, no second handshake
        if px is not None:
            out["seg_cap"] = px.seg_cap
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,port", [("ok", 29621), ("alloc", 29622), ("open", 29623)])
def test_transport_is_agreed_collectively(mode, port):
    """a rank whose allocation / IPC open fails must not leave the others on the peer path (different
    collectives -> hang or a silently wrong table).  Both exchanges report success collectively."""
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_handshake_worker, args=(r, world, port, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        assert res[r]["bucket_ok"] == (mode == "ok")
        assert res[r]["peer"] == (mode == "ok") and res[r]["peer_again"] == res[r]["peer"]
        assert res[r]["geometry"] == (11, 100)          # the maximum of the proposals, the same on every rank
    if mode == "ok":
        assert res[0]["seg_cap"] == res[1]["seg_cap"] == 66


def test_plan_buckets_pure():
    sys.path.insert(0, os.path.join(ROOT, "pycuda-euler_b200"))
    from eulercuda.dist import plan_buckets
    nb1, sc1 = plan_buckets(138_000_000, 32, 1, 4_600_000, cap=1792)
    assert 8000 < nb1 < 10000 and 15_000_000 < sc1 < 40_000_000      # ~0.16 records per base, x 1.5
    nb8, sc8 = plan_buckets(138_000_000, 32, 8, 5_290_000, cap=1792)
    assert sc8 * 8 < sc1 * 1.01 + 8 * 4096 and nb8 >= nb1
    assert plan_buckets(0, 10, 2) == (1, 4096)
