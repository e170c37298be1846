"""k-mer-space partition across the GPUs of one box: one process per GPU, one all-to-all.

The reference distributes by Spark ``mapPartitions(assemble2)`` over independent read partitions
with no exchange and no merge (src/cli_spark_gpu.py:37), which breaks the graph at partition
boundaries.  Here every vertex (k-mer) has one owning rank; each rank encodes its shard of the
reads, sends every canonical l-mer to the owner(s) of its prefix / suffix k-mer in ONE
``all_to_all_single`` over NCCL, and builds its part of the graph from what it receives (see
csrc/dist.cu).  torch.distributed is plumbing only; all compute is in libeuler_b200.so.
"""
import os

import numpy as np


def plan_exchange(send_counts, all_to_all_counts):
    """Host logic of the exchange: offsets of the send buffer and the receive split sizes.

    send_counts: int sequence [world] (keys this rank sends to each rank);
    all_to_all_counts: callable taking that list and returning what every rank sends to us.
    Returns (send_off uint64[world], recv_counts list[int])."""
    send_counts = [int(x) for x in send_counts]
    send_off = np.zeros(len(send_counts), dtype=np.uint64)
    if len(send_counts) > 1:
        send_off[1:] = np.cumsum(send_counts[:-1], dtype=np.uint64)
    recv_counts = [int(x) for x in all_to_all_counts(send_counts)]
    return send_off, recv_counts


def torch_count_exchange(group=None, device="cuda"):
    """all_to_all of the per-destination counts with torch.distributed (NCCL on GPU, gloo on CPU)."""
    import torch
    import torch.distributed as dist

    def fn(send_counts):
        world = dist.get_world_size(group)
        t_in = torch.tensor(send_counts, dtype=torch.int64, device=device)
        t_out = torch.empty(world, dtype=torch.int64, device=device)
        if dist.get_backend(group) == "gloo":
            # gloo has no all_to_all_single on every build: gather + pick our column
            rows = [torch.empty(world, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(rows, t_in.cpu(), group=group)
            me = dist.get_rank(group)
            return [int(r[me]) for r in rows]
        dist.all_to_all_single(t_out, t_in, group=group)
        return t_out.tolist()
    return fn


def global_id_bases(n_vertices, n_lmers, n_edges, group=None, device="cuda"):
    """Global ids = local id + exclusive scan of the per-rank counts: one all_gather of three
    integers.  Returns {"vertex_base", "lmer_base", "edge_base", "vertices", "lmers", "edges"} with the
    bases of this rank and the totals over all ranks."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = torch.tensor([int(n_vertices), int(n_lmers), int(n_edges)], dtype=torch.int64, device=device)
    rows = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(rows, mine, group=group)
    table = torch.stack(rows).cpu().numpy()
    base = table[:rank].sum(axis=0) if rank else np.zeros(3, np.int64)
    tot = table.sum(axis=0)
    return {"vertex_base": int(base[0]), "lmer_base": int(base[1]), "edge_base": int(base[2]),
            "vertices": int(tot[0]), "lmers": int(tot[1]), "edges": int(tot[2])}


_BUFFERS = {}


def _buffer(name, n, device):
    """grow-only cached int64 CUDA buffers (no allocation in steady state)"""
    import torch
    t = _BUFFERS.get((name, str(device)))
    if t is None or t.numel() < n:
        t = torch.empty(int(n * 1.1) + 1024, dtype=torch.int64, device=device)
        _BUFFERS[(name, str(device))] = t
    return t


class PeerExchange:
    """Peer-mapped receive buffers: rank r's scatter kernel stores the keys for rank d straight into
    region r of d's buffer over NVLink (CUDA IPC), so the all-to-all is fused into the partition pass."""

    def __init__(self, ctx, rank, world, seg_cap, group=None):
        """seg_cap is in 8-byte words (an l-mer key is one word for l <= 32, two above)"""
        import torch.distributed as dist
        # every rank lays its receive buffer out as `world` regions of seg_cap words and writes into region
        # `rank` of its peers: the stride must be the same everywhere, so the ranks agree on the largest
        # proposal (their shards, hence their estimates, differ)
        proposals = [None] * world
        dist.all_gather_object(proposals, int(seg_cap), group=group)
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.seg_cap = (max(proposals) + 1) & ~1
        # Two receive buffers used in turn: a rank that is one step ahead (it may start the next scatter as
        # soon as the count exchange of this step is over) must not overwrite keys its peers are still
        # counting.  Nobody can be two steps ahead -- the next count exchange needs everyone.
        self.nbuf = 1 if os.environ.get("EULER_B200_PEER_DOUBLE", "1") == "0" else 2
        self.phase = 0
        self.bases, self.base_ptr, handle, err = [], None, None, None
        try:
            self.base_ptr, handle = ctx.dist_recv_alloc(self.seg_cap * world * self.nbuf)
        except Exception as e:    # still take part in every collective below: a rank that left early would hang the others
            err = e
        self.local_ptr = self.base_ptr
        table = [None] * world
        dist.all_gather_object(table, (handle, err is None), group=group)
        if err is None and all(t[1] for t in table):
            try:
                for d in range(world):
                    self.bases.append(self.base_ptr if d == rank else ctx.dist_peer_open(table[d][0]))
            except Exception as e:
                err = e
        else:
            err = err or RuntimeError("a peer could not allocate its receive buffer")
        if err is not None:
            self.close()
            raise err
        self.next_step(first=True)

    def next_step(self, first=False):
        """switch to the other receive buffer (every rank calls this once per exchange, in lock step)"""
        import torch.distributed as dist
        if self.nbuf == 1 and not first:
            dist.barrier(group=self.group)     # single buffer: wait until every rank has consumed the last step
        shift = 8 * self.seg_cap * self.world * (self.phase % self.nbuf)
        self.phase += 1
        self.local_ptr = self.base_ptr + shift
        # my region inside every destination's current buffer
        self.dst_ptrs = [b + shift + 8 * self.seg_cap * self.rank for b in self.bases]

    def close(self):
        for d, b in enumerate(self.bases):
            if d != self.rank:
                try:
                    self.ctx.dist_peer_close(b)
                except Exception:
                    pass
        self.bases = []


def _peer_exchange(ctx, rank, world, seg_cap, group):
    """The key-exchange object of this context, created collectively on first use and kept ON the context (a new
    Context can never inherit mappings of a freed buffer).  The transport is agreed collectively: every rank reports
    whether its allocation and IPC opens succeeded, and if any failed all of them close their mappings and use NCCL.
    A rank that needs more than the agreed capacity reports the overflow through the count exchange, every rank takes
    the exact-size fallback for that step and drops the object (`_drop_peer_exchange`), and the next step agrees on a
    larger one (the cached capacity is in 8-byte words, so a change of key width re-agrees as well)."""
    import torch.distributed as dist
    px = getattr(ctx, "_peer_exchange", None)
    if px is not None and px is not False and (px.world != world or px.seg_cap < seg_cap):
        px.close()
        px = None
    if px is None:
        err = None
        try:
            px = PeerExchange(ctx, rank, world, max(int(seg_cap), getattr(ctx, "_peer_min", 0)), group)
        except Exception as e:          # no peer access from this rank: everyone must fall back together
            err, px = e, None
        oks = [None] * world
        dist.all_gather_object(oks, err is None, group=group)
        if not all(oks):
            if px is not None:
                px.close()
            px = False                  # remembered: NCCL all_to_all from now on
        ctx._peer_exchange = px
    return px if px else None


def _drop_peer_exchange(ctx, world):
    px = getattr(ctx, "_peer_exchange", None)
    if px:
        ctx._peer_min = int(px.seg_cap * 1.5)
        px.close()
    ctx._peer_exchange = None


def build_partitioned(ctx, d_reads, d_off, nreads, n_bases, l, rank, world, distinct_hint=0, group=None, slack=1.25,
                      use_peer=True):
    """Run the partitioned hot path on this rank.  d_reads / d_off are CUDA tensors (uint8 / int64).

    One pass over the reads scatters the canonical l-mer keys into `world` fixed-capacity segments
    (expected size x `slack`): with peer access straight into the owners' receive buffers over
    NVLink, otherwise into a local send buffer followed by an NCCL all_to_all.  If a segment
    overflows (skewed minimizers) the exchange is redone with exact sizes over NCCL.
    Returns (stats, info)."""
    import time
    import torch
    import torch.distributed as dist
    dev = d_reads.device
    # The record exchange + per-bucket build is the fast path while a rank's table stays in the small-bucket regime
    # (a known distinct count of at most BKT_MAX_DISTINCT canonical l-mers per rank; the caller's hint is the same on
    # every rank, so the choice is collective).  Larger or unknown tables take the key exchange below with its L2-blocked
    # global tables.  EULER_B200_DIST_KEYS=1 / 0 forces the one or the other.
    forced = os.environ.get("EULER_B200_DIST_KEYS")
    bucketed = (0 < int(distinct_hint) <= BKT_MAX_DISTINCT) if forced is None else forced == "0"
    if world > 1 and l <= 32 and use_peer and bucketed:
        res = build_partitioned_bucketed(ctx, d_reads, d_off, nreads, n_bases, l, rank, world, distinct_hint, group)
        if res is not None:
            return res
    kw = 2 if l > 32 else 1          # 8-byte words per key (128-bit keys above l = 32, csrc/wide_dist.cu)
    t0 = time.perf_counter()
    if world == 1 and kw == 2:
        cap1 = max(n_bases - nreads * (l - 1), 1) + 16
        send = _buffer("send", cap1 * kw, dev)
        counts = ctx.dist_scatter_segments(d_reads.data_ptr(), d_off.data_ptr(), nreads, n_bases, l, 1, send.data_ptr(), cap1)
        st = ctx.dist_build(send.data_ptr(), int(counts[0]), l, 0, 1, distinct_hint)
        return st, {"n_lmer_windows": int(counts[1]), "n_kmer_windows": int(counts[2]), "sent_keys": int(counts[0]),
                    "recv_keys": int(counts[0]), "exchange_bytes": 0, "exact_fallback": False, "transport": "none"}
    if world == 1:
        counts = ctx.dist_count(d_reads.data_ptr(), d_off.data_ptr(), nreads, n_bases, l, 1)
        send = _buffer("send", int(counts[0]), dev)
        ctx.dist_scatter(d_reads.data_ptr(), d_off.data_ptr(), nreads, n_bases, l, 1, send.data_ptr(), np.zeros(1, np.uint64))
        ctx.sync()
        st = ctx.dist_build(send.data_ptr(), int(counts[0]), l, 0, 1, distinct_hint)
        return st, {"n_lmer_windows": int(counts[1]), "n_kmer_windows": int(counts[2]), "sent_keys": int(counts[0]),
                    "recv_keys": int(counts[0]), "exchange_bytes": 0, "exact_fallback": False, "transport": "none"}
    # upper bound of windows; ~1.05 copies of each go out with minimizer ownership
    windows_ub = max(n_bases - nreads * (l - 1), 1)
    seg_cap = (int(windows_ub * slack * 1.15 / world) + 4096 + 1) & ~1
    px = _peer_exchange(ctx, rank, world, seg_cap * kw, group) if use_peer else None
    if px is not None:
        seg_cap = px.seg_cap // kw
        if kw == 2 and px.seg_cap % 2:     # cached buffer with an odd word stride: regions would not be 16-byte aligned
            px = None
    if px is not None:
        px.next_step()
        counts = ctx.dist_scatter_peers(d_reads.data_ptr(), d_off.data_ptr(), nreads, n_bases, l, world, px.dst_ptrs, seg_cap)
    else:
        send = _buffer("send", seg_cap * world * kw, dev)
        counts = ctx.dist_scatter_segments(d_reads.data_ptr(), d_off.data_ptr(), nreads, n_bases, l, world, send.data_ptr(),
                                           seg_cap)
    send_counts = counts[:world].astype(np.int64).tolist()
    n_l, n_k = int(counts[world]), int(counts[world + 1])
    t1 = time.perf_counter()
    # the count exchange doubles as the barrier that makes every rank's peer stores visible
    msg = send_counts + [1 if max(send_counts) > seg_cap else 0]
    table = torch.tensor(msg, dtype=torch.int64, device=dev)
    gathered = torch.empty((world, world + 1), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered, table, group=group)
    gathered = gathered.cpu().numpy()
    recv_counts = [int(gathered[src, rank]) for src in range(world)]
    exact = bool(gathered[:, world].max())
    t2 = time.perf_counter()
    if exact:
        _drop_peer_exchange(ctx, world)     # collective knowledge: every rank sees the flag
    if exact and kw == 2:
        raise RuntimeError("128-bit keys: a destination segment overflowed (skewed minimizers); raise `slack`")
    if exact:   # rare: redo with exact sizes on every rank, over NCCL
        counts = ctx.dist_count(d_reads.data_ptr(), d_off.data_ptr(), nreads, n_bases, l, world)
        send_counts = counts[:world].astype(np.int64).tolist()
        send_off = np.zeros(world, np.uint64)
        send_off[1:] = np.cumsum(send_counts[:-1])
        send = _buffer("send", int(sum(send_counts)), dev)
        ctx.dist_scatter(d_reads.data_ptr(), d_off.data_ptr(), nreads, n_bases, l, world, send.data_ptr(), send_off)
        ctx.sync()
        recv_counts = torch_count_exchange(group)(send_counts)
        starts = [int(x) for x in send_off]
        px = None
    else:
        starts = [d * seg_cap for d in range(world)]
    nkeys = int(sum(recv_counts))
    if px is not None:
        t3 = time.perf_counter()
        st = ctx.dist_build_regions(px.local_ptr, seg_cap, recv_counts, l, rank, world, distinct_hint)
        transport = "peer stores over NVLink (CUDA IPC), fused into the scatter kernel"
    else:
        recv = _buffer("recv", nkeys * kw, dev)
        out_list, in_list, pos = [], [], 0
        for src in range(world):
            out_list.append(recv[pos * kw:(pos + recv_counts[src]) * kw])
            pos += recv_counts[src]
        for d in range(world):
            in_list.append(send[starts[d] * kw:(starts[d] + send_counts[d]) * kw])
        dist.all_to_all(out_list, in_list, group=group)
        torch.cuda.current_stream().synchronize()
        t3 = time.perf_counter()
        st = ctx.dist_build(recv.data_ptr(), nkeys, l, rank, world, distinct_hint)
        transport = "NCCL all_to_all"
    t4 = time.perf_counter()
    info = {"n_lmer_windows": n_l, "n_kmer_windows": n_k, "sent_keys": int(sum(send_counts)), "recv_keys": nkeys,
            "exchange_bytes": 8 * kw * (int(sum(send_counts)) - send_counts[rank]), "exact_fallback": exact, "transport": transport,
            "phase_ms": {"partition": 1e3 * (t1 - t0), "count_exchange": 1e3 * (t2 - t1), "all_to_all": 1e3 * (t3 - t2),
                         "build": 1e3 * (t4 - t3)}}
    return st, info


# ---------------------------------------------------------------------------------------------------------------
# The minimizer-bucketed form (csrc/bucket.cuh): what crosses NVLink is the reads cut into 2-bit packed minimizer
# runs (16-byte records of up to 17 l-mers), stored by the partition kernel straight into the owners' bucket
# regions; the owner then builds every bucket in shared memory.  Used for l <= 32.
BKT_CAP = 1536   # slots of a per-bucket shared-memory table (csrc/pipeline.cu default)
BKT_MAX_DISTINCT = 50_000_000   # per-rank distinct canonical l-mers up to which the bucketed path is taken by default


def plan_buckets(n_bases, l, world, distinct_hint=0, cap=None):
    """(buckets per rank, records per (destination, source) stream) for shards of at most n_bases bases per rank.
    distinct_hint = expected distinct canonical l-mers PER RANK (0 = unknown: every window distinct)."""
    cap = cap or int(os.environ.get("EULER_B200_BKT_CAP", BKT_CAP))
    k = l - 1
    w = k - min(k, 12) + 1
    rec_per_base = 2.0 / (w + 1.0) + 1.0 / 16.0 + 0.01
    est = int(distinct_hint) or max(int(n_bases), 1)
    # a hint that is far too small must not pile a whole shard's worth of l-mers into a handful of shared-memory
    # tables (same floor as pipeline_run_bucketed): at least one bucket per 128 Ki bases
    nbpr = max(int(est * 1.06 / (0.30 * cap)) + 1, int(n_bases) >> 17)
    scap = int(n_bases * rec_per_base / world * 1.5) + 4096
    return nbpr, scap


class BucketExchange:
    """Geometry agreed by all ranks, two peer-visible receive areas per rank (step parity: a rank may scatter step
    i+1 while a peer still builds step i; nobody can be two steps ahead, the flag exchange needs everyone) and the
    peers' areas opened through CUDA IPC.  Created collectively; `ok` is the COLLECTIVE outcome, so either every rank
    uses the object or none does."""

    def __init__(self, ctx, rank, world, nb_per_rank, scap, group=None):
        import torch
        import torch.distributed as dist
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        props = [None] * world
        dist.all_gather_object(props, (int(nb_per_rank), int(scap)), group=group)
        self.nb_per_rank = max(p[0] for p in props)
        self.scap = max(p[1] for p in props)
        self.phase = 0
        self.plan_key = (0, 0)   # (l, largest shard in bases) the geometry was planned for
        self.local, self.areas, self.opened = [None, None], [None, None], []
        handles, err = [None, None], None
        try:
            for which in (0, 1):
                self.local[which], handles[which] = ctx.bkt_area_alloc(which, world, self.scap)
        except Exception as e:   # reported to everyone below
            err = e
        table = [None] * world
        dist.all_gather_object(table, (handles, err is None), group=group)
        if all(t[1] for t in table):
            try:
                for which in (0, 1):
                    ptrs = []
                    for d in range(world):
                        if d == rank:
                            ptrs.append(self.local[which])
                        else:
                            ptr = ctx.dist_peer_open(table[d][0][which])
                            self.opened.append(ptr)
                            ptrs.append(ptr)
                    self.areas[which] = ptrs
            except Exception as e:
                err = e
        oks = [None] * world
        dist.all_gather_object(oks, err is None, group=group)
        self.ok = all(oks)
        self.error = err
        if not self.ok:
            self.close()

    def close(self):
        for ptr in self.opened:
            try:
                self.ctx.dist_peer_close(ptr)
            except Exception:
                pass
        self.opened = []


def _bucket_exchange(ctx, rank, world, nbpr, scap, group):
    """the exchange object kept ON the context (closed with it), re-created collectively when the geometry grows"""
    bx = getattr(ctx, "_bucket_exchange", None)
    if bx is not None and bx.world == world and bx.ok:
        return bx
    bx = BucketExchange(ctx, rank, world, nbpr, scap, group)
    ctx._bucket_exchange = bx
    return bx


def build_partitioned_bucketed(ctx, d_reads, d_off, nreads, n_bases, l, rank, world, distinct_hint=0, group=None, _retry=0):
    """One step of the bucketed multi-GPU path on this rank.  d_reads / d_off: CUDA tensors.  The caller's current
    torch stream must be the ctx stream (collectives and kernels are ordered on it).  Returns (stats, info), or None
    when the ranks have no peer access to each other (the caller then takes the key exchange).

    Every decision that changes what the ranks do next (re-planning the geometry, redoing the scatter) is taken from
    words that were MAX-reduced over all ranks, so all ranks take it together."""
    import time
    import torch
    import torch.distributed as dist
    dev = d_reads.device
    t0 = time.perf_counter()
    bx = getattr(ctx, "_bucket_exchange", None)
    if bx is None or bx.world != world:
        # first use: the plan needs the largest shard, which only a collective knows
        nmax = torch.tensor([int(n_bases)], dtype=torch.int64, device=dev)
        dist.all_reduce(nmax, op=dist.ReduceOp.MAX, group=group)
        n_plan = int(nmax.item())
        nbpr, scap = plan_buckets(n_plan, l, world, distinct_hint)
        bx = _bucket_exchange(ctx, rank, world, nbpr, scap, group)
        bx.plan_key = (int(l), n_plan)
        ctx._bucket_want = 0
    if not bx.ok:
        return None
    which = bx.phase & 1
    bx.phase += 1
    words = _buffer("bkt_words", 8, dev)
    serr = None
    try:
        ctx.bkt_scatter(d_reads.data_ptr(), d_off.data_ptr(), nreads, n_bases, l, rank, world, bx.nb_per_rank, bx.scap, bx.areas[which],
                        d_out=words.data_ptr())
    except Exception as e:   # reported through the collective below: every rank raises, nobody waits
        serr = e
        words.zero_()
    # flags, fullest stream, wanted geometry, shard size, l: MAX over ranks.  The collective is also the barrier after
    # which every rank's peer stores are complete (each rank's scatter precedes its contribution on its stream).
    extra = torch.tensor([getattr(ctx, "_bucket_want", 0), int(n_bases), int(l), 0 if serr is None else 1], dtype=torch.int64, device=dev)
    msg = torch.cat([words[2:4], extra])
    dist.all_reduce(msg, op=dist.ReduceOp.MAX, group=group)
    host = torch.cat([words[:2], msg]).cpu().numpy()   # the one host round trip between scatter and build
    t1 = time.perf_counter()
    n_l, n_k, flags, max_region, want_all, n_max, l_max, failed = (int(x) for x in host)
    if failed:
        raise RuntimeError("bucketed scatter failed on rank %d: %s" % (rank, serr) if serr is not None
                           else "bucketed scatter failed on another rank")
    stale = bx.plan_key[0] != l_max or n_max > 2 * bx.plan_key[1] or 2 * n_max < bx.plan_key[1]   # planned for another workload
    regeom = want_all and (want_all > 2 * bx.nb_per_rank or 2 * want_all < bx.nb_per_rank)
    if (flags & 0x10) or stale or regeom:
        # collective knowledge (everyone sees the same reduced words): re-plan and redo this step
        if _retry >= 4:
            raise RuntimeError("bucketed exchange: the geometry did not settle")
        if stale:
            nbpr, scap = plan_buckets(n_max, l, world, distinct_hint)
        else:
            nbpr = want_all if regeom else bx.nb_per_rank
            scap = int(max_region * 1.25) + 4096 if (flags & 0x10) else bx.scap
        bx.close()
        bx.ok = False
        ctx._bucket_exchange = None
        ctx._bucket_want = 0
        bx = _bucket_exchange(ctx, rank, world, nbpr, scap, group)
        bx.plan_key = (int(l), n_max)
        if not bx.ok:
            return None
        return build_partitioned_bucketed(ctx, d_reads, d_off, nreads, n_bases, l, rank, world, distinct_hint, group, _retry + 1)
    # the build can fail on ONE rank (memory, a bucket that does not fit shared memory): the outcome is made
    # collective before anyone moves on, or the other ranks would wait for it in the next collective for ever
    st, err = None, None
    try:
        st = ctx.bkt_build(bx.local[which], l, rank, world, bx.nb_per_rank, bx.scap, distinct_hint)
    except Exception as e:
        err = e
    okw = torch.tensor([0 if err is None else 1], dtype=torch.int64, device=dev)
    dist.all_reduce(okw, op=dist.ReduceOp.MAX, group=group)
    if int(okw.item()):
        raise RuntimeError("bucketed build failed on rank %d: %s" % (rank, err) if err is not None
                           else "bucketed build failed on another rank")
    t2 = time.perf_counter()
    # geometry for the next steps from what was counted (reported at the next exchange, adopted by all ranks together)
    cap = int(os.environ.get("EULER_B200_BKT_CAP", BKT_CAP))
    ctx._bucket_want = int((st.distinct_lmers + 1) // 2 * 1.06 / (0.30 * cap)) + 1
    info = {"n_lmer_windows": n_l, "n_kmer_windows": n_k, "sent_keys": 0, "recv_keys": 0,
            "exchange_bytes": int(16 * max_region * (world - 1)),   # upper bound: the fullest stream x the peers
            "exact_fallback": False,
            "transport": "16-byte minimizer-run records stored as runs into one stream per destination rank over NVLink (CUDA IPC), regrouped by the owner",
            "geometry": {"nb_per_rank": bx.nb_per_rank, "scap": bx.scap, "fullest_stream": max_region, "fullest_bucket": int(st.bucket_records)},
            "phase_ms": {"partition+exchange": 1e3 * (t1 - t0), "build": 1e3 * (t2 - t1), "scatter_kernel": float(st.ms_count),
                         "build_kernel": float(st.ms_build_kernel)}}
    return st, info


def emulate_partitioned_bucketed(ctx, shards, l, world, nb_per_rank=None, scap=None):
    """Single-process emulation of `world` ranks on one GPU (tests) through the entry points the production path
    uses (euler_bkt_scatter with one area per destination, euler_bkt_build per area).  Same return value as
    emulate_partitioned."""
    import torch
    import _native as N
    n_max = max(int(off[-1]) for _, off in shards) if shards else 0
    a, b = plan_buckets(max(n_max, 1), l, world)
    nbpr, scap = nb_per_rank or a, scap or b
    while True:
        nbytes = world * scap * 16 + world * 8 + 256
        areas = [torch.zeros(nbytes // 8 + 1, dtype=torch.int64, device="cuda") for _ in range(world)]
        ptrs = [t.data_ptr() for t in areas]
        windows, worst, overflow = [], 0, False
        for r, (buf, off) in enumerate(shards):
            d_buf = torch.from_numpy(np.ascontiguousarray(buf)).cuda() if len(buf) else torch.zeros(16, dtype=torch.uint8, device="cuda")
            d_off = torch.from_numpy(np.ascontiguousarray(off).astype(np.int64)).cuda()
            out = ctx.bkt_scatter(d_buf.data_ptr(), d_off.data_ptr(), len(off) - 1, int(off[-1]), l, r, world, nbpr, scap, ptrs)
            overflow |= bool(int(out[2]) & 0x10)
            worst = max(worst, int(out[3]))
            windows.append((int(out[0]), int(out[1])))
        if not overflow:
            break
        scap = int(worst * 1.25) + 16     # what the production path does collectively (build_partitioned_bucketed)
    ctx.sync()
    res = []
    for d in range(world):
        st = ctx.bkt_build(ptrs[d], l, d, world, nbpr, scap, 0)
        names = ("LMER_KEYS", "LMER_VALUES", "LMER_OFFSETS", "KMER_KEYS", "LCOUNT", "ECOUNT", "LSTART", "ESTART", "EV", "EDGE_V1", "EDGE_V2")
        art = {name: ctx.download(getattr(N, "ART_" + name)) for name in names}
        art["stats"] = st.as_dict()
        art["recv_keys"] = 0
        res.append(art)
    return res, windows


def emulate_partitioned(ctx, shards, l, world):
    """Single-process emulation of `world` ranks on one GPU (tests): shards = list of
    (uint8 reads array, uint64 offsets array) per rank.  Returns the per-rank artefact dicts."""
    import torch
    import _native as N
    buckets = [[None] * world for _ in range(world)]
    windows = []
    for r, (buf, off) in enumerate(shards):
        d_buf = torch.from_numpy(np.ascontiguousarray(buf)).cuda() if len(buf) else torch.zeros(16, dtype=torch.uint8, device="cuda")
        d_off = torch.from_numpy(np.ascontiguousarray(off).astype(np.int64)).cuda()
        nreads, n_bases = len(off) - 1, int(off[-1])
        if l > 32:      # 16-byte keys: fixed-capacity segments (csrc/wide_dist.cu), two int64 words per key
            cap = max(n_bases, 1) + 16
            send = torch.empty(cap * world * 2, dtype=torch.int64, device="cuda")
            counts = ctx.dist_scatter_segments(d_buf.data_ptr(), d_off.data_ptr(), nreads, n_bases, l, world, send.data_ptr(), cap)
            windows.append((int(counts[world]), int(counts[world + 1])))
            for d in range(world):
                buckets[r][d] = send[2 * cap * d:2 * (cap * d + int(counts[d]))].clone()
            continue
        counts = ctx.dist_count(d_buf.data_ptr(), d_off.data_ptr(), nreads, n_bases, l, world)
        sc = counts[:world].astype(np.int64)
        windows.append((int(counts[world]), int(counts[world + 1])))
        send_off = np.zeros(world, np.uint64)
        send_off[1:] = np.cumsum(sc[:-1])
        send = torch.empty(max(int(sc.sum()), 1), dtype=torch.int64, device="cuda")
        ctx.dist_scatter(d_buf.data_ptr(), d_off.data_ptr(), nreads, n_bases, l, world, send.data_ptr(), send_off)
        ctx.sync()
        for d in range(world):
            buckets[r][d] = send[int(send_off[d]):int(send_off[d]) + int(sc[d])].clone()
    out = []
    for d in range(world):
        recv = torch.cat([buckets[r][d] for r in range(world)]) if world else None
        if recv.numel() == 0:
            recv = torch.zeros(1, dtype=torch.int64, device="cuda")
            nkeys = 0
        else:
            nkeys = recv.numel() // (2 if l > 32 else 1)
        st = ctx.dist_build(recv.data_ptr(), nkeys, l, d, world, 0)
        names = ("LMER_KEYS", "LMER_VALUES", "LMER_OFFSETS", "KMER_KEYS", "LCOUNT", "ECOUNT", "LSTART", "ESTART", "EV",
                 "EDGE_V1", "EDGE_V2") + (("LMER_KEYS_HI", "KMER_KEYS_HI") if l > 32 else ())
        art = {name: ctx.download(getattr(N, "ART_" + name)) for name in names}
        art["stats"] = st.as_dict()
        art["recv_keys"] = nkeys
        out.append(art)
    return out, windows
