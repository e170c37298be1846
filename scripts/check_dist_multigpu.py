"""Real multi-GPU check of the partitioned path with UNEVEN shards (run under torchrun on >= 2 GPUs):
FASTQ bytes split at record boundaries -> device ingestion -> partition + peer exchange -> per-rank build;
the union of the per-rank both-strand k-mer tables must equal the oracle's table of all reads."""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/pycuda-euler_b200'); sys.path.insert(0,'/root/repo/tests')
import oracle, _native as N
from eulercuda.dist import build_partitioned
from assembler import split_records
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
ctx = N.Context(rank)
G, L, cov = 300000, 100, 25
n = G*cov//L
r = oracle.synth_reads(G, L, err_ppm=3000, first=0, count=n)
data = b"".join(b'@r%d\n%s\n+\n%s\n' % (i, bytes(r[i*L:(i+1)*L]), b'I'*L) for i in range(n))
a, b = split_records(data, world, 2)[rank]
nreads, nbases = ctx.ingest(data[a:b], 2)
buf, off = ctx.ingest_download()
first = 0 if rank == 0 else n - nreads
ok_reads = np.array_equal(buf, r[first*L:(first+nreads)*L]) and np.array_equal(off, oracle.fixed_offsets(nreads, L))
print(rank, "ingested reads equal source slice:", ok_reads, nreads, nbases, flush=True)
K = 31
for variant in ("ingested", "direct"):
    src = buf if variant == "ingested" else r[first*L:(first+nreads)*L]
    d_reads = torch.zeros(len(src)+16, dtype=torch.uint8, device="cuda"); d_reads[:len(src)] = torch.from_numpy(np.ascontiguousarray(src)).cuda()
    d_off = torch.from_numpy(np.ascontiguousarray(off).astype(np.int64)).cuda()
    st, info = build_partitioned(ctx, d_reads, d_off, nreads, nbases, K, rank, world, 0)
    k = ctx.download(N.ART_LMER_KEYS); v = ctx.download(N.ART_LMER_VALUES)
    g = [None]*world if rank == 0 else None
    dist.gather_object((k, v), g, dst=0)
    if rank == 0:
        ak = np.concatenate([x[0] for x in g]); av = np.concatenate([x[1] for x in g])
        lo, hi, vals = oracle.count_mers(r, oracle.fixed_offsets(n, L), K)
        o = np.argsort(ak, kind="stable")
        same = ak.size == lo.size and np.array_equal(ak[o], lo) and np.array_equal(av[o], vals)
        print(variant, "oracle", lo.size, "dist", ak.size, "EQUAL" if same else "DIFF", flush=True)
        if not same and ak.size == lo.size and np.array_equal(ak[o], lo):
            bad = np.flatnonzero(av[o] != vals)
            for i in bad[:4]:
                km = oracle.decode_key(lo[i], 0, K)
                s = bytes(r).decode()
                pos = [p for p in range(len(s)) if s.startswith(km, p)][:5]
                print("  kmer", km, "oracle", vals[i], "dist", av[o][i], "positions", [(p // L, p % L) for p in pos], flush=True)
dist.barrier(); dist.destroy_process_group()
