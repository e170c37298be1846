// dist.cu -- k-mer-space partition across the GPUs of one box (the reference has no
// analogue: its only parallelism is Spark mapPartitions over independent read partitions,
// src/cli_spark_gpu.py:37).
//
// Ownership: a vertex (k-mer) v belongs to rank owner(canon(v)).  Every forward l-mer window x is
// sent, as its canonical key, to the owner of its prefix k-mer and to the owner of its suffix
// k-mer (once if they coincide).  After ONE all-to-all each rank therefore holds, with full
// multiplicity, exactly the l-mers incident to its vertices, and builds its part of the graph with
// no further data-path exchange:
//   * vertex table / ids / lcount / ecount / lstart / estart / EulerVertex of the vertices it owns;
//   * the edge record of a (both-strand) l-mer x lives on owner(prefix(x)) ("home"), so v1 is always
//     local; v2 is local when the suffix has the same owner, else EULER_NO_ID (resolved at the
//     component merge, where per-GPU tables are joined).
// Global ids = local id + exclusive scan of the per-rank counts (done by the caller).
#include "encode.cuh"
#include "kernels.h"
#include "scan.cuh"

#define DB 256

__device__ __forceinline__ u32 owner_of(u64 canon_kmer, u32 nranks)
{
    const u64 h = (canon_kmer ^ (canon_kmer >> 31)) * 0xD6E8FEB86659FD93ull;
    return (u32)(((h >> 32) * (u64)nranks) >> 32);
}
__device__ __forceinline__ u32 owner_kmer(u64 v, u32 k, u32 nranks)
{
    const u64 r = revcomp64(v, k);
    return owner_of(v < r ? v : r, nranks);
}

// ---- pass over the reads: per-destination counts (SCATTER=false) or key scatter (SCATTER=true) --
template <bool SCATTER>
__global__ void __launch_bounds__(DB, 4) dist_partition_kernel(const uint4 *__restrict__ buf16, u64 n_bases,
                                                                const u32 *__restrict__ start_bits, u32 l, u32 nranks, u64 ntiles,
                                                                u64 *__restrict__ counts /* [nranks] + N_l + N_k */,
                                                                u64 *__restrict__ cursors, u64 *__restrict__ send)
{
    const int lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const u32 k = l - 1, top = 2 * (l - 1);
    const u64 lmask = key_mask_d(l), kmask = key_mask_d(k);
    const unsigned lt = (1u << lane) - 1u;
    u32 nl_tot = 0, nk_tot = 0;
    u32 cnt_local[16];
#pragma unroll
    for (int d = 0; d < 16; d++) cnt_local[d] = 0;

    for (u64 tile = warp; tile < ntiles; tile += nwarps) {
        const long long chunk = (long long)(tile * ENC_ADV) - ENC_HALO + lane;
        const Chunk c = load_chunk(buf16, chunk, n_bases, start_bits);
        const u32 p1 = __shfl_up_sync(0xffffffffu, c.codes, 1), p2 = __shfl_up_sync(0xffffffffu, c.codes, 2);
        const u32 v1 = __shfl_up_sync(0xffffffffu, c.vmask, 1), v2 = __shfl_up_sync(0xffffffffu, c.vmask, 2);
        const u32 s1 = __shfl_up_sync(0xffffffffu, c.smask, 1), s2 = __shfl_up_sync(0xffffffffu, c.smask, 2);
        u64 f = ((u64)p2 << 32) | p1;
        u64 rc = revcomp64(f & lmask, l);
        const u32 pv = (v2 << 16) | v1, ps = (s2 << 16) | s1;
        u32 vrun = (pv == 0xffffffffu) ? 32u : (u32)__ffs(~pv) - 1u;
        u32 srun = ps ? (u32)__ffs(ps) - 1u : 32u;
        u32 codes = c.codes;
        u32 vm = (lane < ENC_HALO) ? 0u : (c.vmask << 16);
        u32 sm = c.smask << 16;
        if (lane < ENC_HALO) vrun = 0;
#pragma unroll 1
        for (int i = 0; i < 16; i++) {
            const u32 cc = codes >> 30;
            codes <<= 2;
            const bool valid = (vm >> 31) != 0, start = (sm >> 31) != 0;
            vm <<= 1;
            sm <<= 1;
            f = (f << 2) | cc;
            rc = (rc >> 2) | ((u64)(3u - cc) << top);
            vrun = valid ? vrun + 1u : 0u;
            srun = start ? 0u : srun + 1u;
            nk_tot += (vrun >= k && srun + 1u >= k) ? 1u : 0u;
            const bool ok = vrun >= l && srun + 1u >= l;
            nl_tot += ok ? 1u : 0u;
            const u64 fm = f & lmask;
            // rc(prefix x) = suffix(rc x), rc(suffix x) = prefix(rc x): owners come for free
            const u64 p = fm >> 2, s = fm & kmask, rp = rc & kmask, rs = rc >> 2;
            const u32 o1 = ok ? owner_of(p < rp ? p : rp, nranks) : 0xffffffffu;
            u32 o2 = ok ? owner_of(s < rs ? s : rs, nranks) : 0xffffffffu;
            if (o2 == o1) o2 = 0xffffffffu;
            const u64 key = fm < rc ? fm : rc;
#pragma unroll
            for (int pass = 0; pass < 2; pass++) {
                const u32 o = pass ? o2 : o1;
                const unsigned peers = __match_any_sync(0xffffffffu, o);
                if (o == 0xffffffffu) continue;
                const int leader = __ffs(peers) - 1;
                if (!SCATTER) {
                    if (lane == leader) {
#pragma unroll
                        for (int d = 0; d < 16; d++)
                            if ((u32)d == o) cnt_local[d] += __popc(peers);
                    }
                } else {
                    u64 base = 0;
                    if (lane == leader) base = atomicAdd(cursors + o, (u64)__popc(peers));
                    base = __shfl_sync(peers, base, leader);
                    send[base + __popc(peers & lt)] = key;
                }
            }
        }
    }
    if (!SCATTER) {
#pragma unroll
        for (int d = 0; d < 16; d++) {
            u32 t = cnt_local[d];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0 && t && (u32)d < nranks) atomicAdd(counts + d, (u64)t);
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            nl_tot += __shfl_xor_sync(0xffffffffu, nl_tot, o);
            nk_tot += __shfl_xor_sync(0xffffffffu, nk_tot, o);
        }
        if (lane == 0) {
            if (nl_tot) atomicAdd(counts + 16, (u64)nl_tot);
            if (nk_tot) atomicAdd(counts + 17, (u64)nk_tot);
        }
    }
}

int dist_partition(euler_ctx *ctx, bool scatter, const void *d_buf, u64 n_bases, const u32 *d_bits, u32 l, u32 nranks,
                   u64 *d_counts, u64 *d_cursors, u64 *d_send)
{
    if (!n_bases) return EULER_OK;
    const u64 nchunks = (n_bases + 15) / 16;
    const u64 ntiles = (nchunks + ENC_ADV - 1) / ENC_ADV;
    u64 grid = (u64)ctx->num_sms * 4;
    const u64 need = (ntiles + DB / 32 - 1) / (DB / 32);
    if (grid > need) grid = need;
    if (!scatter)
        dist_partition_kernel<false><<<(unsigned)grid, DB, 0, ctx->stream>>>((const uint4 *)d_buf, n_bases, d_bits, l, nranks,
                                                                              ntiles, d_counts, nullptr, nullptr);
    else
        dist_partition_kernel<true><<<(unsigned)grid, DB, 0, ctx->stream>>>((const uint4 *)d_buf, n_bases, d_bits, l, nranks,
                                                                             ntiles, nullptr, d_cursors, d_send);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- count received canonical keys into the local table ----------------------------------------
__global__ void __launch_bounds__(DB, 4) dist_count_keys_kernel(const u64 *__restrict__ keys_in, u64 n, u64 *__restrict__ tab_keys,
                                                                 u32 *__restrict__ tab_cnt, u64 cap, u64 *__restrict__ stats)
{
    const u32 nbuckets = (u32)(cap / EULER_BUCKET);
    const u32 max_probe = nbuckets < 4096 ? nbuckets : 4096;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 t0 = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool overflow = false;
    // every lane of a warp runs the same number of iterations (the loop bound is warp-uniform)
    const u64 iters = (n + stride * 4 - 1) / (stride * 4);
    for (u64 it = 0; it < iters; it++) {
        u64 key[4];
        u32 bucket[4];
        K4 q[4];
        u32 pend = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const u64 idx = (it * 4 + i) * stride + t0;
            key[i] = idx < n ? keys_in[idx] : 0;
            if (idx < n) pend |= 1u << i;
            bucket[i] = (u32)hash_bucket(key[i], nbuckets);
        }
        u32 probes = 0;
        while (__any_sync(0xffffffffu, pend != 0)) {
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (pend & (1u << i)) q[i] = ld_bucket_cg(tab_keys + (u64)bucket[i] * EULER_BUCKET);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (pend & (1u << i)) {
                    const int j = bucket_claim(tab_keys + (u64)bucket[i] * EULER_BUCKET, q[i], key[i]);
                    if (j >= 0) {
                        atomicAdd(tab_cnt + (u64)bucket[i] * EULER_BUCKET + j, 1u);
                        pend &= ~(1u << i);
                    } else if (++bucket[i] == nbuckets) {
                        bucket[i] = 0;
                    }
                }
            }
            if (++probes >= max_probe && pend) { overflow = true; pend = 0; }
        }
    }
    if (overflow) atomicOr((unsigned long long *)(stats + 2), 1ull);
}

int dist_count_keys(euler_ctx *ctx, const u64 *d_keys, u64 n, u64 *tab_keys, u32 *tab_cnt, u64 cap, u64 *d_stats)
{
    if (!n) return EULER_OK;
    u64 grid = (u64)ctx->num_sms * 8;
    const u64 need = (n + DB * 4 - 1) / (DB * 4);
    if (grid > need) grid = need;
    dist_count_keys_kernel<<<(unsigned)grid, DB, 0, ctx->stream>>>(d_keys, n, tab_keys, tab_cnt, cap, d_stats);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- ownership-aware graph stage ----------------------------------------------------------------
__global__ void __launch_bounds__(DB) dist_vertex_insert_kernel(const u64 *__restrict__ lt_keys, u64 lt_cap, u32 l,
                                                                 u64 *__restrict__ vt_keys, u64 vt_cap, u32 rank, u32 nranks,
                                                                 u64 *flags)
{
    const u64 slot = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= lt_cap) return;
    const u64 key = lt_keys[slot];
    if (key == EULER_EMPTY_KEY) return;
    const u32 k = l - 1;
    const u64 kmask = key_mask_d(k);
    const u64 max_probe = vt_cap / EULER_BUCKET < 4096 ? vt_cap / EULER_BUCKET : 4096;
    const u64 p = key >> 2, s = key & kmask;
    const u64 rp = revcomp64(p, k), rs = revcomp64(s, k);
    const u64 cp = p < rp ? p : rp, cs = s < rs ? s : rs;
    bool bad = false;
    if (owner_of(cp, nranks) == rank) bad |= table_insert(vt_keys, vt_cap, cp, max_probe) == EULER_NO_SLOT;
    if (owner_of(cs, nranks) == rank) bad |= table_insert(vt_keys, vt_cap, cs, max_probe) == EULER_NO_SLOT;
    if (bad) atomicOr((unsigned long long *)flags, 2ull);
}

int dist_vertex_insert(euler_ctx *ctx, const u64 *lt_keys, u64 lt_cap, u32 l, u64 *vt_keys, u64 vt_cap, u32 rank, u32 nranks,
                       u64 *d_flags)
{
    dist_vertex_insert_kernel<<<grid_for(lt_cap, DB), DB, 0, ctx->stream>>>(lt_keys, lt_cap, l, vt_keys, vt_cap, rank, nranks,
                                                                            d_flags);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// strands of the slot's canonical l-mer that are homed on this rank: bit0 = c (prefix owned),
// bit1 = rc(c) (suffix owned; same record as bit0 for a palindrome)
__device__ __forceinline__ u32 homed_strands(u64 c, u32 l, u32 rank, u32 nranks, bool &pal)
{
    const u32 k = l - 1;
    pal = c == revcomp64(c, l);
    const bool own_p = owner_kmer(c >> 2, k, nranks) == rank;
    const bool own_s = owner_kmer(c & key_mask_d(k), k, nranks) == rank;
    return (own_p ? 1u : 0u) | ((own_s && !pal) ? 2u : 0u);
}

struct DistLtScanPolicy {
    typedef u64 T;
    const u64 *keys;
    const u32 *cnt;
    u32 l, rank, nranks;
    u32 *base, *eoff;
    __device__ __forceinline__ u64 load(u64 i) const
    {
        const u64 c = keys[i];
        if (c == EULER_EMPTY_KEY) return 0ull;
        bool pal;
        const u32 h = homed_strands(c, l, rank, nranks, pal);
        const u64 n = cnt[i];
        const u64 records = (h & 1u) + ((h >> 1) & 1u);
        const u64 edges = pal ? ((h & 1u) ? 2 * n : 0) : n * records;
        return (edges << 32) | records;
    }
    __device__ __forceinline__ void store(u64 i, u64 ex, u64, bool valid) const
    {
        if (valid) { base[i] = (u32)ex; eoff[i] = (u32)(ex >> 32); }
    }
};

int dist_lt_scan(euler_ctx *ctx, const u64 *lt_keys, const u32 *lt_cnt, u64 cap, u32 l, u32 rank, u32 nranks, u32 *base,
                 u32 *eoff, u64 *d_total_packed)
{
    return scan_run(ctx, DistLtScanPolicy{lt_keys, lt_cnt, l, rank, nranks, base, eoff}, cap, d_total_packed);
}

__global__ void __launch_bounds__(DB) dist_edges_kernel(const u64 *__restrict__ lt_keys, const u32 *__restrict__ lt_cnt,
                                                         const u32 *__restrict__ base, const u32 *__restrict__ eoff, u64 cap, u32 l,
                                                         VertexTable vt, u32 rank, u32 nranks, u64 *__restrict__ lkeys,
                                                         u32 *__restrict__ lvals, u32 *__restrict__ loffs, u32 *__restrict__ ev1,
                                                         u32 *__restrict__ ev2, u32 *__restrict__ lcount, u32 *__restrict__ ecount)
{
    const u64 slot = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= cap) return;
    const u64 c = lt_keys[slot];
    if (c == EULER_EMPTY_KEY) return;
    const u32 n = lt_cnt[slot];
    const u32 k = l - 1;
    const u64 kmask = key_mask_d(k);
    const u64 r = revcomp64(c, l);
    const bool pal = c == r;
    const u64 p = c >> 2, s = c & kmask;
    const u64 rp = revcomp64(p, k), rs = revcomp64(s, k);
    const u64 cp = p < rp ? p : rp, cs = s < rs ? s : rs;
    const bool own_p = owner_of(cp, nranks) == rank, own_s = owner_of(cs, nranks) == rank;
    u32 id_p = EULER_NO_ID, id_rp = EULER_NO_ID, id_s = EULER_NO_ID, id_rs = EULER_NO_ID;
    if (own_p) {
        const u64 sp = table_find(vt.keys, vt.cap, cp);
        if (sp == EULER_NO_SLOT) return;
        const u32 a = vt.id0[sp], b = (p == rp) ? a : a + 1u;
        id_p = (p == cp) ? a : b;
        id_rp = (p == cp) ? b : a;
    }
    if (own_s) {
        const u64 ss = table_find(vt.keys, vt.cap, cs);
        if (ss == EULER_NO_SLOT) return;
        const u32 a = vt.id0[ss], b = (s == rs) ? a : a + 1u;
        id_s = (s == cs) ? a : b;
        id_rs = (s == cs) ? b : a;
    }
    const u32 m0 = pal ? 2u * n : n;
    u32 idx = base[slot], eo = eoff[slot];
    const u32 first_c = (u32)((c >> (2 * k)) & 3), first_r = (u32)((r >> (2 * k)) & 3);
    if (own_p) {  // strand c is homed here (leaves p); rc(c) enters rc(p), which is ours too
        lkeys[idx] = c; lvals[idx] = m0; loffs[idx] = eo; ev1[idx] = id_p; ev2[idx] = id_s;
        lcount[4ull * id_p + (u32)(c & 3)] = m0;
        if (!pal) ecount[4ull * id_rp + first_r] = n;
        idx++; eo += m0;
    }
    if (own_s) {  // c enters s (ours); rc(c) leaves rc(s) and is homed here
        ecount[4ull * id_s + first_c] = m0;
        if (!pal) {
            lkeys[idx] = r; lvals[idx] = n; loffs[idx] = eo; ev1[idx] = id_rs; ev2[idx] = id_rp;
            lcount[4ull * id_rs + (u32)(r & 3)] = n;
        }
    }
}

int dist_edges(euler_ctx *ctx, const u64 *lt_keys, const u32 *lt_cnt, const u32 *base, const u32 *eoff, u64 cap, u32 l,
               const VertexTable &vt, u32 rank, u32 nranks, u64 *lkeys, u32 *lvals, u32 *loffs, u32 *ev1, u32 *ev2, u32 *lcount,
               u32 *ecount)
{
    dist_edges_kernel<<<grid_for(cap, DB), DB, 0, ctx->stream>>>(lt_keys, lt_cnt, base, eoff, cap, l, vt, rank, nranks, lkeys, lvals,
                                                                 loffs, ev1, ev2, lcount, ecount);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}
