// wide.cu -- 128-bit keys: l-mers of 33..64 bases (k = 32..63), BASELINE.json configs[4] (k = 63).
// The reference stops at 64-bit KEY_T (pyencode.py:22-33, l <= 32); this path extends the same
// pipeline (canonical l-mer table -> vertex table -> D1-D6) to two-word keys.  It is the
// correctness-first variant: one thread walks one read with rolling 128-bit forward /
// reverse-complement registers; the table is open addressing over 32-byte buckets of two 16-byte
// keys, claimed with a single 128-bit compare-and-swap (atom.cas.b128, sm_90+).  Everything after
// the tables (scans, EulerVertex, expanded edges, tour, contigs) is shared with the 64-bit path.
#include "kernels.h"
#include "scan.cuh"
#include "sort.cuh"
#include "tmp.cuh"
#include "wide.cuh"
#include "encode.cuh"

#define WB 256

// ---- count: one thread per read ------------------------------------------------------------------
__global__ void __launch_bounds__(WB) wide_count_kernel(const unsigned char *__restrict__ buf, const u64 *__restrict__ off,
                                                         u64 nreads, u32 l, K128 *__restrict__ keys, u32 *__restrict__ cnt, u64 cap,
                                                         u64 *__restrict__ stats)
{
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nreads) return;
    const u32 k = l - 1, top = 2 * (l - 1);
    const K128 lmask = mask128(l);
    const u32 nb = (u32)(cap / WIDE_BUCKET);
    const u32 max_probe = nb < 8192 ? nb : 8192;
    K128 f = {0, 0}, rc = {0, 0};
    u32 run = 0;
    u64 nl = 0, nk = 0;
    bool overflow = false;
    for (u64 t = off[j]; t < off[j + 1]; t++) {
        const unsigned char c = buf[t];
        const unsigned char up = c & 0xDF;
        const bool ok = up == 'A' || up == 'C' || up == 'G' || up == 'T';
        if (!ok) { run = 0; continue; }
        const u32 cc = ((c >> 1) ^ (c >> 2)) & 3u;   // A0 C1 G2 T3 (pyencode.py:42 codeF)
        f = and128(shl2_or(f, cc), lmask);
        rc = shr2_or_top(rc, 3u - cc, top);
        run++;
        if (run >= k) nk++;
        if (run >= l) {
            nl++;
            const K128 key = lt128(f, rc) ? f : rc;
            const u64 slot = wide_insert(keys, cap, key, max_probe);
            if (slot == EULER_NO_SLOT) overflow = true;
            else atomicAdd(cnt + slot, 1u);
        }
    }
    if (nl) atomicAdd(stats + 0, nl);
    if (nk) atomicAdd(stats + 1, nk);
    if (overflow) atomicOr((unsigned long long *)(stats + 2), 1ull);
}

int wide_count(euler_ctx *ctx, const void *d_buf, const u64 *d_off, u64 nreads, u32 l, K128 *keys, u32 *cnt, u64 cap, u64 *d_stats)
{
    if (!nreads) return EULER_OK;
    wide_count_kernel<<<grid_for(nreads, WB), WB, 0, ctx->stream>>>((const unsigned char *)d_buf, d_off, nreads, l, keys, cnt, cap,
                                                                   d_stats);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- count, tiled: the 64-bit count_compact_kernel (encode.cu) over two-word keys ---------------------------
// A warp covers 32 chunks of 16 bytes; the first WT_HALO = 4 lanes only supply the 63 bases of left context, the
// other 28 own the windows that end in their chunk.  Phase 1 rolls the 128-bit forward / reverse-complement
// registers through the 16 positions in lock step and packs the valid canonical keys into shared memory; phase
// 2 probes them two per lane: one step per key first, then the keys whose home bucket was full together.
#define WT_HALO 4
#define WT_ADV (32 - WT_HALO)
#define WT_BLOCK 128
#define WT_KEYS (WT_ADV * 16)
#define WT_KPL 2
__global__ void __launch_bounds__(WT_BLOCK, 4) wide_count_tiled_kernel(const uint4 *__restrict__ buf16, u64 n_bases,
                                                                       const u32 *__restrict__ start_bits, u32 l,
                                                                       K128 *__restrict__ keys, u32 *__restrict__ cnt, u64 cap,
                                                                       u64 ntiles, u64 *__restrict__ stats)
{
    __shared__ K128 s_keys[(WT_BLOCK / 32) * WT_KEYS];
    K128 *stage = s_keys + (threadIdx.x >> 5) * WT_KEYS;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const u32 nb = (u32)(cap / WIDE_BUCKET);
    const u32 max_probe = nb < 8192 ? nb : 8192;
    const u32 k = l - 1, top = 2 * (l - 1);
    const K128 lmask = mask128(l);
    u32 nl_tot = 0, nk_tot = 0, fresh = 0;
    bool overflow = false;

    for (u64 tile = warp; tile < ntiles; tile += nwarps) {
        const long long chunk = (long long)(tile * WT_ADV) - WT_HALO + lane;
        const Chunk c = load_chunk(buf16, chunk, n_bases, start_bits);
        const u32 p1 = __shfl_up_sync(0xffffffffu, c.codes, 1), p2 = __shfl_up_sync(0xffffffffu, c.codes, 2);
        const u32 p3 = __shfl_up_sync(0xffffffffu, c.codes, 3), p4 = __shfl_up_sync(0xffffffffu, c.codes, 4);
        const u32 v1 = __shfl_up_sync(0xffffffffu, c.vmask, 1), v2 = __shfl_up_sync(0xffffffffu, c.vmask, 2);
        const u32 v3 = __shfl_up_sync(0xffffffffu, c.vmask, 3), v4 = __shfl_up_sync(0xffffffffu, c.vmask, 4);
        const u32 s1 = __shfl_up_sync(0xffffffffu, c.smask, 1), s2 = __shfl_up_sync(0xffffffffu, c.smask, 2);
        const u32 s3 = __shfl_up_sync(0xffffffffu, c.smask, 3), s4 = __shfl_up_sync(0xffffffffu, c.smask, 4);
        // the 64 bases before this chunk (p1 = the chunk just before), their validity / read starts (bit 0 = most recent)
        K128 f{((u64)p2 << 32) | p1, ((u64)p4 << 32) | p3};
        K128 rc = revcomp128(and128(f, lmask), l);
        const u64 pv = ((u64)v4 << 48) | ((u64)v3 << 32) | ((u64)v2 << 16) | v1;
        const u64 ps = ((u64)s4 << 48) | ((u64)s3 << 32) | ((u64)s2 << 16) | s1;
        u32 vrun = (lane < WT_HALO) ? 0u : ((pv == ~0ull) ? 64u : (u32)__ffsll((long long)~pv) - 1u);
        u32 srun = ps ? (u32)__ffsll((long long)ps) - 1u : 64u;
        u32 codes = c.codes;
        u32 vm = (lane < WT_HALO) ? 0u : (c.vmask << 16);   // halo lanes own no windows
        u32 sm = c.smask << 16;
        u32 nvalid = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const u32 cc = codes >> 30;
            codes <<= 2;
            const bool valid = (vm >> 31) != 0, start = (sm >> 31) != 0;
            vm <<= 1;
            sm <<= 1;
            f = and128(shl2_or(f, cc), lmask);
            rc = shr2_or_top(rc, 3u - cc, top);
            vrun = valid ? vrun + 1u : 0u;
            srun = start ? 0u : srun + 1u;
            nk_tot += (vrun >= k && srun + 1u >= k) ? 1u : 0u;
            const bool ok = vrun >= l && srun + 1u >= l;
            const unsigned okm = __ballot_sync(0xffffffffu, ok);
            if (ok) stage[nvalid + __popc(okm & lt_mask)] = lt128(f, rc) ? f : rc;
            nvalid += __popc(okm);
        }
        nl_tot += nvalid;   // warp-uniform
        __syncwarp();
        u32 ncur = nvalid;
#pragma unroll 1
        for (int pass = 0; pass < 2 && ncur; pass++) {
            u32 nleft = 0;
#pragma unroll 1
            for (u32 g = 0; g < ncur; g += 32 * WT_KPL) {
                K128 key[WT_KPL];
                u64 b[WT_KPL];
                u32 pend = 0;
#pragma unroll
                for (int i = 0; i < WT_KPL; i++) {
                    const u32 idx = g + i * 32 + lane;
                    key[i] = K128{0, 0};
                    if (idx < ncur) {
                        key[i] = stage[idx];
                        pend |= 1u << i;
                    }
                    b[i] = wide_hash_bucket(key[i], nb);
                    if (pass && ++b[i] == nb) b[i] = 0;   // stragglers: the home bucket was full
                }
                __syncwarp();
                u32 probes = 0;
                while (__any_sync(0xffffffffu, pend != 0)) {
#pragma unroll
                    for (int i = 0; i < WT_KPL; i++) {
                        if (pend & (1u << i)) {
                            const u64 slot = wide_probe_step(keys, b[i], key[i], fresh);
                            if (slot != EULER_NO_SLOT) {
                                atomicAdd(cnt + slot, 1u);
                                pend &= ~(1u << i);
                            } else if (++b[i] == nb) {
                                b[i] = 0;
                            }
                        }
                    }
                    if (pass == 0) break;
                    if (++probes >= max_probe && pend) { overflow = true; pend = 0; }
                }
                if (pass == 0) {
#pragma unroll
                    for (int i = 0; i < WT_KPL; i++) {
                        const bool left = (pend >> i) & 1u;
                        const unsigned lm = __ballot_sync(0xffffffffu, left);
                        if (left) stage[nleft + __popc(lm & lt_mask)] = key[i];
                        nleft += __popc(lm);
                    }
                }
            }
            __syncwarp();
            ncur = nleft;
        }
        __syncwarp();
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        nk_tot += __shfl_xor_sync(0xffffffffu, nk_tot, d);
        fresh += __shfl_xor_sync(0xffffffffu, fresh, d);
    }
    if (lane == 0) {
        if (nl_tot) atomicAdd(stats + 0, (u64)nl_tot);
        if (nk_tot) atomicAdd(stats + 1, (u64)nk_tot);
        if (fresh) atomicAdd(stats + 5, (u64)fresh);
    }
    if (overflow) atomicOr((unsigned long long *)(stats + 2), 1ull);
}

int wide_count_tiled(euler_ctx *ctx, const void *d_buf, u64 n_bases, const u32 *d_bits, u32 l, K128 *keys, u32 *cnt, u64 cap,
                     u64 *d_stats)
{
    if (!n_bases) return EULER_OK;
    const u64 nchunks = (n_bases + 15) / 16;
    const u64 ntiles = (nchunks + WT_ADV - 1) / WT_ADV;
    u64 grid = (u64)ctx->num_sms * 4;
    const u64 need = (ntiles + WT_BLOCK / 32 - 1) / (WT_BLOCK / 32);
    if (grid > need) grid = need;
    wide_count_tiled_kernel<<<(unsigned)grid, WT_BLOCK, 0, ctx->stream>>>((const uint4 *)d_buf, n_bases, d_bits, l, keys, cnt, cap,
                                                                           ntiles, d_stats);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

int wide_table_clear(euler_ctx *ctx, K128 *keys, u32 *vals, u64 cap)
{
    CUDA_TRY(ctx, cudaMemsetAsync(keys, 0xFF, cap * sizeof(K128), ctx->stream));
    if (vals) CUDA_TRY(ctx, cudaMemsetAsync(vals, 0, cap * sizeof(u32), ctx->stream));
    return EULER_OK;
}

// ---- vertex table ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(WB) wide_vertex_insert_kernel(const K128 *__restrict__ lt_keys, u64 lt_cap, u32 l,
                                                                 K128 *__restrict__ vt_keys, u64 vt_cap, u64 *flags)
{
    const u64 slot = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= lt_cap) return;
    const K128 c = lt_keys[slot];
    if (is_empty128(c)) return;
    const u32 k = l - 1;
    const u32 nb = (u32)(vt_cap / WIDE_BUCKET);
    const u32 max_probe = nb < 8192 ? nb : 8192;
    const K128 p = shr128(c, 2), s = and128(c, mask128(k));
    const u64 a = wide_insert(vt_keys, vt_cap, canon128(p, k), max_probe);
    const u64 b = wide_insert(vt_keys, vt_cap, canon128(s, k), max_probe);
    if (a == EULER_NO_SLOT || b == EULER_NO_SLOT) atomicOr((unsigned long long *)flags, 2ull);
}

int wide_vertex_insert(euler_ctx *ctx, const K128 *lt_keys, u64 lt_cap, u32 l, K128 *vt_keys, u64 vt_cap, u64 *d_flags)
{
    wide_vertex_insert_kernel<<<grid_for(lt_cap, WB), WB, 0, ctx->stream>>>(lt_keys, lt_cap, l, vt_keys, vt_cap, d_flags);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

struct WideSlotWeight {
    const K128 *keys;
    u32 len;
    __device__ __forceinline__ u32 operator()(u64 i) const
    {
        const K128 x = keys[i];
        if (is_empty128(x)) return 0u;
        return eq128(x, revcomp128(x, len)) ? 1u : 2u;
    }
};
int wide_slot_scan(euler_ctx *ctx, const K128 *keys, u64 cap, u32 len, u32 *d_base, u64 *d_total)
{
    return scan_exclusive(ctx, WideSlotWeight{keys, len}, cap, d_base, d_total);
}

__global__ void __launch_bounds__(WB) wide_compact_vertices_kernel(const K128 *__restrict__ vt_keys, const u32 *__restrict__ base,
                                                                    u64 cap, u32 k, u64 *__restrict__ vk_lo, u64 *__restrict__ vk_hi)
{
    const u64 slot = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= cap) return;
    const K128 c = vt_keys[slot];
    if (is_empty128(c)) return;
    const u32 idx = base[slot];
    const K128 r = revcomp128(c, k);
    vk_lo[idx] = c.lo; vk_hi[idx] = c.hi;
    if (!eq128(c, r)) { vk_lo[idx + 1] = r.lo; vk_hi[idx + 1] = r.hi; }
}
int wide_compact_vertices(euler_ctx *ctx, const K128 *vt_keys, const u32 *vt_base, u64 vt_cap, u32 k, u64 *vk_lo, u64 *vk_hi)
{
    wide_compact_vertices_kernel<<<grid_for(vt_cap, WB), WB, 0, ctx->stream>>>(vt_keys, vt_base, vt_cap, k, vk_lo, vk_hi);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// both-strand (key, multiplicity) pairs in slot order (readLmersKmersCuda eulercuda.py:141-178, palindrome 2c)
__global__ void __launch_bounds__(WB) wide_compact_lmers_kernel(const K128 *__restrict__ lt_keys, const u32 *__restrict__ lt_cnt,
                                                                 const u32 *__restrict__ base, u64 cap, u32 l,
                                                                 u64 *__restrict__ lk_lo, u64 *__restrict__ lk_hi,
                                                                 u32 *__restrict__ lvals)
{
    const u64 slot = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= cap) return;
    const K128 c = lt_keys[slot];
    if (is_empty128(c)) return;
    const u32 n = lt_cnt[slot], idx = base[slot];
    const K128 r = revcomp128(c, l);
    const bool pal = eq128(c, r);
    lk_lo[idx] = c.lo; lk_hi[idx] = c.hi; lvals[idx] = pal ? 2u * n : n;
    if (!pal) { lk_lo[idx + 1] = r.lo; lk_hi[idx + 1] = r.hi; lvals[idx + 1] = n; }
}
int wide_compact_lmers(euler_ctx *ctx, const K128 *lt_keys, const u32 *lt_cnt, const u32 *lt_base, u64 lt_cap, u32 l, u64 *lk_lo,
                       u64 *lk_hi, u32 *lvals)
{
    wide_compact_lmers_kernel<<<grid_for(lt_cap, WB), WB, 0, ctx->stream>>>(lt_keys, lt_cnt, lt_base, lt_cap, l, lk_lo, lk_hi, lvals);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- ascending 128-bit order (canonical ids, B14): two stable LSD sorts of a permutation -------------
__global__ void __launch_bounds__(WB) wide_iota_kernel(u32 *p, u64 n)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (u32)i;
}
__global__ void __launch_bounds__(WB) wide_gather_u64_kernel(const u64 *__restrict__ src, const u32 *__restrict__ perm, u64 n,
                                                              u64 *__restrict__ dst)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[perm[i]];
}
__global__ void __launch_bounds__(WB) wide_gather_u32_kernel(const u32 *__restrict__ src, const u32 *__restrict__ perm, u64 n,
                                                              u32 *__restrict__ dst)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[perm[i]];
}
int wide_sort(euler_ctx *ctx, u64 *lo, u64 *hi, u32 *vals, u64 n, int nbits)
{
    if (n < 2) return EULER_OK;
    const u32 nblocks = (u32)((n + RS_TILE - 1) / RS_TILE);
    DevTmp<u64> k1(ctx, n), k2(ctx, n), kt(ctx, n);
    DevTmp<u32> perm(ctx, n), pt(ctx, n), hist(ctx, (u64)256 * nblocks);
    TMP_CHECK(ctx, k1); TMP_CHECK(ctx, k2); TMP_CHECK(ctx, kt); TMP_CHECK(ctx, perm); TMP_CHECK(ctx, pt); TMP_CHECK(ctx, hist);
    const unsigned g = grid_for(n, WB);
    wide_iota_kernel<<<g, WB, 0, ctx->stream>>>(perm, n);
    CUDA_TRY(ctx, cudaMemcpyAsync(k1.get(), lo, n * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
    EULER_TRY(radix_sort_pairs(ctx, k1.get(), perm.get(), n, nbits < 64 ? nbits : 64, kt.get(), pt.get(), hist.get()));
    if (nbits > 64) {
        wide_gather_u64_kernel<<<g, WB, 0, ctx->stream>>>(hi, perm, n, k2);
        EULER_TRY(radix_sort_pairs(ctx, k2.get(), perm.get(), n, nbits - 64, kt.get(), pt.get(), hist.get()));
    }
    wide_gather_u64_kernel<<<g, WB, 0, ctx->stream>>>(lo, perm, n, k1);
    wide_gather_u64_kernel<<<g, WB, 0, ctx->stream>>>(hi, perm, n, k2);
    CUDA_TRY(ctx, cudaMemcpyAsync(lo, k1.get(), n * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(hi, k2.get(), n * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
    if (vals) {
        wide_gather_u32_kernel<<<g, WB, 0, ctx->stream>>>(vals, perm, n, pt);
        CUDA_TRY(ctx, cudaMemcpyAsync(vals, pt.get(), n * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

__global__ void __launch_bounds__(WB) wide_assign_sorted_ids_kernel(const u64 *__restrict__ vk_lo, const u64 *__restrict__ vk_hi,
                                                                     u64 nv, const K128 *__restrict__ vt_keys, u64 vt_cap, u32 k,
                                                                     u32 *__restrict__ id0, u32 *__restrict__ id1)
{
    const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nv) return;
    const K128 x{vk_lo[r], vk_hi[r]};
    const K128 rc = revcomp128(x, k);
    const bool is_c = !lt128(rc, x);
    const u64 slot = wide_find(vt_keys, vt_cap, is_c ? x : rc);
    if (slot == EULER_NO_SLOT) return;
    if (is_c) id0[slot] = (u32)r;
    if (eq128(x, rc) || !is_c) id1[slot] = (u32)r;
}
int wide_assign_sorted_ids(euler_ctx *ctx, const u64 *vk_lo, const u64 *vk_hi, u64 nv, const K128 *vt_keys, u64 vt_cap, u32 k,
                           u32 *id0, u32 *id1)
{
    if (!nv) return EULER_OK;
    wide_assign_sorted_ids_kernel<<<grid_for(nv, WB), WB, 0, ctx->stream>>>(vk_lo, vk_hi, nv, vt_keys, vt_cap, k, id0, id1);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- D1 debruijnCount (pydebruijn.py:107-141) + compressed edges over explicit both-strand l-mers -----
__device__ __forceinline__ u32 wide_vt_lookup(const K128 *vt_keys, const u32 *id0, const u32 *id1, u64 cap, u32 k, K128 v)
{
    const K128 r = revcomp128(v, k);
    const bool is_c = !lt128(r, v);
    const u64 slot = wide_find(vt_keys, cap, is_c ? v : r);
    if (slot == EULER_NO_SLOT) return EULER_NO_ID;
    if (is_c) return id0[slot];
    return id1 ? id1[slot] : id0[slot] + 1u;
}
__global__ void __launch_bounds__(WB) wide_degree_slots_kernel(const u64 *__restrict__ lk_lo, const u64 *__restrict__ lk_hi,
                                                                const u32 *__restrict__ lvals, u64 nl, u32 l,
                                                                const K128 *__restrict__ vt_keys, const u32 *__restrict__ id0,
                                                                const u32 *__restrict__ id1, u64 vt_cap, u32 *__restrict__ lcount,
                                                                u32 *__restrict__ ecount, u32 *__restrict__ ev1,
                                                                u32 *__restrict__ ev2, unsigned char *__restrict__ tf)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nl) return;
    const K128 x{lk_lo[i], lk_hi[i]};
    const u32 m = lvals[i], k = l - 1;
    const u32 pid = wide_vt_lookup(vt_keys, id0, id1, vt_cap, k, shr128(x, 2));
    const u32 sid = wide_vt_lookup(vt_keys, id0, id1, vt_cap, k, and128(x, mask128(k)));
    const u32 to = (u32)(x.lo & 3), from = base_at(x, l, 0);
    if (pid != EULER_NO_ID) lcount[((u64)pid << 2) + to] = m;
    if (sid != EULER_NO_ID) ecount[((u64)sid << 2) + from] = m;
    ev1[i] = pid; ev2[i] = sid;
    tf[i] = (unsigned char)(to | (from << 2));
}
int wide_degree_slots(euler_ctx *ctx, const u64 *lk_lo, const u64 *lk_hi, const u32 *lvals, u64 nl, u32 l, const K128 *vt_keys,
                      const u32 *id0, const u32 *id1, u64 vt_cap, u32 *lcount, u32 *ecount, u32 *ev1, u32 *ev2, unsigned char *tf)
{
    if (!nl) return EULER_OK;
    wide_degree_slots_kernel<<<grid_for(nl, WB), WB, 0, ctx->stream>>>(lk_lo, lk_hi, lvals, nl, l, vt_keys, id0, id1, vt_cap, lcount,
                                                                       ecount, ev1, ev2, tf);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}
