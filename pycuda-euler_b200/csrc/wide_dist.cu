// wide_dist.cu -- 128-bit keys through the k-mer-space partition (BASELINE.json configs[4]: the k = 63
// leg of the k sweep on 8 GPUs).  Same ownership rule and the same per-rank graph contract as dist.cu
// (owner = scaled low bits of the minimizer score; an l-mer goes to the owners of its prefix and
// suffix vertex; edges are homed on the prefix owner), over two-word keys and with the
// correctness-first kernels of wide.cu: one thread walks one read.
#include "kernels.h"
#include "scan.cuh"
#include "wide.cuh"

#define WDB 128
#define WD_M 12   // == DIST_M of dist.cu

__device__ __forceinline__ u32 wd_owner(u32 score, u32 nranks) { return ((score & 0xffffu) * nranks) >> 16; }

// minimizer score of a len-mer (len >= WD_M): min over its m-mers of the scrambled canonical m-mer
__device__ __forceinline__ u32 wd_min_score(K128 x, u32 len)
{
    const u32 mmask = (1u << (2 * WD_M)) - 1u;
    const K128 r = revcomp128(x, len);
    u32 best = 0xffffffffu;
    for (u32 j = 0; j + WD_M <= len; j++) {
        const u32 w = (u32)shr128(x, 2 * (len - WD_M - j)).lo & mmask;
        const u32 rw = (u32)shr128(r, 2 * j).lo & mmask;
        const u32 sc = mmer_score(w < rw ? w : rw);
        best = sc < best ? sc : best;
    }
    return best;
}

// ---- sender: one thread per read, two passes (count per destination, then emit) -----------------
// The minimizer of the k-mer ending at t is a sliding minimum over the last WK = k - m + 1 m-mer
// scores (ring buffer, rescan only when the minimum leaves the window); the prefix k-mer of the
// l-mer ending at t is the k-mer ending at t - 1, so one sequence serves both ends.
template <bool EMIT>
__device__ __forceinline__ void wd_walk(const unsigned char *__restrict__ buf, u64 beg, u64 end, u32 l, u32 nranks,
                                        u32 (&cnt)[8], u64 (&pos)[8], K128 *__restrict__ send, const u64 *__restrict__ seg_off,
                                        u64 seg_cap, u64 &nl, u64 &nk)
{
    const u32 k = l - 1, top = 2 * (l - 1), WK = k - WD_M + 1;
    const u32 mmask = (1u << (2 * WD_M)) - 1u;
    const K128 lmask = mask128(l);
    K128 f = {0, 0}, rc = {0, 0};
    u32 mf = 0, mr = 0, run = 0;
    u32 ring[64];
    u32 minv = 0xffffffffu, prevmin = 0xffffffffu;
    u64 minpos = 0;
    for (u64 t = beg; t < end; t++) {
        const unsigned char c = buf[t];
        const unsigned char up = c & 0xDF;
        if (!(up == 'A' || up == 'C' || up == 'G' || up == 'T')) { run = 0; continue; }
        const u32 cc = ((c >> 1) ^ (c >> 2)) & 3u;
        f = and128(shl2_or(f, cc), lmask);
        rc = shr2_or_top(rc, 3u - cc, top);
        mf = ((mf << 2) | cc) & mmask;
        mr = (mr >> 2) | ((3u - cc) << (2 * (WD_M - 1)));
        run++;
        if (run >= WD_M) ring[t & 63] = mmer_score(mf < mr ? mf : mr);
        if (run < k) continue;
        if (!EMIT) nk++;
        const u32 s = ring[t & 63];
        if (run == k || minpos + WK <= t) {   // first k-mer of this run, or the minimum slid out: rescan
            minv = 0xffffffffu;
            for (u32 j = 0; j < WK; j++) {
                const u32 v = ring[(t - j) & 63];
                if (v < minv) { minv = v; minpos = t - j; }
            }
        } else if (s <= minv) {
            minv = s;
            minpos = t;
        }
        if (run >= l) {
            const u32 o1 = wd_owner(prevmin, nranks), o2 = wd_owner(minv, nranks);
            if (!EMIT) {
                nl++;
                cnt[o1]++;
                if (o2 != o1) cnt[o2]++;
            } else {
                const K128 key = lt128(f, rc) ? f : rc;
                const u64 a = pos[o1]++;
                if (a < seg_cap) send[seg_off[o1] + a] = key;
                if (o2 != o1) {
                    const u64 b = pos[o2]++;
                    if (b < seg_cap) send[seg_off[o2] + b] = key;
                }
            }
        }
        prevmin = minv;
    }
}

__global__ void __launch_bounds__(WDB) wide_dist_scatter_kernel(const unsigned char *__restrict__ buf, const u64 *__restrict__ off,
                                                                 u64 nreads, u32 l, u32 nranks, u64 *__restrict__ counts,
                                                                 u64 *__restrict__ cursors, K128 *__restrict__ send,
                                                                 const u64 *__restrict__ seg_off, u64 seg_cap)
{
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nreads) return;
    u32 cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    u64 pos[8];
    u64 nl = 0, nk = 0;
    const u64 beg = off[j], end = off[j + 1];
    wd_walk<false>(buf, beg, end, l, nranks, cnt, pos, send, seg_off, seg_cap, nl, nk);
    if (!nk) return;
    for (u32 d = 0; d < nranks; d++) pos[d] = cnt[d] ? atomicAdd((unsigned long long *)(cursors + d), (unsigned long long)cnt[d]) : 0ull;
    if (nl) atomicAdd((unsigned long long *)(counts + 16), (unsigned long long)nl);
    atomicAdd((unsigned long long *)(counts + 17), (unsigned long long)nk);
    if (nl) wd_walk<true>(buf, beg, end, l, nranks, cnt, pos, send, seg_off, seg_cap, nl, nk);
}

// seg_off in units of 16-byte keys from `send` (send == NULL: absolute addresses / 16, peer buffers)
int wide_dist_scatter(euler_ctx *ctx, const void *d_buf, const u64 *d_off, u64 nreads, u32 l, u32 nranks, u64 *d_counts,
                      u64 *d_cursors, void *d_send, const u64 *d_seg_off, u64 seg_cap)
{
    if (!nreads) return EULER_OK;
    if (nranks > 8) return euler_fail(ctx, EULER_ERR_ARG, "128-bit keys: at most 8 ranks");
    wide_dist_scatter_kernel<<<grid_for(nreads, WDB), WDB, 0, ctx->stream>>>((const unsigned char *)d_buf, d_off, nreads, l, nranks,
                                                                            d_counts, d_cursors, (K128 *)d_send, d_seg_off, seg_cap);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- receiver ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wide_count_keys_kernel(const K128 *__restrict__ in, u64 n, K128 *__restrict__ keys,
                                                               u32 *__restrict__ cnt, u64 cap, u64 *__restrict__ stats)
{
    // two keys per lane in flight; warp-uniform probing rounds as in dist_count_keys_kernel
    const u32 nb = (u32)(cap / WIDE_BUCKET);
    const u32 max_probe = nb < 8192 ? nb : 8192;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 t0 = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 iters = (n + stride * 2 - 1) / (stride * 2);
    u32 fresh = 0;
    bool overflow = false;
    for (u64 it = 0; it < iters; it++) {
        K128 key[2];
        u64 b[2];
        u32 pend = 0;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const u64 idx = (it * 2 + i) * stride + t0;
            key[i] = K128{0, 0};
            if (idx < n) {
                key[i] = in[idx];
                pend |= 1u << i;
            }
            b[i] = wide_hash_bucket(key[i], nb);
        }
        u32 probes = 0;
        while (__any_sync(0xffffffffu, pend != 0)) {
#pragma unroll
            for (int i = 0; i < 2; i++) {
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
            if (++probes >= max_probe && pend) { overflow = true; pend = 0; }
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) fresh += __shfl_xor_sync(0xffffffffu, fresh, o);
    if ((threadIdx.x & 31) == 0 && fresh) atomicAdd((unsigned long long *)(stats + 5), (unsigned long long)fresh);   // distinct keys held
    if (overflow) atomicOr((unsigned long long *)(stats + 2), 1ull);
}
int wide_count_keys(euler_ctx *ctx, const void *d_keys, u64 n, K128 *keys, u32 *cnt, u64 cap, u64 *d_stats)
{
    if (!n) return EULER_OK;
    u64 grid = (u64)ctx->num_sms * 8;
    const u64 need = (n + 511) / 512;
    if (grid > need) grid = need;
    wide_count_keys_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>((const K128 *)d_keys, n, keys, cnt, cap, d_stats);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// own[slot]: bit0 = prefix(c) is ours (strand c is homed here), bit1 = suffix(c) is ours (strand rc(c) is homed here)
__global__ void __launch_bounds__(256) wide_own_flags_kernel(const K128 *__restrict__ keys, u64 cap, u32 l, u32 rank, u32 nranks,
                                                              unsigned char *__restrict__ own)
{
    const u64 slot = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= cap) return;
    const K128 c = keys[slot];
    if (is_empty128(c)) { own[slot] = 0; return; }
    const u32 k = l - 1;
    const u32 op = wd_owner(wd_min_score(shr128(c, 2), k), nranks);
    const u32 os = wd_owner(wd_min_score(and128(c, mask128(k)), k), nranks);
    own[slot] = (unsigned char)((op == rank ? 1u : 0u) | (os == rank ? 2u : 0u));
}
int wide_own_flags(euler_ctx *ctx, const K128 *keys, u64 cap, u32 l, u32 rank, u32 nranks, unsigned char *own)
{
    wide_own_flags_kernel<<<grid_for(cap, 256), 256, 0, ctx->stream>>>(keys, cap, l, rank, nranks, own);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

__global__ void __launch_bounds__(256) wide_dist_vertex_insert_kernel(const K128 *__restrict__ lt_keys, u64 lt_cap, u32 l,
                                                                       const unsigned char *__restrict__ own,
                                                                       K128 *__restrict__ vt_keys, u64 vt_cap, u64 *flags)
{
    const u64 slot = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= lt_cap) return;
    const u32 o = own[slot];
    if (!o) return;
    const K128 c = lt_keys[slot];
    const u32 k = l - 1;
    const u32 nb = (u32)(vt_cap / WIDE_BUCKET);
    const u32 max_probe = nb < 8192 ? nb : 8192;
    bool bad = false;
    if (o & 1u) bad |= wide_insert(vt_keys, vt_cap, canon128(shr128(c, 2), k), max_probe) == EULER_NO_SLOT;
    if (o & 2u) bad |= wide_insert(vt_keys, vt_cap, canon128(and128(c, mask128(k)), k), max_probe) == EULER_NO_SLOT;
    if (bad) atomicOr((unsigned long long *)flags, 2ull);
}
int wide_dist_vertex_insert(euler_ctx *ctx, const K128 *lt_keys, u64 lt_cap, u32 l, const unsigned char *own, K128 *vt_keys,
                            u64 vt_cap, u64 *d_flags)
{
    wide_dist_vertex_insert_kernel<<<grid_for(lt_cap, 256), 256, 0, ctx->stream>>>(lt_keys, lt_cap, l, own, vt_keys, vt_cap, d_flags);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// homed strands per slot: c if bit0, rc(c) if bit1 (a palindrome is one l-mer, both bits equal)
struct WideHomedWeight {
    const K128 *keys;
    const unsigned char *own;
    u32 l;
    __device__ __forceinline__ u32 operator()(u64 i) const
    {
        const u32 o = own[i];
        if (!o) return 0u;
        const K128 x = keys[i];
        if (eq128(x, revcomp128(x, l))) return (o & 1u) ? 1u : 0u;
        return (o & 1u) + ((o >> 1) & 1u);
    }
};
int wide_homed_scan(euler_ctx *ctx, const K128 *keys, const unsigned char *own, u64 cap, u32 l, u32 *d_base, u64 *d_total)
{
    return scan_exclusive(ctx, WideHomedWeight{keys, own, l}, cap, d_base, d_total);
}

__global__ void __launch_bounds__(256) wide_compact_homed_kernel(const K128 *__restrict__ lt_keys, const u32 *__restrict__ lt_cnt,
                                                                  const unsigned char *__restrict__ own,
                                                                  const u32 *__restrict__ base, u64 cap, u32 l,
                                                                  u64 *__restrict__ lk_lo, u64 *__restrict__ lk_hi,
                                                                  u32 *__restrict__ lvals)
{
    const u64 slot = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= cap) return;
    const u32 o = own[slot];
    if (!o) return;
    const K128 c = lt_keys[slot];
    const u32 n = lt_cnt[slot];
    u32 idx = base[slot];
    const K128 r = revcomp128(c, l);
    if (eq128(c, r)) {
        if (o & 1u) { lk_lo[idx] = c.lo; lk_hi[idx] = c.hi; lvals[idx] = 2u * n; }
        return;
    }
    if (o & 1u) { lk_lo[idx] = c.lo; lk_hi[idx] = c.hi; lvals[idx] = n; idx++; }
    if (o & 2u) { lk_lo[idx] = r.lo; lk_hi[idx] = r.hi; lvals[idx] = n; }
}
int wide_compact_homed(euler_ctx *ctx, const K128 *lt_keys, const u32 *lt_cnt, const unsigned char *own, const u32 *base, u64 cap,
                       u32 l, u64 *lk_lo, u64 *lk_hi, u32 *lvals)
{
    wide_compact_homed_kernel<<<grid_for(cap, 256), 256, 0, ctx->stream>>>(lt_keys, lt_cnt, own, base, cap, l, lk_lo, lk_hi, lvals);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// in-edges of owned vertices whose l-mer is homed elsewhere (prefix on another rank): ecount only
__global__ void __launch_bounds__(256) wide_foreign_in_edges_kernel(const K128 *__restrict__ lt_keys, const u32 *__restrict__ lt_cnt,
                                                                     const unsigned char *__restrict__ own, u64 cap, u32 l,
                                                                     const K128 *__restrict__ vt_keys, const u32 *__restrict__ id0,
                                                                     u64 vt_cap, u32 *__restrict__ ecount)
{
    const u64 slot = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= cap) return;
    const u32 o = own[slot];
    if (o != 1u && o != 2u) return;
    const K128 c = lt_keys[slot];
    const u32 n = lt_cnt[slot], k = l - 1;
    // o == 2: strand c has a foreign prefix and our suffix; o == 1: strand rc(c) does
    const K128 x = (o == 2u) ? c : revcomp128(c, l);
    const K128 s = and128(x, mask128(k));
    const K128 rs = revcomp128(s, k);
    const bool s_can = !lt128(rs, s);
    const u64 vs = wide_find(vt_keys, vt_cap, s_can ? s : rs);
    if (vs == EULER_NO_SLOT) return;
    const u32 sid = s_can ? id0[vs] : id0[vs] + 1u;
    ecount[((u64)sid << 2) + base_at(x, l, 0)] = n;
}
int wide_foreign_in_edges(euler_ctx *ctx, const K128 *lt_keys, const u32 *lt_cnt, const unsigned char *own, u64 cap, u32 l,
                          const K128 *vt_keys, const u32 *id0, u64 vt_cap, u32 *ecount)
{
    wide_foreign_in_edges_kernel<<<grid_for(cap, 256), 256, 0, ctx->stream>>>(lt_keys, lt_cnt, own, cap, l, vt_keys, id0, vt_cap,
                                                                             ecount);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}
