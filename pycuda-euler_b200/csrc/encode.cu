// encode.cu -- encoder kernels and the fused encode+count kernel (the hot kernel of the path).
#include "encode.cuh"
#include "kernels.h"
#include <stdlib.h>

// ---- read-start bitmap: bit t set iff a read starts at byte t ---------------------------------
__global__ void mark_starts_kernel(const u64 *__restrict__ off, u64 nreads, u32 *__restrict__ bits)
{
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nreads) return;
    const u64 t = off[j];
    atomicOr(bits + (t >> 5), 1u << (t & 31));
}

int enc_mark_starts(euler_ctx *ctx, const u64 *d_off, u64 nreads, u64 n_bases, u32 *d_bits)
{
    const u64 words = n_bases / 32 + 2;
    CUDA_TRY(ctx, cudaMemsetAsync(d_bits, 0, words * sizeof(u32), ctx->stream));
    if (nreads) {
        mark_starts_kernel<<<grid_for(nreads, 256), 256, 0, ctx->stream>>>(d_off, nreads, d_bits);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    return EULER_OK;
}

// ---- module-level encoder: one value per window start byte ------------------------------------
__global__ void __launch_bounds__(256) encode_positions_kernel(const uint4 *__restrict__ buf16, u64 n_bases,
                                                                const u32 *__restrict__ start_bits, u32 l,
                                                                u64 *__restrict__ out_fwd, u64 *__restrict__ out_rc,
                                                                unsigned char *__restrict__ out_valid, u64 ntiles)
{
    const int lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 tile = warp; tile < ntiles; tile += nwarps) {
        const long long chunk = (long long)(tile * ENC_ADV) - ENC_HALO + lane;
        const Chunk c = load_chunk(buf16, chunk, n_bases, start_bits);
        for_each_window(c, l, lane, [&](int i, u64 key) {
            const u64 start = (u64)chunk * 16 + i + 1 - l;
            out_fwd[start] = key;
            if (out_rc) out_rc[start] = revcomp64(key, l);
            if (out_valid) out_valid[start] = 1;
        });
    }
}

int enc_positions(euler_ctx *ctx, const void *d_buf, u64 n_bases, const u32 *d_bits, u32 l, u64 *d_fwd, u64 *d_rc,
                  unsigned char *d_valid)
{
    CUDA_TRY(ctx, cudaMemsetAsync(d_fwd, 0, n_bases * sizeof(u64), ctx->stream));
    if (d_rc) CUDA_TRY(ctx, cudaMemsetAsync(d_rc, 0, n_bases * sizeof(u64), ctx->stream));
    if (d_valid) CUDA_TRY(ctx, cudaMemsetAsync(d_valid, 0, n_bases, ctx->stream));
    if (!n_bases) return EULER_OK;
    const u64 nchunks = (n_bases + 15) / 16;
    const u64 ntiles = (nchunks + ENC_ADV - 1) / ENC_ADV;
    const unsigned grid = (unsigned)min((u64)ctx->num_sms * 8, (ntiles + 7) / 8);
    encode_positions_kernel<<<grid ? grid : 1, 256, 0, ctx->stream>>>((const uint4 *)d_buf, n_bases, d_bits, l, d_fwd,
                                                                      d_rc, d_valid, ntiles);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

__global__ void compute_kmers_kernel(const u64 *__restrict__ lmers, u64 n, u64 mask, u64 *__restrict__ pk,
                                     u64 *__restrict__ sk)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 x = lmers[i];
    pk[i] = (x & (mask << 2)) >> 2;  // LMER_PREFIX pyencode.py:109
    sk[i] = x & mask;                // LMER_SUFFIX pyencode.py:110
}

int enc_compute_kmers(euler_ctx *ctx, const u64 *d_lmers, u64 n, u64 mask, u64 *d_pk, u64 *d_sk)
{
    if (!n) return EULER_OK;
    compute_kmers_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(d_lmers, n, mask, d_pk, d_sk);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- fused encode + canonicalise + count (hot kernel) -----------------------------------------
// Every forward window contributes +1 to its canonical key min(x, rc(x)); the both-strand table of
// the reference (eulercuda.py:141-161) is count[x] = count[rc x] = c, or 2c for palindromes.
// Table: SoA keys u64[cap] / counts u32[cap], linear probing, EMPTY = all-ones (never canonical).
#define CNT_BLOCK 256

// Sliding-window minimum of the m-mer scores: s[t], t = position + NPREV, holds the scores of the
// W-1 positions before the chunk followed by the chunk's 16.  win[i] = min s over the W m-mers of
// the l-mer ending at chunk position i.  Log-doubling min tree, all indices compile-time.
template <int W>
struct WinMin {
    static constexpr int NPREV = W - 1;
    static constexpr int N = NPREV + 16;
    static constexpr int E = (W >= 16) ? 4 : (W >= 8) ? 3 : (W >= 4) ? 2 : (W >= 2) ? 1 : 0;
    static constexpr int P = 1 << E;
    __device__ __forceinline__ static void run(u32 (&s)[N], u32 (&win)[16])
    {
        // after step e, s[t] = min of the original s[t .. t + 2^e - 1]
#pragma unroll
        for (int e = 0; e < E; e++) {
            const int h = 1 << e;
#pragma unroll
            for (int t = 0; t + h < N; t++) s[t] = s[t] < s[t + h] ? s[t] : s[t + h];
        }
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int lo = i + NPREV - (W - 1);  // first m-mer of the window
            const int hi = i + NPREV - (P - 1);  // start of the last P-block inside the window
            win[i] = s[lo] < s[hi] ? s[lo] : s[hi];
        }
    }
};

// Rolling formulation: each lane walks the 16 bases of its chunk, keeping the forward and
// reverse-complement l-mers in two 64-bit registers (shift in / shift out), plus two run lengths
// (valid bases, bases since the last read start) that decide whether the window ending here is a
// whole l-mer / k-mer of one read.  Keys are produced four at a time and probed right away, so the
// live state is ~60 registers and the loop body stays inside the instruction cache.
// W > 0: minimizer-ordered table (W = l - m + 1 m-mers per l-mer); W == 0: plain hash;
// W == -1: minimizer-ordered with a per-window brute-force minimum (any l);
// W == -2: plain hash of the canonical PREFIX k-mer (co-hashed with the vertex table, common.cuh).
template <int W>
__global__ void __launch_bounds__(CNT_BLOCK, 4) count_canonical_kernel(const uint4 *__restrict__ buf16, u64 n_bases,
                                                                        const u32 *__restrict__ start_bits, u32 l,
                                                                        u64 *__restrict__ tab_keys, u32 *__restrict__ tab_cnt,
                                                                        u64 cap, TableHash th, u64 ntiles, u64 *__restrict__ stats,
                                                                        u64 part_lo, u64 part_hi, int count_windows)
{
    __shared__ u32 s_win[W > 0 ? 16 * CNT_BLOCK : 1];
    const int lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const u32 nbuckets = (u32)(cap / EULER_BUCKET);
    const u32 max_probe = nbuckets < 4096 ? nbuckets : 4096;
    const u32 k = l - 1;
    const u32 top = 2 * (l - 1);
    const u64 kmask = key_mask_d(l);
    u32 nl_tot = 0, nk_tot = 0;
    bool overflow = false;

    for (u64 tile = warp; tile < ntiles; tile += nwarps) {
        const long long chunk = (long long)(tile * ENC_ADV) - ENC_HALO + lane;
        const Chunk c = load_chunk(buf16, chunk, n_bases, start_bits);
        const u32 p1 = __shfl_up_sync(0xffffffffu, c.codes, 1), p2 = __shfl_up_sync(0xffffffffu, c.codes, 2);
        const u32 v1 = __shfl_up_sync(0xffffffffu, c.vmask, 1), v2 = __shfl_up_sync(0xffffffffu, c.vmask, 2);
        const u32 s1 = __shfl_up_sync(0xffffffffu, c.smask, 1), s2 = __shfl_up_sync(0xffffffffu, c.smask, 2);
        // state after the base just before this chunk
        u64 f = ((u64)p2 << 32) | p1;
        u64 rc = revcomp64(f & kmask, l);

        if constexpr (W > 0) {
            // minimizer of every l-mer ending in this chunk: scores of the chunk's 16 m-mers, the
            // W-1 before them from the two lanes below, sliding minimum, parked in shared memory
            constexpr int NP = W - 1;
            const u32 m = th.m;
            const u32 mmask = m >= 16 ? 0xffffffffu : ((1u << (2 * m)) - 1u);
            const u32 rsh = 2 * (l - m);
            u32 sc[16];
            {
                u64 f2 = f, r2 = rc;
                u32 cd = c.codes;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const u32 cc = cd >> 30;
                    cd <<= 2;
                    f2 = (f2 << 2) | cc;
                    r2 = (r2 >> 2) | ((u64)(3u - cc) << top);
                    const u32 w = (u32)f2 & mmask, rw = (u32)(r2 >> rsh) & mmask;
                    sc[i] = mmer_score(w < rw ? w : rw);
                }
            }
            u32 sv[NP + 16];
#pragma unroll
            for (int j = 0; j < NP; j++) {
                const int pos = j - NP;  // chunk-relative position, negative
                sv[j] = (pos >= -16) ? __shfl_up_sync(0xffffffffu, sc[(pos + 16) & 15], 1)
                                     : __shfl_up_sync(0xffffffffu, sc[(pos + 32) & 15], 2);
            }
#pragma unroll
            for (int i = 0; i < 16; i++) sv[NP + i] = sc[i];
            u32 win[16];
            WinMin<W>::run(sv, win);
#pragma unroll
            for (int i = 0; i < 16; i++) s_win[i * CNT_BLOCK + threadIdx.x] = win[i];
        }

        const u32 pv = (v2 << 16) | v1, ps = (s2 << 16) | s1;   // bit 0 = the most recent base
        u32 vrun = (pv == 0xffffffffu) ? 32u : (u32)__ffs(~pv) - 1u;
        u32 srun = ps ? (u32)__ffs(ps) - 1u : 32u;
        u32 codes = c.codes;
        u32 vm = (lane < ENC_HALO) ? 0u : (c.vmask << 16);     // halo lanes own no windows
        u32 sm = c.smask << 16;
        if (lane < ENC_HALO) vrun = 0;

#pragma unroll 1
        for (int b = 0; b < 4; b++) {
            u64 key[4];
            u32 bucket[4];
            u32 pend = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
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
                const u64 fm = f & kmask;
                key[i] = fm < rc ? fm : rc;
                pend |= (ok ? 1u : 0u) << i;
                if constexpr (W == -2) bucket[i] = prefix_home_bucket(key[i], fm < rc ? rc : fm, kmask >> 2, nbuckets);
            }
            // Probing proceeds in warp-uniform rounds: every round first issues the 256-bit bucket
            // loads of all still-pending keys (4 independent 32 B requests per lane in flight), then
            // resolves them, so lanes that need another bucket take it together.
            K4 q[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if constexpr (W == -2) { }
                else if constexpr (W > 0) bucket[i] = (u32)home_from_score(key[i], s_win[(b * 4 + i) * CNT_BLOCK + threadIdx.x], nbuckets, th);
                else if constexpr (W < 0) bucket[i] = (pend & (1u << i)) ? (u32)table_home(key[i], l, nbuckets, th) : 0u;
                else bucket[i] = (u32)hash_bucket(key[i], nbuckets);
                // L2 blocking knob: this launch only owns home buckets in [part_lo, part_hi)
                if (bucket[i] < part_lo || bucket[i] >= part_hi) pend &= ~(1u << i);
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
    }
    // per-warp reduction of the window counters
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        nl_tot += __shfl_xor_sync(0xffffffffu, nl_tot, d);
        nk_tot += __shfl_xor_sync(0xffffffffu, nk_tot, d);
    }
    if (lane == 0 && count_windows) {
        if (nl_tot) atomicAdd(stats + 0, (u64)nl_tot);
        if (nk_tot) atomicAdd(stats + 1, (u64)nk_tot);
    }
    if (overflow) atomicOr((unsigned long long *)(stats + 2), 1ull);
}

// ---- the default count kernel: compact the valid windows of a tile, then probe --------------------------
// In the kernel above every lane probes the windows of its own positions, so the lanes whose window is
// not a whole l-mer of one read (31 % of the positions of 100 bp reads at l = 32) idle through the probe
// code: ncu shows 21 of 32 threads active per instruction.  Here phase 1 rolls the 16 positions of the
// tile in lock step and packs the valid canonical keys into a per-warp shared-memory buffer (one ballot
// per position gives every lane its slot); phase 2 probes them 4 per lane with all lanes busy.  COHASH
// selects the home rule (plain hash of the key / of the canonical prefix k-mer, common.cuh).
#define CC_BLOCK 128
#define CC_KEYS (ENC_ADV * 16)   // valid windows per tile <= 480
#ifndef CC_KPL
#define CC_KPL 3   // keys in flight per lane (measured: 2..4 within 2 %)
#endif
#ifndef CC_MINB
#define CC_MINB 9
#endif
// MERGED: the table is an array of 32-byte buckets {key, key, key, counts} -- three keys and one word of
// counters (16 bits for slots 0 and 1 in the low half, 32 bits for slot 2 in the high half; 32-bit REDs:
// the 64-bit RED of a 3 x 21-bit layout ran at half rate) -- so the counter of a key lives in the sector
// the probe just loaded: one DRAM sector per insert instead of two (keys[] and counts[] of the SoA
// layout).  `cap` then counts 8-byte words (4 per bucket).  A 16-bit counter that overflows carries into
// its neighbour or is lost; the caller detects that through the checksum (sum of counts != windows) and
// reruns with the SoA layout.
#define MERGED_KEYS 3
template <bool COHASH, bool MERGED>
__global__ void __launch_bounds__(CC_BLOCK, CC_MINB) count_compact_kernel(const uint4 *__restrict__ buf16, u64 n_bases,
                                                                     const u32 *__restrict__ start_bits, u32 l,
                                                                     u64 *__restrict__ tab_keys, u32 *__restrict__ tab_cnt, u64 cap,
                                                                     u64 ntiles, u64 *__restrict__ stats)
{
    __shared__ u64 s_keys[(CC_BLOCK / 32) * CC_KEYS];
    u64 *stage = s_keys + (threadIdx.x >> 5) * CC_KEYS;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const u32 nbuckets = (u32)(cap / EULER_BUCKET);
    const u32 max_probe = nbuckets < 4096 ? nbuckets : 4096;
    const u32 k = l - 1, top = 2 * (l - 1);
    const u64 lmask = key_mask_d(l), kmask = lmask >> 2;
    u32 nl_tot = 0, nk_tot = 0;
    bool overflow = false;

    for (u64 tile = warp; tile < ntiles; tile += nwarps) {
        const long long chunk = (long long)(tile * ENC_ADV) - ENC_HALO + lane;
        const Chunk c = load_chunk(buf16, chunk, n_bases, start_bits);
        const u32 p1 = __shfl_up_sync(0xffffffffu, c.codes, 1), p2 = __shfl_up_sync(0xffffffffu, c.codes, 2);
        const u32 v1 = __shfl_up_sync(0xffffffffu, c.vmask, 1), v2 = __shfl_up_sync(0xffffffffu, c.vmask, 2);
        const u32 s1 = __shfl_up_sync(0xffffffffu, c.smask, 1), s2 = __shfl_up_sync(0xffffffffu, c.smask, 2);
        u64 f = ((u64)p2 << 32) | p1;
        u64 rc = revcomp64(f & lmask, l);
        const u32 pv = (v2 << 16) | v1, ps = (s2 << 16) | s1;   // bit 0 = the most recent base
        u32 vrun = (lane < ENC_HALO) ? 0u : ((pv == 0xffffffffu) ? 32u : (u32)__ffs(~pv) - 1u);
        u32 srun = ps ? (u32)__ffs(ps) - 1u : 32u;
        u32 codes = c.codes;
        u32 vm = (lane < ENC_HALO) ? 0u : (c.vmask << 16);      // halo lanes own no windows
        u32 sm = c.smask << 16;
        // phase 1: roll, pack the valid canonical keys (co-hash: keys whose prefix is the reverse strand get bit
        // patterns resolved again in phase 2 from the key itself, so only the key is staged)
        u32 nvalid = 0;
#pragma unroll
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
            const unsigned okm = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const u64 fm = f & lmask;
                stage[nvalid + __popc(okm & lt_mask)] = fm < rc ? fm : rc;
            }
            nvalid += __popc(okm);
        }
        nl_tot += nvalid;   // warp-uniform: counted once per warp below
        __syncwarp();
        // phase 2: probe CC_KPL keys per lane, every lane busy.  Pass 0 gives every key ONE probe step (its home
        // bucket); the few keys whose home is full (~7 %) are packed back into the staging buffer and finished
        // together in pass 1 -- otherwise every group of 96 keys would pay extra warp-wide rounds for its 6 stragglers.
        u32 ncur = nvalid;
#pragma unroll 1
        for (int pass = 0; pass < 2 && ncur; pass++) {
            u32 nleft = 0;
#pragma unroll 1
            for (u32 g = 0; g < ncur; g += 32 * CC_KPL) {
                u64 key[CC_KPL];
                u32 bucket[CC_KPL];
                K4 q[CC_KPL];
                u32 pend = 0;
#pragma unroll
                for (int i = 0; i < CC_KPL; i++) {
                    const u32 idx = g + i * 32 + lane;
                    key[i] = 0;
                    if (idx < ncur) {
                        key[i] = stage[idx];
                        pend |= 1u << i;
                    }
                    if (COHASH) bucket[i] = prefix_home_bucket(key[i], revcomp64(key[i], l), kmask, nbuckets);
                    else bucket[i] = (u32)hash_bucket(key[i], nbuckets);
                    if (pass && ++bucket[i] == nbuckets) bucket[i] = 0;   // stragglers: the home bucket was full
                }
                __syncwarp();   // all keys of this group are in registers before stragglers overwrite the buffer
                u32 probes = 0;
                while (__any_sync(0xffffffffu, pend != 0)) {
#pragma unroll
                    for (int i = 0; i < CC_KPL; i++)
                        if (pend & (1u << i)) q[i] = ld_bucket_cg(tab_keys + (u64)bucket[i] * EULER_BUCKET);
#pragma unroll
                    for (int i = 0; i < CC_KPL; i++) {
                        if (pend & (1u << i)) {
                            u64 *bk = tab_keys + (u64)bucket[i] * EULER_BUCKET;
                            int fe;
                            int j = bucket_match<MERGED ? MERGED_KEYS : EULER_BUCKET>(q[i], key[i], fe);
                            while (j < 0 && fe < (MERGED ? MERGED_KEYS : EULER_BUCKET)) {   // claim the first empty slot
                                const u64 old = atomicCAS(bk + fe, EULER_EMPTY_KEY, key[i]);
                                if (old == EULER_EMPTY_KEY || old == key[i]) j = fe;
                                else fe++;   // lost the race for this slot: the next one is tried through its own CAS
                            }
                            if (j >= 0) {
                                if (MERGED) atomicAdd((u32 *)(bk + 3) + (j >> 1), j == 1 ? 0x10000u : 1u);
                                else atomicAdd(tab_cnt + (u64)bucket[i] * EULER_BUCKET + j, 1u);
                                pend &= ~(1u << i);
                            } else if (++bucket[i] == nbuckets) {
                                bucket[i] = 0;
                            }
                        }
                    }
                    if (pass == 0) break;
                    if (++probes >= max_probe && pend) { overflow = true; pend = 0; }
                }
                if (pass == 0) {   // pack the stragglers to the front of the buffer (fewer than the keys already consumed)
#pragma unroll
                    for (int i = 0; i < CC_KPL; i++) {
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
    for (int d = 16; d >= 1; d >>= 1) nk_tot += __shfl_xor_sync(0xffffffffu, nk_tot, d);
    if (lane == 0) {
        if (nl_tot) atomicAdd(stats + 0, (u64)nl_tot);
        if (nk_tot) atomicAdd(stats + 1, (u64)nk_tot);
    }
    if (overflow) atomicOr((unsigned long long *)(stats + 2), 1ull);
}

// merged table -> the SoA layout the graph stage reads: keys[4 b + j] / cnt[4 b + j] for j < 3, slot 3 of every bucket empty
__global__ void __launch_bounds__(256) merged_unpack_kernel(const uint4 *__restrict__ tab, u64 nbuckets, uint4 *__restrict__ keys,
                                                            uint4 *__restrict__ cnt)
{
    const u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbuckets) return;
    const uint4 lo = tab[2 * b], hi = tab[2 * b + 1];   // {k0, k1}, {k2, counters}
    keys[2 * b] = lo;
    keys[2 * b + 1] = make_uint4(hi.x, hi.y, 0xffffffffu, 0xffffffffu);
    cnt[b] = make_uint4(hi.z & 0xffffu, hi.z >> 16, hi.w, 0u);
}
__global__ void __launch_bounds__(256) merged_clear_kernel(uint4 *__restrict__ tab, u64 nbuckets)
{
    const u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbuckets) return;
    tab[2 * b] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    tab[2 * b + 1] = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);
}
// count into a merged table `tab` (cap_words = 4 * buckets, caller-cleared with enc_merged_clear), then unpack
int enc_merged_clear(euler_ctx *ctx, u64 *tab, u64 cap_words)
{
    const u64 nb = cap_words / 4;
    merged_clear_kernel<<<grid_for(nb, 256), 256, 0, ctx->stream>>>((uint4 *)tab, nb);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}
int enc_count_merged(euler_ctx *ctx, const void *d_buf, u64 n_bases, const u32 *d_bits, u32 l, u64 *tab, u64 cap_words, bool cohash,
                     u64 *d_stats)
{
    if (!n_bases) return EULER_OK;
    const u64 nchunks = (n_bases + 15) / 16;
    const u64 ntiles = (nchunks + ENC_ADV - 1) / ENC_ADV;
    u64 g2 = (u64)ctx->num_sms * CC_MINB;
    const u64 need2 = (ntiles + CC_BLOCK / 32 - 1) / (CC_BLOCK / 32);
    if (g2 > need2) g2 = need2;
    if (cohash)
        count_compact_kernel<true, true><<<(unsigned)g2, CC_BLOCK, 0, ctx->stream>>>((const uint4 *)d_buf, n_bases, d_bits, l, tab,
                                                                                    nullptr, cap_words, ntiles, d_stats);
    else
        count_compact_kernel<false, true><<<(unsigned)g2, CC_BLOCK, 0, ctx->stream>>>((const uint4 *)d_buf, n_bases, d_bits, l, tab,
                                                                                     nullptr, cap_words, ntiles, d_stats);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}
int enc_merged_unpack(euler_ctx *ctx, const u64 *tab, u64 cap_words, u64 *keys, u32 *cnt)
{
    const u64 nb = cap_words / 4;
    merged_unpack_kernel<<<grid_for(nb, 256), 256, 0, ctx->stream>>>((const uint4 *)tab, nb, (uint4 *)keys, (uint4 *)cnt);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

int enc_count_canonical(euler_ctx *ctx, const void *d_buf, u64 n_bases, const u32 *d_bits, u32 l, u64 *tab_keys,
                        u32 *tab_cnt, u64 cap, TableHash th, u64 *d_stats)
{
    if (!n_bases) return EULER_OK;
    const u64 nchunks = (n_bases + 15) / 16;
    const u64 ntiles = (nchunks + ENC_ADV - 1) / ENC_ADV;
    const u64 warps_per_block = CNT_BLOCK / 32;
    u64 grid = (u64)ctx->num_sms * 4;
    const u64 need = (ntiles + warps_per_block - 1) / warps_per_block;
    if (grid > need) grid = need;
    // L2-blocking knob (EULER_B200_COUNT_PARTS): one pass over the reads per contiguous bucket range.
    // Measured: re-encoding costs more than the misses it saves, so the default is a single pass.
    u64 parts = 1;
    const char *env = getenv("EULER_B200_COUNT_PARTS");
    if (env && atoi(env) > 0) parts = (u64)atoi(env);
    const int W = th.span_nb ? (int)(l - th.m + 1) : (th.m == EULER_PREFIX_HOME ? -2 : 0);
    static int compact = -1;
    if (compact < 0) {
        const char *e = getenv("EULER_B200_COUNT_COMPACT");   // 0: the per-position kernel above
        compact = (e && atoi(e) == 0) ? 0 : 1;
    }
    if (compact && parts == 1 && (W == 0 || W == -2)) {
        u64 g2 = (u64)ctx->num_sms * CC_MINB;
        const u64 need2 = (ntiles + CC_BLOCK / 32 - 1) / (CC_BLOCK / 32);
        if (g2 > need2) g2 = need2;
        if (W == 0)
            count_compact_kernel<false, false><<<(unsigned)g2, CC_BLOCK, 0, ctx->stream>>>((const uint4 *)d_buf, n_bases, d_bits, l, tab_keys,
                                                                                   tab_cnt, cap, ntiles, d_stats);
        else
            count_compact_kernel<true, false><<<(unsigned)g2, CC_BLOCK, 0, ctx->stream>>>((const uint4 *)d_buf, n_bases, d_bits, l, tab_keys,
                                                                                  tab_cnt, cap, ntiles, d_stats);
        CUDA_TRY(ctx, cudaGetLastError());
        return EULER_OK;
    }
    for (u64 p = 0; p < parts; p++) {
        const u64 nb = cap / EULER_BUCKET;
        const u64 lo = nb * p / parts, hi = nb * (p + 1) / parts;
        const unsigned g = (unsigned)(grid ? grid : 1);
        const int cw = p == 0 ? 1 : 0;
#define LAUNCH_CNT(WW)                                                                                          \
    count_canonical_kernel<WW><<<g, CNT_BLOCK, 0, ctx->stream>>>((const uint4 *)d_buf, n_bases, d_bits, l, tab_keys, \
                                                                 tab_cnt, cap, th, ntiles, d_stats, lo, hi, cw)
        if (W == 0) LAUNCH_CNT(0);
        else if (W == -2) LAUNCH_CNT(-2);
        else if (W == 21) LAUNCH_CNT(21);   // l = 32 (k = 31), m = 12
        else if (W == 11) LAUNCH_CNT(11);   // l = 22 (k = 21), m = 12
        else LAUNCH_CNT(-1);                // any other l: per-window brute-force minimizer
#undef LAUNCH_CNT
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}
