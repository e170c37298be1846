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

// Owner of a vertex = re-mixed score of its minimizer (smallest scrambled canonical m-mer), scaled
// to the rank count.  Adjacent k-mers almost always share their minimizer, so the prefix and suffix
// vertex of an l-mer have the same owner ~(1 - 2/(k-m+2)) of the time: ~1.1 copies of every l-mer
// cross the fabric instead of ~(2 - 1/N) with a per-k-mer hash.
#define DIST_M 12
__device__ __forceinline__ u32 dist_m(u32 k) { return k < DIST_M ? k : DIST_M; }
__device__ __forceinline__ u32 owner_from_score(u32 score, u32 nranks)
{
    // The minimum of many scores hugs 0 in its HIGH bits; its low 16 bits (low16(c * odd) xor higher
    // product bits, see mmer_score) stay uniform, so they are what is scaled to the rank count.
    return ((score & 0xffffu) * nranks) >> 16;
}
// owners of the prefix and suffix k-mer of an l-mer (either orientation: the minimizer is strand symmetric)
__device__ __forceinline__ void owners_of_lmer(u64 x, u32 l, u32 nranks, u32 &own_prefix, u32 &own_suffix)
{
    u32 all, but_last, but_first;
    min_scores(x, l, dist_m(l - 1), all, but_last, but_first);
    own_prefix = owner_from_score(but_last, nranks);
    own_suffix = owner_from_score(but_first, nranks);
}

// sliding minima for the partition kernel: s[t], t = position + WK, holds the m-mer scores of the
// WK positions before the chunk and the chunk's 16; for the l-mer ending at chunk position i the
// prefix k-mer owns m-mers ending at positions i-WK .. i-1 and the suffix k-mer i-WK+1 .. i.
template <int WK>
struct WinMin2 {
    static constexpr int N = WK + 16;
    static constexpr int E = (WK >= 16) ? 4 : (WK >= 8) ? 3 : (WK >= 4) ? 2 : (WK >= 2) ? 1 : 0;
    static constexpr int P = 1 << E;
    __device__ __forceinline__ static void run(u32 (&s)[N], u32 (&wp)[16], u32 (&ws)[16])
    {
#pragma unroll
        for (int e = 0; e < E; e++) {
            const int h = 1 << e;
#pragma unroll
            for (int t = 0; t + h < N; t++) s[t] = s[t] < s[t + h] ? s[t] : s[t + h];
        }
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const u32 a = s[i], b = s[i + WK - P], c = s[i + 1], d = s[i + 1 + WK - P];
            wp[i] = a < b ? a : b;
            ws[i] = c < d ? c : d;
        }
    }
};

// ---- pass over the reads: per-destination counts (SCATTER=false) or key scatter (SCATTER=true) --
// SCATTER writes into fixed-capacity per-destination segments (seg_cap keys each); a cursor that
// runs past its segment only counts (the caller then falls back to exact sizes).
#define PB 128          // partition kernel block: 4 warps, each with an 8 KB staging area
#define PB_STAGE 1024   // keys per warp staging area (a tile emits at most 30 * 16 * 2 = 960)
template <bool SCATTER, int WK>
__global__ void __launch_bounds__(PB, 4) dist_partition_kernel(const uint4 *__restrict__ buf16, u64 n_bases,
                                                                const u32 *__restrict__ start_bits, u32 l, u32 nranks, u64 ntiles,
                                                                u64 *__restrict__ counts /* [16] per dest, [16] N_l, [17] N_k */,
                                                                u64 *__restrict__ cursors, u64 *__restrict__ send,
                                                                const u64 *__restrict__ seg_off, u64 seg_cap)
{
    __shared__ u32 s_own[16 * PB];
    __shared__ u64 s_stage[SCATTER ? (PB / 32) * PB_STAGE : 1];
    u64 *stage = s_stage + (SCATTER ? (threadIdx.x >> 5) * PB_STAGE : 0);
    const int lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const u32 k = l - 1, top = 2 * (l - 1);
    const u64 lmask = key_mask_d(l);
    const u32 m = dist_m(k);
    u32 nl_tot = 0, nk_tot = 0;

    for (u64 tile = warp; tile < ntiles; tile += nwarps) {
        const long long chunk = (long long)(tile * ENC_ADV) - ENC_HALO + lane;
        const Chunk c = load_chunk(buf16, chunk, n_bases, start_bits);
        const u32 p1 = __shfl_up_sync(0xffffffffu, c.codes, 1), p2 = __shfl_up_sync(0xffffffffu, c.codes, 2);
        const u32 v1 = __shfl_up_sync(0xffffffffu, c.vmask, 1), v2 = __shfl_up_sync(0xffffffffu, c.vmask, 2);
        const u32 s1 = __shfl_up_sync(0xffffffffu, c.smask, 1), s2 = __shfl_up_sync(0xffffffffu, c.smask, 2);
        u64 f = ((u64)p2 << 32) | p1;
        u64 rc = revcomp64(f & lmask, l);
        if constexpr (WK > 0) {
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
            u32 sv[WK + 16];
#pragma unroll
            for (int j = 0; j < WK; j++) {
                const int pos = j - WK;
                sv[j] = (pos >= -16) ? __shfl_up_sync(0xffffffffu, sc[(pos + 16) & 15], 1)
                                     : __shfl_up_sync(0xffffffffu, sc[(pos + 32) & 15], 2);
            }
#pragma unroll
            for (int i = 0; i < 16; i++) sv[WK + i] = sc[i];
            u32 wp[16], ws[16];
            WinMin2<WK>::run(sv, wp, ws);
#pragma unroll
            for (int i = 0; i < 16; i++)
                s_own[i * PB + threadIdx.x] = owner_from_score(wp[i], nranks) | (owner_from_score(ws[i], nranks) << 8);
        }
        const u32 pv = (v2 << 16) | v1, ps = (s2 << 16) | s1;
        const u32 vrun0 = (lane < ENC_HALO) ? 0u : ((pv == 0xffffffffu) ? 32u : (u32)__ffs(~pv) - 1u);
        const u32 srun0 = ps ? (u32)__ffs(ps) - 1u : 32u;
        const u32 vm0 = (lane < ENC_HALO) ? 0u : (c.vmask << 16);   // halo lanes own no windows
        const u32 sm0 = c.smask << 16;
        const u64 f0 = f, rc0 = rc;

        // pass 1: classify the 16 windows of this lane; per-destination counts in packed 16-bit fields
        // (4 destinations per u64; a warp sends at most 32 * 32 keys to one destination per tile)
        u64 cnt[4] = {0, 0, 0, 0};
        u32 okmask = 0;
        {
            u32 vrun = vrun0, srun = srun0, codes = c.codes, vm = vm0, sm = sm0;
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
                if (!ok) continue;
                nl_tot++;
                okmask |= 1u << i;
                u32 o1, o2;
                if constexpr (WK > 0) {
                    const u32 o = s_own[i * PB + threadIdx.x];
                    o1 = o & 0xffu;
                    o2 = o >> 8;
                } else {
                    owners_of_lmer(f & lmask, l, nranks, o1, o2);
                    s_own[i * PB + threadIdx.x] = o1 | (o2 << 8);
                }
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    if ((o1 >> 2) == (u32)g) cnt[g] += 1ull << (16 * (o1 & 3));
                    if (o2 != o1 && (o2 >> 2) == (u32)g) cnt[g] += 1ull << (16 * (o2 & 3));
                }
            }
        }
        const u32 ngroups = (nranks + 3) >> 2;
        if (!SCATTER) {
#pragma unroll
            for (int g = 0; g < 4; g++) {
                if ((u32)g < ngroups) {
                    u64 t = cnt[g];
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                    if ((lane >> 2) == g && (u32)lane < nranks) {
                        const u64 mine = (t >> (16 * (lane & 3))) & 0xffffull;
                        if (mine) atomicAdd(counts + lane, mine);
                    }
                }
            }
            continue;
        }
        // warp-exclusive scan of the packed counts; ONE cursor atomic per destination per tile
        u64 excl[4] = {0, 0, 0, 0};
        u64 base_reg = 0;
        u32 mine_reg = 0;   // lane d: keys this warp sends to destination d in this tile
#pragma unroll
        for (int g = 0; g < 4; g++) {
            if ((u32)g < ngroups) {
                u64 inc = cnt[g];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const u64 t = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += t;
                }
                excl[g] = inc - cnt[g];
                const u64 tot = __shfl_sync(0xffffffffu, inc, 31);
                if ((lane >> 2) == g && (u32)lane < nranks) {
                    const u64 mine = (tot >> (16 * (lane & 3))) & 0xffffull;
                    mine_reg = (u32)mine;
                    if (mine) base_reg = atomicAdd(cursors + lane, mine);
                }
            }
        }
        const u64 seg_reg = (u32)lane < nranks ? seg_off[lane] : 0ull;
        // warp-local layout of the staging area: destination after destination
        u32 woff_reg = mine_reg;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, woff_reg, d);
            if (lane >= d) woff_reg += t;
        }
        woff_reg -= mine_reg;
        // pass 2: roll again and stage every key at woff[dest] + keys of lower lanes + own keys so far
        {
            u64 f2 = f0, r2 = rc0;
            u32 codes = c.codes;
            u64 run[4] = {0, 0, 0, 0};
#pragma unroll 1
            for (int i = 0; i < 16; i++) {
                const u32 cc = codes >> 30;
                codes <<= 2;
                f2 = (f2 << 2) | cc;
                r2 = (r2 >> 2) | ((u64)(3u - cc) << top);
                const bool ok = (okmask >> i) & 1u;
                const u32 o = ok ? s_own[i * PB + threadIdx.x] : 0u;
                const u32 o1 = o & 0xffu, o2 = o >> 8;
                const u64 fm = f2 & lmask;
                const u64 key = fm < r2 ? fm : r2;
#pragma unroll
                for (int pass = 0; pass < 2; pass++) {
                    const u32 dst = pass ? o2 : o1;
                    const u32 wo = __shfl_sync(0xffffffffu, woff_reg, dst & 31);
                    if (!ok || (pass && o2 == o1)) continue;
                    u64 e = 0, r = 0;
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        if ((dst >> 2) == (u32)g) { e = excl[g]; r = run[g]; run[g] += 1ull << (16 * (dst & 3)); }
                    }
                    const u32 sh = 16 * (dst & 3);
                    stage[wo + (u32)((e >> sh) & 0xffffull) + (u32)((r >> sh) & 0xffffull)] = key;
                }
            }
        }
        __syncwarp();
        // copy-out: contiguous, fully coalesced runs per destination (local HBM or a peer's over NVLink)
        for (u32 d = 0; d < nranks; d++) {
            const u32 tot = __shfl_sync(0xffffffffu, mine_reg, d);
            const u64 base = __shfl_sync(0xffffffffu, base_reg, d);
            const u64 seg = __shfl_sync(0xffffffffu, seg_reg, d);
            const u32 wo = __shfl_sync(0xffffffffu, woff_reg, d);
            for (u32 j = lane; j < tot; j += 32) {
                const u64 at = base + j;
                if (at < seg_cap) send[seg + at] = stage[wo + j];
            }
        }
        __syncwarp();
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

// ---- the same pass for nranks <= 8 with rolling minimizers (the multi-GPU hot kernel) -------------
// One roll for the m-mer scores and the sliding minima, owners kept as 4-bit fields of two u64
// registers, per-lane per-destination counts as 8-bit fields of ONE u64, everything unrolled so that
// nothing is indexed dynamically (no local memory, no shared owner table).  The warp scan runs on
// 16-bit fields; prefix sums ACROSS the fields of a word are one multiply by 0x0001000100010001.
#define P8_FIELDS 0x0001000100010001ull
#ifndef P8_MINB
#define P8_MINB 4   // resident blocks per SM (128 registers); 5 and 6 spill
#endif
__device__ __forceinline__ u64 spread8to16(u32 x)
{
    u64 t = x;
    t = (t | (t << 16)) & 0x0000FFFF0000FFFFull;
    return (t | (t << 8)) & 0x00FF00FF00FF00FFull;
}
template <bool SCATTER, int WK>
__global__ void __launch_bounds__(PB, P8_MINB) dist_partition8_kernel(const uint4 *__restrict__ buf16, u64 n_bases,
                                                                 const u32 *__restrict__ start_bits, u32 l, u32 nranks, u64 ntiles,
                                                                 u64 *__restrict__ counts, u64 *__restrict__ cursors,
                                                                 u64 *__restrict__ send, const u64 *__restrict__ seg_off, u64 seg_cap)
{
    __shared__ u64 s_stage[SCATTER ? (PB / 32) * PB_STAGE : 1];
    u64 *stage = s_stage + (SCATTER ? (threadIdx.x >> 5) * PB_STAGE : 0);
    const int lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const u32 k = l - 1, top = 2 * (l - 1);
    const u64 lmask = key_mask_d(l);
    const u32 mmask = (1u << (2 * DIST_M)) - 1u;   // WK > 0 implies k >= DIST_M
    const u32 rsh = 2 * WK;
    const bool wide8 = nranks > 4;
    u32 nl_tot = 0, nk_tot = 0;
    const u64 seg_reg = (u32)lane < nranks ? seg_off[lane] : 0ull;

    for (u64 tile = warp; tile < ntiles; tile += nwarps) {
        const long long chunk = (long long)(tile * ENC_ADV) - ENC_HALO + lane;
        const Chunk c = load_chunk(buf16, chunk, n_bases, start_bits);
        const u32 p1 = __shfl_up_sync(0xffffffffu, c.codes, 1), p2 = __shfl_up_sync(0xffffffffu, c.codes, 2);
        const u32 v1 = __shfl_up_sync(0xffffffffu, c.vmask, 1), v2 = __shfl_up_sync(0xffffffffu, c.vmask, 2);
        const u32 s1 = __shfl_up_sync(0xffffffffu, c.smask, 1), s2 = __shfl_up_sync(0xffffffffu, c.smask, 2);
        const u64 f0 = ((u64)p2 << 32) | p1;
        const u64 rc0 = revcomp64(f0 & lmask, l);
        // m-mer scores of the 16 positions of this chunk, then the WK before them from the lanes below
        u32 wp[16], ws[16];
        {
            u32 sc[16];
            u64 f2 = f0, r2 = rc0;
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
            u32 sv[WK + 16];
#pragma unroll
            for (int j = 0; j < WK; j++) {
                const int pos = j - WK;
                sv[j] = (pos >= -16) ? __shfl_up_sync(0xffffffffu, sc[(pos + 16) & 15], 1)
                                     : __shfl_up_sync(0xffffffffu, sc[(pos + 32) & 15], 2);
            }
#pragma unroll
            for (int i = 0; i < 16; i++) sv[WK + i] = sc[i];
            WinMin2<WK>::run(sv, wp, ws);
        }
        // validity, owners, per-destination counts
        const u32 pv = (v2 << 16) | v1, ps = (s2 << 16) | s1;
        u32 vrun = (lane < ENC_HALO) ? 0u : ((pv == 0xffffffffu) ? 32u : (u32)__ffs(~pv) - 1u);
        u32 srun = ps ? (u32)__ffs(ps) - 1u : 32u;
        u32 vm = (lane < ENC_HALO) ? 0u : (c.vmask << 16);   // halo lanes own no windows
        u32 sm = c.smask << 16;
        u64 own1 = 0, own2 = 0, cnt = 0;
        u32 okmask = 0, nk = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const bool valid = (vm >> 31) != 0, start = (sm >> 31) != 0;
            vm <<= 1;
            sm <<= 1;
            vrun = valid ? vrun + 1u : 0u;
            srun = start ? 0u : srun + 1u;
            nk += (vrun >= k && srun + 1u >= k) ? 1u : 0u;
            const bool ok = vrun >= l && srun + 1u >= l;
            const u32 o1 = owner_from_score(wp[i], nranks), o2 = owner_from_score(ws[i], nranks);
            own1 |= (u64)o1 << (4 * i);
            own2 |= (u64)o2 << (4 * i);
            if (ok) {
                okmask |= 1u << i;
                cnt += 1ull << (8 * o1);
                if (o2 != o1) cnt += 1ull << (8 * o2);
            }
        }
        nk_tot += nk;
        nl_tot += __popc(okmask);
        // warp scan of the counts (16-bit fields: destinations 0-3 in lo, 4-7 in hi)
        const u64 cLo = spread8to16((u32)cnt), cHi = wide8 ? spread8to16((u32)(cnt >> 32)) : 0ull;
        u64 iLo = cLo, iHi = cHi;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u64 t = __shfl_up_sync(0xffffffffu, iLo, d);
            if (lane >= d) iLo += t;
        }
        if (wide8) {
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u64 t = __shfl_up_sync(0xffffffffu, iHi, d);
                if (lane >= d) iHi += t;
            }
        }
        const u64 tLo = __shfl_sync(0xffffffffu, iLo, 31), tHi = wide8 ? __shfl_sync(0xffffffffu, iHi, 31) : 0ull;
        const u32 mine_reg = (u32)(((lane & 4) ? tHi : tLo) >> (16 * (lane & 3))) & 0xffffu;   // lane d < 8: keys for destination d
        if (!SCATTER) {
            if ((u32)lane < nranks && mine_reg) atomicAdd(counts + lane, (u64)mine_reg);
            continue;
        }
        u64 base_reg = 0;
        if ((u32)lane < nranks && mine_reg) base_reg = atomicAdd(cursors + lane, (u64)mine_reg);   // ONE cursor atomic per destination per tile
        // staging layout: destination after destination; start of each = prefix sum across the fields
        const u64 fLo = tLo * P8_FIELDS, fHi = tHi * P8_FIELDS + (fLo >> 48) * P8_FIELDS;
        const u64 wLo = fLo - tLo, wHi = fHi - tHi;
        u64 posLo = wLo + (iLo - cLo), posHi = wHi + (iHi - cHi);
        const u32 woff_reg = (u32)(((lane & 4) ? wHi : wLo) >> (16 * (lane & 3))) & 0xffffu;
        {
            u64 f = f0, rc = rc0;
            u32 codes = c.codes;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const u32 cc = codes >> 30;
                codes <<= 2;
                f = (f << 2) | cc;
                rc = (rc >> 2) | ((u64)(3u - cc) << top);
                if ((okmask >> i) & 1u) {
                    const u64 fm = f & lmask;
                    const u64 key = fm < rc ? fm : rc;
                    const u32 o1 = (u32)(own1 >> (4 * i)) & 15u, o2 = (u32)(own2 >> (4 * i)) & 15u;
                    {
                        const u32 sh = 16 * (o1 & 3);
                        u32 slot;
                        if (o1 & 4) { slot = (u32)(posHi >> sh) & 0xffffu; posHi += 1ull << sh; }
                        else { slot = (u32)(posLo >> sh) & 0xffffu; posLo += 1ull << sh; }
                        stage[slot] = key;
                    }
                    if (o2 != o1) {
                        const u32 sh = 16 * (o2 & 3);
                        u32 slot;
                        if (o2 & 4) { slot = (u32)(posHi >> sh) & 0xffffu; posHi += 1ull << sh; }
                        else { slot = (u32)(posLo >> sh) & 0xffffu; posLo += 1ull << sh; }
                        stage[slot] = key;
                    }
                }
            }
        }
        __syncwarp();
        // copy-out: contiguous, fully coalesced runs per destination (local HBM or a peer's over NVLink)
        for (u32 d = 0; d < nranks; d++) {
            const u32 tot = __shfl_sync(0xffffffffu, mine_reg, d);
            const u64 base = __shfl_sync(0xffffffffu, base_reg, d);
            const u64 seg = __shfl_sync(0xffffffffu, seg_reg, d);
            const u32 wo = __shfl_sync(0xffffffffu, woff_reg, d);
            for (u32 j = lane; j < tot; j += 32) {
                const u64 at = base + j;
                if (at < seg_cap) send[seg + at] = stage[wo + j];
            }
        }
        __syncwarp();
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

int dist_partition(euler_ctx *ctx, bool scatter, const void *d_buf, u64 n_bases, const u32 *d_bits, u32 l, u32 nranks,
                   u64 *d_counts, u64 *d_cursors, u64 *d_send, const u64 *d_seg_off, u64 seg_cap)
{
    if (!n_bases) return EULER_OK;
    const u64 nchunks = (n_bases + 15) / 16;
    const u64 ntiles = (nchunks + ENC_ADV - 1) / ENC_ADV;
    u64 grid = (u64)ctx->num_sms * 4;
    const u64 need = (ntiles + PB / 32 - 1) / (PB / 32);
    if (grid > need) grid = need;
    const u32 k = l - 1;
    const int WK = (int)(l - (k < DIST_M ? k : DIST_M));
    const uint4 *b16 = (const uint4 *)d_buf;
    const unsigned g = (unsigned)grid;
#define LAUNCH_PART(S, WW)                                                                                              \
    dist_partition_kernel<S, WW><<<g, PB, 0, ctx->stream>>>(b16, n_bases, d_bits, l, nranks, ntiles, d_counts, d_cursors, \
                                                            d_send, d_seg_off, seg_cap)
#define LAUNCH_P8(S, WW)                                                                                                  \
    dist_partition8_kernel<S, WW><<<g8, PB, 0, ctx->stream>>>(b16, n_bases, d_bits, l, nranks, ntiles, d_counts, d_cursors, \
                                                              d_send, d_seg_off, seg_cap)
    if (nranks <= 8 && (WK == 20 || WK == 10)) {
        u64 grid8 = (u64)ctx->num_sms * P8_MINB;
        if (grid8 > need) grid8 = need;
        const unsigned g8 = (unsigned)grid8;
        if (scatter) { if (WK == 20) LAUNCH_P8(true, 20); else LAUNCH_P8(true, 10); }
        else { if (WK == 20) LAUNCH_P8(false, 20); else LAUNCH_P8(false, 10); }
        CUDA_TRY(ctx, cudaGetLastError());
        return EULER_OK;
    }
#undef LAUNCH_P8
    if (!scatter) {
        if (WK == 20) LAUNCH_PART(false, 20);
        else if (WK == 10) LAUNCH_PART(false, 10);
        else LAUNCH_PART(false, -1);
    } else {
        if (WK == 20) LAUNCH_PART(true, 20);
        else if (WK == 10) LAUNCH_PART(true, 10);
        else LAUNCH_PART(true, -1);
    }
#undef LAUNCH_PART
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- count received canonical keys into the local table ----------------------------------------
// The key stream is read once: it goes through L2 with an evict-first policy so that it does not
// push the table (the only data with reuse) out.  stats[5] += number of slots newly claimed, i.e. the
// distinct canonical keys this rank holds -- the size the next run of the same input needs.
__device__ __forceinline__ u64 ld_evict_first_u64(const u64 *p, u64 policy)
{
    u64 v;
    asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ int bucket_claim_fresh(u64 *bucket_keys, const K4 &q, u64 key, u32 &fresh)
{
#pragma unroll
    for (int j = 0; j < EULER_BUCKET; j++) {
        const u64 kv = q.k[j];
        if (kv == key) return j;
        if (kv == EULER_EMPTY_KEY) {
            const u64 old = atomicCAS(bucket_keys + j, EULER_EMPTY_KEY, key);
            if (old == EULER_EMPTY_KEY) { fresh++; return j; }
            if (old == key) return j;
        }
    }
    return -1;
}
// home bucket of a received canonical l-mer: by the key, or (co-hashed tables, common.cuh) by its canonical prefix k-mer
__device__ __forceinline__ u32 dist_home_bucket(u64 key, u32 cohash_l, u32 nbuckets)
{
    if (!cohash_l) return (u32)hash_bucket(key, nbuckets);
    return prefix_home_bucket(key, revcomp64(key, cohash_l), key_mask_d(cohash_l - 1), nbuckets);
}
__global__ void __launch_bounds__(DB, 4) dist_count_keys_kernel(const u64 *__restrict__ keys_in, u64 n, u64 *__restrict__ tab_keys,
                                                                 u32 *__restrict__ tab_cnt, u64 cap, u32 cohash_l,
                                                                 u64 *__restrict__ stats)
{
    const u32 nbuckets = (u32)(cap / EULER_BUCKET);
    const u32 max_probe = nbuckets < 4096 ? nbuckets : 4096;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 t0 = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool overflow = false;
    u32 fresh = 0;
    u64 policy;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
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
            key[i] = idx < n ? ld_evict_first_u64(keys_in + idx, policy) : 0;
            if (idx < n) pend |= 1u << i;
            bucket[i] = dist_home_bucket(key[i], cohash_l, nbuckets);
        }
        u32 probes = 0;
        while (__any_sync(0xffffffffu, pend != 0)) {
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (pend & (1u << i)) q[i] = ld_bucket_cg(tab_keys + (u64)bucket[i] * EULER_BUCKET);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (pend & (1u << i)) {
                    const int j = bucket_claim_fresh(tab_keys + (u64)bucket[i] * EULER_BUCKET, q[i], key[i], fresh);
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
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) fresh += __shfl_xor_sync(0xffffffffu, fresh, o);
    if ((threadIdx.x & 31) == 0 && fresh) atomicAdd(stats + 5, (u64)fresh);
    if (overflow) atomicOr((unsigned long long *)(stats + 2), 1ull);
}

int dist_count_keys(euler_ctx *ctx, const u64 *d_keys, u64 n, u64 *tab_keys, u32 *tab_cnt, u64 cap, u32 cohash_l, u64 *d_stats)
{
    if (!n) return EULER_OK;
    u64 grid = (u64)ctx->num_sms * 8;
    const u64 need = (n + DB * 4 - 1) / (DB * 4);
    if (grid > need) grid = need;
    dist_count_keys_kernel<<<(unsigned)grid, DB, 0, ctx->stream>>>(d_keys, n, tab_keys, tab_cnt, cap, cohash_l, d_stats);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- L2 blocking for tables far larger than L2 -----------------------------------------------------
// A 3 GB table (1 Gbp over 8 GPUs) turns every insert into DRAM sector traffic plus TLB misses.  The
// received keys are first split into `nparts` runs by the high bits of their table hash, so that
// each run only touches one contiguous ~48 MB stretch of the table; the runs are then counted one
// after another with that stretch resident in L2.  Order inside a run is irrelevant, so the split
// is a counting sort per 4096-key tile in shared memory with ONE global cursor atomic per
// (tile, part) and coalesced run copies.  Parts have a fixed capacity (the hash is uniform); a part
// that overflows raises flags[0] and the caller falls back to the unblocked kernel.
#define BK_THREADS 256
#define BK_ITEMS 16
#define BK_TILE (BK_THREADS * BK_ITEMS)
#define BK_MAXP 256
__device__ __forceinline__ u32 key_part(u64 key, u32 nparts, u32 cohash_l)
{
    if (cohash_l) {
        u32 flip;
        key = prefix_home_key(key, revcomp64(key, cohash_l), key_mask_d(cohash_l - 1), flip);
    }
    const u64 h = (key ^ (key >> 29)) * 0x9E3779B97F4A7C15ull;   // same mix as hash_bucket: parts are bucket ranges
    return (u32)(((h >> 32) * (u64)nparts) >> 32);
}
__global__ void __launch_bounds__(BK_THREADS) dist_block_keys_kernel(const u64 *__restrict__ in, u64 n, u32 nparts,
                                                                      u64 *__restrict__ cursors, u64 *__restrict__ out, u64 part_cap,
                                                                      u32 cohash_l, u64 *__restrict__ flags)
{
    __shared__ u64 stage[BK_TILE];
    __shared__ u32 hist[BK_MAXP], loff[BK_MAXP], fill[BK_MAXP], wsum[BK_THREADS / 32];
    __shared__ u64 gbase[BK_MAXP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u64 ntiles = (n + BK_TILE - 1) / BK_TILE;
    u64 policy;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    bool over = false;
    for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        hist[tid] = 0;
        fill[tid] = 0;
        __syncthreads();
        const u64 base = tile * BK_TILE;
        const u32 cnt = (u32)(n - base < BK_TILE ? n - base : BK_TILE);
        u64 k[BK_ITEMS];
        u32 part[BK_ITEMS];
#pragma unroll
        for (int j = 0; j < BK_ITEMS; j++) {
            const u32 i = j * BK_THREADS + tid;
            if (i < cnt) {
                k[j] = ld_evict_first_u64(in + base + i, policy);
                part[j] = key_part(k[j], nparts, cohash_l);
                atomicAdd(&hist[part[j]], 1u);
            }
        }
        __syncthreads();
        {   // exclusive scan of the 256 bins: shuffle scan per warp, then warp totals
            const u32 v = hist[tid];
            u32 inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            u32 woff = 0;
#pragma unroll
            for (int w = 0; w < BK_THREADS / 32; w++)
                if (w < warp) woff += wsum[w];
            loff[tid] = woff + inc - v;
            gbase[tid] = v ? atomicAdd((unsigned long long *)(cursors + tid), (unsigned long long)v) : 0ull;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < BK_ITEMS; j++) {
            const u32 i = j * BK_THREADS + tid;
            if (i < cnt) stage[loff[part[j]] + atomicAdd(&fill[part[j]], 1u)] = k[j];
        }
        __syncthreads();
        for (u32 i = tid; i < cnt; i += BK_THREADS) {
            const u64 key = stage[i];
            const u32 pp = key_part(key, nparts, cohash_l);
            const u64 at = gbase[pp] + (i - loff[pp]);
            if (at < part_cap) out[(u64)pp * part_cap + at] = key;
            else over = true;
        }
        __syncthreads();
    }
    if (over) atomicOr((unsigned long long *)flags, 1ull);
}

int dist_block_keys(euler_ctx *ctx, const u64 *d_keys, u64 n, u32 nparts, u64 *d_cursors, u64 *d_out, u64 part_cap, u32 cohash_l,
                    u64 *d_flags)
{
    if (!n) return EULER_OK;
    if (nparts > BK_MAXP) return euler_fail(ctx, EULER_ERR_ARG, "too many table parts");
    u64 grid = (u64)ctx->num_sms * 4;
    const u64 need = (n + BK_TILE - 1) / BK_TILE;
    if (grid > need) grid = need;
    dist_block_keys_kernel<<<(unsigned)grid, BK_THREADS, 0, ctx->stream>>>(d_keys, n, nparts, d_cursors, d_out, part_cap, cohash_l, d_flags);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- ownership-aware graph stage ----------------------------------------------------------------
__global__ void __launch_bounds__(DB) dist_vertex_insert_kernel(const u64 *__restrict__ lt_keys, u64 lt_cap, u32 l,
                                                                 u64 *__restrict__ vt_keys, u64 vt_cap,
                                                                 const unsigned char *__restrict__ own_flags, u64 *flags)
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
    const u32 own = own_flags[slot];
    if (own & 1u) bad |= table_insert(vt_keys, vt_cap, cp, max_probe) == EULER_NO_SLOT;
    if (own & 2u) bad |= table_insert(vt_keys, vt_cap, cs, max_probe) == EULER_NO_SLOT;
    if (bad) atomicOr((unsigned long long *)flags, 2ull);
}

int dist_vertex_insert(euler_ctx *ctx, const u64 *lt_keys, u64 lt_cap, u32 l, u64 *vt_keys, u64 vt_cap,
                       const unsigned char *own_flags, u64 *d_flags)
{
    dist_vertex_insert_kernel<<<grid_for(lt_cap, DB), DB, 0, ctx->stream>>>(lt_keys, lt_cap, l, vt_keys, vt_cap, own_flags,
                                                                            d_flags);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// Ownership of the slot's canonical l-mer c, computed once (one pass over its m-mers) and cached:
// bit0 = prefix(c) is ours (strand c is homed here), bit1 = suffix(c) is ours (strand rc(c) is homed here)
struct DistLtScanPolicy {
    typedef u64 T;
    const u64 *keys;
    const u32 *cnt;
    u32 l, rank, nranks;
    u32 *base, *eoff;
    unsigned char *own_flags;
    __device__ __forceinline__ u64 load(u64 i) const
    {
        const u64 c = keys[i];
        if (c == EULER_EMPTY_KEY) { own_flags[i] = 0; return 0ull; }
        u32 op, os;
        owners_of_lmer(c, l, nranks, op, os);
        const bool pal = c == revcomp64(c, l);
        const u32 own = (op == rank ? 1u : 0u) | (os == rank ? 2u : 0u);
        own_flags[i] = (unsigned char)own;
        const u64 n = cnt[i];
        const u64 records = pal ? (own & 1u) : ((own & 1u) + ((own >> 1) & 1u));
        const u64 edges = pal ? ((own & 1u) ? 2 * n : 0) : n * records;
        return (edges << 32) | records;
    }
    __device__ __forceinline__ void store(u64 i, u64 ex, u64, bool valid) const
    {
        if (valid) { base[i] = (u32)ex; eoff[i] = (u32)(ex >> 32); }
    }
};

int dist_lt_scan(euler_ctx *ctx, const u64 *lt_keys, const u32 *lt_cnt, u64 cap, u32 l, u32 rank, u32 nranks, u32 *base,
                 u32 *eoff, unsigned char *own_flags, u64 *d_total_packed)
{
    return scan_run(ctx, DistLtScanPolicy{lt_keys, lt_cnt, l, rank, nranks, base, eoff, own_flags}, cap, d_total_packed);
}

__global__ void __launch_bounds__(DB) dist_edges_kernel(const u64 *__restrict__ lt_keys, const u32 *__restrict__ lt_cnt,
                                                         const u32 *__restrict__ base, const u32 *__restrict__ eoff, u64 cap, u32 l,
                                                         VertexTable vt, const unsigned char *__restrict__ own_flags,
                                                         u64 *__restrict__ lkeys,
                                                         u32 *__restrict__ lvals, u32 *__restrict__ loffs, u32 *__restrict__ ev1,
                                                         u32 *__restrict__ ev2, DegOut dg)
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
    const u32 own = own_flags[slot];
    const bool own_p = (own & 1u) != 0, own_s = (own & 2u) != 0;
    u32 id_p = EULER_NO_ID, id_rp = EULER_NO_ID, id_s = EULER_NO_ID, id_rs = EULER_NO_ID;
    if (own_p) {
        const u32 a = table_find_id(vt.keys, vt.cap, cp, hash_bucket(cp, (u32)(vt.cap / EULER_BUCKET)), vt.bbase, k);
        if (a == EULER_NO_ID) return;
        const u32 b = (p == rp) ? a : a + 1u;
        id_p = (p == cp) ? a : b;
        id_rp = (p == cp) ? b : a;
    }
    if (own_s) {
        const u32 a = table_find_id(vt.keys, vt.cap, cs, hash_bucket(cs, (u32)(vt.cap / EULER_BUCKET)), vt.bbase, k);
        if (a == EULER_NO_ID) return;
        const u32 b = (s == rs) ? a : a + 1u;
        id_s = (s == cs) ? a : b;
        id_rs = (s == cs) ? b : a;
    }
    const u32 m0 = pal ? 2u * n : n;
    u32 idx = base[slot], eo = eoff[slot];
    const u32 first_c = (u32)((c >> (2 * k)) & 3), first_r = (u32)((r >> (2 * k)) & 3);
    if (own_p) {  // strand c is homed here (leaves p); rc(c) enters rc(p), which is ours too
        lkeys[idx] = c; lvals[idx] = m0; loffs[idx] = eo; ev1[idx] = id_p; ev2[idx] = id_s;
        deg_put_l(dg, id_p, (u32)(c & 3), m0);
        if (!pal) deg_put_e(dg, id_rp, id_p, first_r, n);
        idx++; eo += m0;
    }
    if (own_s) {  // c enters s (ours); rc(c) leaves rc(s) and is homed here
        deg_put_e(dg, id_s, id_rs, first_c, m0);
        if (!pal) {
            lkeys[idx] = r; lvals[idx] = n; loffs[idx] = eo; ev1[idx] = id_rs; ev2[idx] = id_rp;
            deg_put_l(dg, id_rs, (u32)(r & 3), n);
        }
    }
}

int dist_edges(euler_ctx *ctx, const u64 *lt_keys, const u32 *lt_cnt, const u32 *base, const u32 *eoff, u64 cap, u32 l,
               const VertexTable &vt, const unsigned char *own_flags, u64 *lkeys, u32 *lvals, u32 *loffs, u32 *ev1, u32 *ev2,
               u32 *lcount, u32 *ecount, u32 *deg)
{
    dist_edges_kernel<<<grid_for(cap, DB), DB, 0, ctx->stream>>>(lt_keys, lt_cnt, base, eoff, cap, l, vt, own_flags, lkeys, lvals,
                                                                 loffs, ev1, ev2, DegOut{lcount, ecount, deg});
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}
