// bucket_part.cu -- pass 1 of the bucketed hot path (bucket.cuh): one sweep over the ASCII reads that cuts
// every read into minimizer runs and writes them, 2-bit packed, as 16-byte records into the bucket regions.
// Replaces the encode half of eulercuda.readLmersKmersCuda (eulercuda.py:117-139: encode_lmer_device,
// compute_kmer_device, compute_lmer_complement_device) -- no per-position key ever reaches memory.
//
// Tile geometry is the encoder's (encode.cuh): a warp covers 32 chunks of 16 bytes, lanes [0, ENC_HALO) only
// supply left context.  Per lane: 16 m-mer scores (two funnel shifts + min + scramble each, no rolling
// state), the 20 scores before them from the warp's shared-memory rows (stride 20 words: conflict-free 128-bit
// accesses), a van Herk / Gil-Werman sliding minimum (~4 min per position), validity of all 16 k-mer / l-mer
// windows from two 64-bit masks and the piece boundaries (bit masks).  The tile's pieces are then listed in shared
// memory and handled one per lane per round -- every lane busy, 32 cursor atomics in flight -- each giving one
// 16-byte record: a cursor atomic and one 16-byte store.
#include "bucket.cuh"
#include "encode.cuh"
#ifndef EULER_SIMT_EMU   // (tests/host/simt_part_check.cpp compiles the kernels of this file for the CPU)
#include "kernels.h"
#endif

#define BP_BLOCK 128
#define BP_WARPS (BP_BLOCK / 32)
#define BP_ROW 20                       // words per lane row (16 scores + 4 pad)
#define BP_ROWS 34                      // two rows of padding in front of lane 0
#ifndef BP_MINB
#define BP_MINB 8
#endif

// STREAM = false: records go straight into the bucket regions (one GPU: `dst[0]`, region of bucket b at b * rcap, one
//   cursor per bucket).
// STREAM = true (multi-GPU): records go into ONE stream per destination rank (`dst[r]` + my_rank * rcap, rcap = stream
//   capacity, one cursor per destination), the local bucket id rides in bits 8..31 of the header, and the owner regroups
//   its incoming streams into bucket regions itself (bkt_regroup_kernel).  Isolated 16-byte stores into thousands of
//   regions of a PEER's memory ran at ~1 GB/s on 8 GPUs; here every tile first reserves one contiguous run per
//   destination (one cursor atomic each), so what crosses NVLink are runs of records.
template <int W, bool STREAM>
__global__ void __launch_bounds__(BP_BLOCK, BP_MINB) bkt_partition_kernel(const uint4 *__restrict__ buf16, u64 n_bases,
                                                                           const u32 *__restrict__ start_bits, u32 l, BkGeom geom,
                                                                           u32 my_rank, u32 rcap, uint4 *const *__restrict__ dst,
                                                                           u32 *__restrict__ cursors, u64 ntiles, u64 *__restrict__ stats)
{
    // stream form, per warp and destination rank: pieces of the tile (s_rmain), slots reserved for orphans (s_rorph), orphans
    // written (s_rfill), first piece of the rank in the tile's rank-sorted piece list (s_roff, s_rpos = next free), and the
    // start of the run reserved in the destination's stream (s_rbase)
    __shared__ u32 s_rmain[STREAM ? BP_WARPS : 1][16], s_rorph[STREAM ? BP_WARPS : 1][16], s_rfill[STREAM ? BP_WARPS : 1][16];
    __shared__ u32 s_roff[STREAM ? BP_WARPS : 1][16], s_rpos[STREAM ? BP_WARPS : 1][16], s_rbase[STREAM ? BP_WARPS : 1][16];
    __shared__ unsigned short s_sorted[STREAM ? BP_WARPS : 1][STREAM ? ENC_ADV * 16 : 1];   // the tile's pieces ordered by destination rank
    __shared__ __align__(16) u32 s_rows[BP_WARPS][BP_ROWS * BP_ROW];   // per lane: 16 m-mer scores, then 16 window minima
    __shared__ u32 s_codes[BP_WARPS][34];                              // 2-bit codes of every lane's chunk (two pad entries in front)
    __shared__ u32 s_vl[BP_WARPS][32];                                 // valid l-mer windows of every lane
    __shared__ unsigned short s_desc[BP_WARPS][ENC_ADV * 16];          // the tile's pieces: lane << 8 | end << 4 | start
    const int wib = threadIdx.x >> 5;
    u32 *rows = s_rows[wib];
    const int lane = threadIdx.x & 31;
    u32 *my_row = rows + (lane + 2) * BP_ROW;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    // W > 0 fixes k at compile time (m = 12, k = W + 11): the validity masks and shifts below unroll to constants
    const u32 k = W > 0 ? (u32)(W + BK_M - 1) : l - 1, m = bk_m_of(k);
    u32 nl_tot = 0, nk_tot = 0;
    bool overflow = false;
    if (lane < 2) s_codes[wib][lane] = 0;

    for (u64 tile = warp; tile < ntiles; tile += nwarps) {
        const long long chunk = (long long)(tile * ENC_ADV) - ENC_HALO + lane;
        const Chunk c = load_chunk(buf16, chunk, n_bases, start_bits);
        const u32 p1 = __shfl_up_sync(0xffffffffu, c.codes, 1);
        const u32 v1 = __shfl_up_sync(0xffffffffu, c.vmask, 1), v2 = __shfl_up_sync(0xffffffffu, c.vmask, 2);
        const u32 s1 = __shfl_up_sync(0xffffffffu, c.smask, 1), s2 = __shfl_up_sync(0xffffffffu, c.smask, 2);
        u32 sa[36];
        {
            u32 sc[16];
            bk_chunk_scores(lane >= 1 ? p1 : 0u, c.codes, m, sc);
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<uint4 *>(my_row + i) = make_uint4(sc[i], sc[i + 1], sc[i + 2], sc[i + 3]);
#pragma unroll
            for (int i = 0; i < 16; i++) sa[20 + i] = sc[i];
        }
        s_codes[wib][lane + 2] = c.codes;
        __syncwarp();
        {   // the 20 scores before this chunk: the previous lane's row and the last four of the lane before it
            const uint4 a = *reinterpret_cast<const uint4 *>(my_row - 2 * BP_ROW + 12);
            sa[0] = a.x; sa[1] = a.y; sa[2] = a.z; sa[3] = a.w;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const uint4 b = *reinterpret_cast<const uint4 *>(my_row - BP_ROW + i);
                sa[4 + i] = b.x; sa[5 + i] = b.y; sa[6 + i] = b.z; sa[7 + i] = b.w;
            }
        }
        u32 win[16];
        if constexpr (W > 0) bk_window_min<W>(sa, win);
        else bk_window_min_any(sa, k - m + 1, win);
        __syncwarp();   // every lane has read its neighbours' scores: the rows now carry the window minima
#pragma unroll
        for (int i = 0; i < 16; i += 4) *reinterpret_cast<uint4 *>(my_row + i) = make_uint4(win[i], win[i + 1], win[i + 2], win[i + 3]);
        __syncwarp();
        // ---- piece boundaries of this lane, listed for the whole warp ---------------------------------------------------
        u32 starts = 0, ends = 0, vl16 = 0;
        if (lane >= ENC_HALO) {
            const u64 vmw = ((u64)v2 << 48) | ((u64)v1 << 32) | ((u64)c.vmask << 16);
            const u64 smw = ((u64)s2 << 48) | ((u64)s1 << 32) | ((u64)c.smask << 16);
            const u64 VK = bk_valid_kmers(vmw, smw, k);
            const u32 vk16 = bk_own16(VK);
            vl16 = bk_own16(bk_valid_lmers(VK, smw, k));
            nk_tot += __popc(vk16);
            nl_tot += __popc(vl16);
            bk_piece_masks(bk_eq16(win, my_row[-BP_ROW + 15]), vk16, vl16, starts, ends);
        }
        s_vl[wib][lane] = vl16;
        const u32 np = __popc(starts);
        u32 inc = np;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        const u32 total = __shfl_sync(0xffffffffu, inc, 31);
        u32 at = inc - np;
        while (starts) {
            const u32 s = (u32)__ffs(starts) - 1u, e = (u32)__ffs(ends) - 1u;
            starts &= starts - 1u;
            ends &= ends - 1u;
            s_desc[wib][at++] = (unsigned short)((lane << 8) | (e << 4) | s);
        }
        __syncwarp();
        if constexpr (STREAM) {
            // ---- one run per destination rank for this tile: its pieces first, in the order of the rank-sorted piece list (so
            // that the lanes of one store instruction write CONSECUTIVE records of a stream: on 8 GPUs the partition pass is
            // bound by the number of remote write transactions, not by their bytes), then a slot for every orphan a
            // chunk-leading piece MAY add
            if (lane < 16) { s_rmain[wib][lane] = 0; s_rorph[wib][lane] = 0; s_rfill[wib][lane] = 0; s_rpos[wib][lane] = 0; }
            __syncwarp();
            for (u32 t = lane; t < total; t += 32) {
                const u32 d = s_desc[wib][t];
                const u32 L = d >> 8, s = d & 15u;
                const u32 *row = rows + (L + 2) * BP_ROW;
                atomicAdd(&s_rmain[wib][bk_rank_of(row[s], geom.nranks)], 1u);
                if (s == 0 && (s_vl[wib][L] & 1u)) atomicAdd(&s_rorph[wib][bk_rank_of(row[-BP_ROW + 15], geom.nranks)], 1u);
            }
            __syncwarp();
            {
                const u32 mine = lane < 16 ? s_rmain[wib][lane] : 0u;
                u32 incl = mine;
#pragma unroll
                for (int d = 1; d < 16; d <<= 1) {
                    const u32 t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                if (lane < 16) s_roff[wib][lane] = incl - mine;
                if ((u32)lane < geom.nranks) {
                    const u32 cnt = mine + s_rorph[wib][lane];
                    s_rbase[wib][lane] = cnt ? atomicAdd(cursors + lane, cnt) : 0u;   // ONE cursor atomic per destination per tile
                }
            }
            __syncwarp();
            for (u32 t = lane; t < total; t += 32) {
                const u32 d = s_desc[wib][t];
                const u32 rank = bk_rank_of(rows[((d >> 8) + 2) * BP_ROW + (d & 15u)], geom.nranks);
                s_sorted[wib][s_roff[wib][rank] + atomicAdd(&s_rpos[wib][rank], 1u)] = (unsigned short)d;
            }
            __syncwarp();
            // ---- one piece per lane per round, in rank order: the piece's record at its place in the run, an orphan behind the pieces
            for (u32 t = lane; t < total; t += 32) {
                const u32 d = s_sorted[wib][t];
                const u32 L = d >> 8, e = (d >> 4) & 15u, s = d & 15u;
                const u32 *row = rows + (L + 2) * BP_ROW;
                const u32 w_s = row[s], w_before = s ? row[s - 1] : row[-BP_ROW + 15], w_after = e < 15u ? row[e + 1] : 0u;
                const u32 home = bk_bucket_of(w_s, geom), hrank = geom.nranks > 1 ? home / geom.nb_per_rank : 0u;
                const u32 hpos = s_rbase[wib][hrank] + (t - s_roff[wib][hrank]);
                bool placed = false;
                bk_piece_records(s_codes[wib][L], s_codes[wib][L + 1], s_codes[wib][L + 2], s, e, w_s, w_before, w_after, s_vl[wib][L], k, geom,
                                 [&](u32 bucket, const BkRec &r) {
                                     const u32 rank = geom.nranks > 1 ? bucket / geom.nb_per_rank : 0u;
                                     const u32 lb = bucket - rank * geom.nb_per_rank;
                                     u32 pos;
                                     if (bucket == home) { pos = hpos; placed = true; }   // (an orphan goes to ANOTHER bucket by construction)
                                     else pos = s_rbase[wib][rank] + s_rmain[wib][rank] + atomicAdd(&s_rfill[wib][rank], 1u);
                                     if (pos < rcap) dst[rank][(u64)my_rank * rcap + pos] = make_uint4(r.hdr | (lb << 8), r.d[0], r.d[1], r.d[2]);
                                     else overflow = true;
                                 });
                // a lone k-mer is no record: its place in the run gets an empty one (the owner skips it)
                if (!placed && hpos < rcap) dst[hrank][(u64)my_rank * rcap + hpos] = make_uint4(0u, 0u, 0u, 0u);
            }
            __syncwarp();
            // orphan slots that no orphan took (the chunk-leading piece's neighbour was in the same bucket after all): empty records
            for (u32 r = 0; r < geom.nranks; r++) {
                const u32 first = s_rbase[wib][r] + s_rmain[wib][r], fill = s_rfill[wib][r], cnt = s_rorph[wib][r];
                for (u32 j = fill + lane; j < cnt; j += 32)
                    if (first + j < rcap) dst[r][(u64)my_rank * rcap + first + j] = make_uint4(0u, 0u, 0u, 0u);
            }
        } else {
            // ---- one piece per lane per round: record(s), cursor atomic, 16-byte store ---------------------------------------
            for (u32 t = lane; t < total; t += 32) {
                const u32 d = s_desc[wib][t];
                const u32 L = d >> 8, e = (d >> 4) & 15u, s = d & 15u;
                const u32 *row = rows + (L + 2) * BP_ROW;
                const u32 w_s = row[s], w_before = s ? row[s - 1] : row[-BP_ROW + 15], w_after = e < 15u ? row[e + 1] : 0u;
                bk_piece_records(s_codes[wib][L], s_codes[wib][L + 1], s_codes[wib][L + 2], s, e, w_s, w_before, w_after, s_vl[wib][L], k, geom,
                                 [&](u32 bucket, const BkRec &r) {
                                     const u32 rank = geom.nranks > 1 ? bucket / geom.nb_per_rank : 0u;
                                     const u32 lb = bucket - rank * geom.nb_per_rank;
                                     const u32 pos = atomicAdd(cursors + bucket, 1u);
                                     if (pos < rcap) {
                                         uint4 *region = dst[rank] + ((u64)lb * geom.nranks + my_rank) * rcap;
                                         region[pos] = make_uint4(r.hdr, r.d[0], r.d[1], r.d[2]);
                                     } else {
                                         overflow = true;
                                     }
                                 });
            }
        }
        __syncwarp();   // the rows are rewritten by the next tile
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        nl_tot += __shfl_xor_sync(0xffffffffu, nl_tot, o);
        nk_tot += __shfl_xor_sync(0xffffffffu, nk_tot, o);
    }
    if (lane == 0) {
        if (nl_tot) atomicAdd(stats + 0, (u64)nl_tot);
        if (nk_tot) atomicAdd(stats + 1, (u64)nk_tot);
    }
    if (overflow) atomicOr((unsigned long long *)(stats + 2), (unsigned long long)BKT_FLAG_REGION);
}

#ifndef EULER_SIMT_EMU   // host side
static int bkt_partition_launch(euler_ctx *ctx, bool stream, const void *d_buf, u64 n_bases, const u32 *d_bits, u32 l, u32 nranks,
                                u32 nb_per_rank, u32 my_rank, u32 rcap, uint4 *const *d_dst, u32 *d_cursors, u64 *d_stats)
{
    if (!n_bases) return EULER_OK;
    const u64 nchunks = (n_bases + 15) / 16;
    const u64 ntiles = (nchunks + ENC_ADV - 1) / ENC_ADV;
    u64 grid = (u64)ctx->num_sms * BP_MINB;
    const u64 need = (ntiles + BP_WARPS - 1) / BP_WARPS;
    if (grid > need) grid = need;
    const BkGeom geom = {nranks, nb_per_rank};
    const u32 k = l - 1, W = k - bk_m_of(k) + 1;
    const unsigned g = (unsigned)grid;
#define LAUNCH_BP(WW, SS)                                                                                                        \
    bkt_partition_kernel<WW, SS><<<g, BP_BLOCK, 0, ctx->stream>>>((const uint4 *)d_buf, n_bases, d_bits, l, geom, my_rank, rcap, d_dst, \
                                                                  d_cursors, ntiles, d_stats)
    if (stream) {
        if (W == 20) LAUNCH_BP(20, true);        // k = 31
        else if (W == 10) LAUNCH_BP(10, true);   // k = 21
        else LAUNCH_BP(0, true);
    } else {
        if (W == 20) LAUNCH_BP(20, false);
        else if (W == 10) LAUNCH_BP(10, false);
        else LAUNCH_BP(0, false);
    }
#undef LAUNCH_BP
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

int bkt_partition(euler_ctx *ctx, const void *d_buf, u64 n_bases, const u32 *d_bits, u32 l, u32 nranks, u32 nb_per_rank, u32 my_rank,
                  u32 rcap, uint4 *const *d_dst, u32 *d_cursors, u64 *d_stats)
{
    return bkt_partition_launch(ctx, false, d_buf, n_bases, d_bits, l, nranks, nb_per_rank, my_rank, rcap, d_dst, d_cursors, d_stats);
}
// stream form: d_dst[r] = base of rank r's stream area (nranks streams of scap records), d_cursors[nranks] zeroed by the caller
int bkt_partition_streams(euler_ctx *ctx, const void *d_buf, u64 n_bases, const u32 *d_bits, u32 l, u32 nranks, u32 nb_per_rank,
                          u32 my_rank, u32 scap, uint4 *const *d_dst, u32 *d_cursors, u64 *d_stats)
{
    if (nranks > 16) return euler_fail(ctx, EULER_ERR_ARG, "at most 16 ranks");
    if (nb_per_rank >= (1u << 24)) return euler_fail(ctx, EULER_ERR_ARG, "at most 2^24 buckets per rank (the id rides in the record header)");
    return bkt_partition_launch(ctx, true, d_buf, n_bases, d_bits, l, nranks, nb_per_rank, my_rank, scap, d_dst, d_cursors, d_stats);
}

#endif

// ---- multi-GPU: publish how many records this rank wrote into each owner's stream ---------------------------------------
// the counts of owner d live behind its streams: u64 counts[nranks], entry `my_rank` is ours
__global__ void bkt_push_counts_kernel(const u32 *__restrict__ cursors, uint4 *const *__restrict__ dst, u64 stream_bytes, u32 nranks,
                                       u32 my_rank, u32 scap, u64 *__restrict__ max_out)
{
    const u32 d = threadIdx.x;
    if (d >= nranks) return;
    const u32 c = cursors[d];
    u64 *counts = (u64 *)((char *)dst[d] + stream_bytes);
    counts[my_rank] = c < scap ? c : scap;
    atomicMax((unsigned long long *)max_out, (unsigned long long)c);
}

#ifndef EULER_SIMT_EMU
int bkt_push_counts(euler_ctx *ctx, const u32 *d_cursors, uint4 *const *d_dst, u64 stream_bytes, u32 nranks, u32 my_rank, u32 scap, u64 *d_max)
{
    bkt_push_counts_kernel<<<1, 32, 0, ctx->stream>>>(d_cursors, d_dst, stream_bytes, nranks, my_rank, scap, d_max);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

#endif

// ---- owner side: incoming streams -> bucket regions (local memory, one cursor per bucket) ---------------------------------
__global__ void __launch_bounds__(256) bkt_regroup_kernel(const uint4 *__restrict__ streams, const u64 *__restrict__ counts, u32 nranks,
                                                          u32 scap, u32 nb, u32 rcap, uint4 *__restrict__ records, u32 *__restrict__ cursors,
                                                          u64 *__restrict__ stats)
{
    bool overflow = false, bad = false;
    for (u32 src = 0; src < nranks; src++) {
        const u64 n = counts[src] < scap ? counts[src] : scap;
        const uint4 *in = streams + (u64)src * scap;
        for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
            const uint4 r = ld_stream_v4(in + i);
            if ((r.x & 63u) == 0u) continue;   // an unused reserved slot
            const u32 lb = r.x >> 8;
            if (lb >= nb) { bad = true; continue; }
            const u32 pos = atomicAdd(cursors + lb, 1u);
            if (pos < rcap) records[(u64)lb * rcap + pos] = make_uint4(r.x & 0xffu, r.y, r.z, r.w);
            else overflow = true;
        }
    }
    if (overflow) atomicOr((unsigned long long *)(stats + 2), (unsigned long long)BKT_FLAG_REGION);
    if (bad) atomicOr((unsigned long long *)(stats + 2), (unsigned long long)BKT_FLAG_INTERNAL);
}

#ifndef EULER_SIMT_EMU
int bkt_regroup(euler_ctx *ctx, const void *d_streams, const u64 *d_counts, u32 nranks, u32 scap, u32 nb, u32 rcap, void *d_records,
                u32 *d_cursors, u64 *d_stats)
{
    bkt_regroup_kernel<<<(unsigned)ctx->num_sms * 8, 256, 0, ctx->stream>>>((const uint4 *)d_streams, d_counts, nranks, scap, nb, rcap,
                                                                          (uint4 *)d_records, d_cursors, d_stats);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}
#endif
