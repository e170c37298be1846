// bucket.cuh -- minimizer-bucketed hot path: record format and the per-lane logic of the partition pass.
//
// Replaces what eulercuda.readLmersKmersCuda (eulercuda.py:74-180) + pygpuhash.create_hash_table_device
// (pygpuhash.py:262-315) + pydebruijn.construct_debruijn_graph_device (pydebruijn.py:516-620) compute, with
// two passes instead of random probes into tables far larger than L2:
//
//   pass 1 (bucket_part.cu)  one sweep over the ASCII reads.  Every vertex (k-mer, k = l-1) belongs to the
//          bucket of its MINIMIZER (smallest scrambled canonical m-mer, strand symmetric).  Consecutive
//          k-mers of a read share their minimizer ~(k-m+2)/2 positions in a row; such a run, extended by one
//          base on each side, is cut out of the read and written 2-bit packed as ONE 16-byte record into
//          the bucket's region.  The record spells every l-mer incident to the run's vertices, so after
//          this pass a bucket holds, with full multiplicity, exactly the l-mers incident to the vertices it
//          owns (an l-mer whose prefix and suffix vertex live in different buckets is present in both).
//   pass 2 (bucket_build.cu) one CTA per bucket: count the l-mers of the bucket's records in a SHARED-MEMORY
//          table, derive the owned vertices and their degree slots there, and write the bucket's part of
//          every graph artefact sequentially (positions from a decoupled look-back over the buckets).
//
// The same records are the multi-GPU exchange format: bucket -> owning rank, and the region of a
// (bucket, source rank) pair lives in the owner's memory (written over NVLink by pass 1).
//
// Everything in this header is __host__ __device__ so that tests/host harnesses can run the lane logic on the CPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BK_HD __host__ __device__ __forceinline__
#else
#define BK_HD static inline
#endif

typedef unsigned long long bk_u64;
typedef unsigned int bk_u32;

#define BK_M 12            // minimizer length for k >= 12 (k < 12: m = k, every k-mer is its own minimizer)
#define BK_MAX_BASES 48    // bases per record (3 words of 16)
#define BK_HDR_LFF 0x40u   // the first l-mer's PREFIX vertex belongs to another bucket
#define BK_HDR_RFF 0x80u   // the last l-mer's SUFFIX vertex belongs to another bucket

struct BkRec {   // one 16-byte record
    bk_u32 hdr;   // bits 0..5 number of bases n (l <= n <= 48), bit 6 LFF, bit 7 RFF
    bk_u32 d[3];  // bases, 2 bits each, first base in bits 31:30 of d[0]
};

// ---- portable bit helpers ---------------------------------------------------------------------------
BK_HD bk_u32 bk_brev32(bk_u32 x)
{
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
BK_HD bk_u64 bk_brev64(bk_u64 x)
{
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    return ((bk_u64)bk_brev32((bk_u32)x) << 32) | bk_brev32((bk_u32)(x >> 32));
#endif
}
BK_HD int bk_ffs(bk_u32 x)   // 1-based index of the lowest set bit, 0 if none
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)x);
#else
    return x ? __builtin_ctz(x) + 1 : 0;
#endif
}
BK_HD int bk_popc(bk_u32 x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

// reverse complement of `len` bases packed MSB-first in the low 2*len bits
BK_HD bk_u64 bk_revcomp(bk_u64 x, bk_u32 len)
{
    bk_u64 y = bk_brev64(~x);
    y = ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);
    return y >> (64 - 2 * len);
}

// ---- minimizer score and bucket -----------------------------------------------------------------------
// Same scramble as the multi-GPU ownership rule of round 1 (common.cuh mmer_score): a bijection of the
// m-mer, ordered mostly by its high bits, with well mixed low bits.
BK_HD bk_u32 bk_mmer_score(bk_u32 canon_m)
{
    const bk_u32 s = canon_m * 2654435761u;
    return s ^ (s >> 15);
}
BK_HD bk_u32 bk_m_of(bk_u32 k) { return k < BK_M ? k : BK_M; }

// bucket of a vertex from its minimizer score.  rank = the round-1 owner rule (low 16 bits scaled to the rank
// count); the bucket inside the rank comes from a full remix of the score (the minimum of many scores hugs 0
// in its high bits, a bijective remix of the m-mer identity is uniform again).
struct BkGeom {
    bk_u32 nranks;        // >= 1
    bk_u32 nb_per_rank;   // buckets owned by each rank
};
BK_HD bk_u32 bk_fmix32(bk_u32 h)
{
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}
BK_HD bk_u32 bk_rank_of(bk_u32 score, bk_u32 nranks) { return ((score & 0xffffu) * nranks) >> 16; }
BK_HD bk_u32 bk_bucket_of(bk_u32 score, BkGeom g)
{
    const bk_u32 lb = (bk_u32)(((bk_u64)bk_fmix32(score) * g.nb_per_rank) >> 32);
    return bk_rank_of(score, g.nranks) * g.nb_per_rank + lb;
}

// ---- window validity ----------------------------------------------------------------------------------
// Bit masks over the 48 bases [chunk start - 32, chunk start + 16): base q sits at bit 63 - q.
// bk_run_and(A, len): bit b = AND of A[b .. b+len-1] (zeros enter from the top: no context = invalid).
BK_HD bk_u64 bk_run_and(bk_u64 A, bk_u32 len)
{
    if (len == 0) return ~0ull;
    bk_u32 have = 1;
    bk_u64 R = A;
    while (have * 2 <= len) {
        R &= R >> have;
        have *= 2;
    }
    if (have < len) R &= R >> (len - have);
    return R;
}
// valid k-mer windows: all k bases are ACGT and no read starts after the window's first base
BK_HD bk_u64 bk_valid_kmers(bk_u64 vmw, bk_u64 smw, bk_u32 k) { return bk_run_and(vmw, k) & bk_run_and(~smw, k - 1); }

// valid l-mer windows (l = k + 1) from the valid k-mer windows: two consecutive valid k-mers overlap in k - 1 >= 1
// bases, so they lie in one read; for k = 1 nothing overlaps and the read start has to be excluded explicitly
BK_HD bk_u64 bk_valid_lmers(bk_u64 VK, bk_u64 smw, bk_u32 k) { return VK & (VK >> 1) & (k == 1 ? ~smw : ~0ull); }

// own-position masks (bit i = position i of the lane's chunk) from a 64-bit window mask
BK_HD bk_u32 bk_own16(bk_u64 w) { return bk_brev32((bk_u32)w) & 0xffffu; }

// ---- m-mer scores of the 16 positions of a chunk --------------------------------------------------------
// c1 = codes of the previous chunk, c0 = codes of this chunk (16 bases each, first base in bits 31:30).
// sc[i] = score of the canonical m-mer ENDING at position i (garbage when it reaches before c1: never used).
BK_HD void bk_chunk_scores(bk_u32 c1, bk_u32 c0, bk_u32 m, bk_u32 (&sc)[16])
{
    const bk_u64 cw = ((bk_u64)c1 << 32) | c0;
    const bk_u64 rcw = bk_revcomp(cw, 32);
    const bk_u32 mmask = m >= 16 ? 0xffffffffu : ((1u << (2 * m)) - 1u);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 16; i++) {
        const bk_u32 fm = (bk_u32)(cw >> (2 * (15 - i))) & mmask;
        const bk_u32 rm = (bk_u32)(rcw >> (2 * (17 + i - (int)m))) & mmask;
        sc[i] = bk_mmer_score(fm < rm ? fm : rm);
    }
}

// ---- record assembly ----------------------------------------------------------------------------------
// c2, c1, c0: codes of the two previous chunks and this chunk = bases q = 0..47; the record takes the n
// bases starting at base q0.
BK_HD BkRec bk_make_record(bk_u32 c2, bk_u32 c1, bk_u32 c0, bk_u32 q0, bk_u32 n, bk_u32 flags)
{
    const bk_u64 A = ((bk_u64)c2 << 32) | c1, B = (bk_u64)c0 << 32;
    const bk_u32 sh = 2 * q0;
    bk_u64 hi, lo;
    if (sh == 0) { hi = A; lo = B; }
    else if (sh < 64) { hi = (A << sh) | (B >> (64 - sh)); lo = B << sh; }
    else { hi = B << (sh - 64); lo = 0; }
    BkRec r;
    r.hdr = n | flags;
    r.d[0] = (bk_u32)(hi >> 32);
    r.d[1] = (bk_u32)hi;
    r.d[2] = (bk_u32)(lo >> 32);
    return r;
}

// l-mer j of a record (bases j .. j+l-1), 0 <= j <= 16, l <= 32
BK_HD bk_u64 bk_record_lmer(const BkRec &r, bk_u32 j, bk_u32 l)
{
    const bk_u64 hi = ((bk_u64)r.d[0] << 32) | r.d[1], lo = (bk_u64)r.d[2] << 32;
    const bk_u32 sh = 2 * j;
    const bk_u64 f = sh ? ((hi << sh) | (lo >> (64 - sh))) : hi;
    return f >> (64 - 2 * l);
}

// ---- the pieces of one lane -----------------------------------------------------------------------------
// win[i]      minimizer score of the k-mer ending at position i of this lane's chunk (any addressable memory)
// win_prev    the same for the last position of the previous chunk
// eq16        bit i: win[i] equals the score of the position before (bk_eq16)
// vk16, vl16  valid k-mer / l-mer windows ending at the lane's positions (bit i = position i)
// emit(bucket, record) is called once per record.  A piece never crosses the lane's chunk: the l-mer that
// ends at position 0 is shipped by THIS lane (left flank of its first piece; plus an orphan single-l-mer record
// to the previous chunk's bucket when that differs).
BK_HD bk_u32 bk_eq16(const bk_u32 (&win)[16], bk_u32 win_prev)
{
    bk_u32 eq16 = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 16; i++) eq16 |= (win[i] == (i ? win[i - 1] : win_prev) ? 1u : 0u) << i;
    return eq16;
}
// The records of ONE piece [s, e] (positions of the lane's chunk): the piece itself and, for a piece that starts the
// chunk right after a k-mer of another bucket, the orphan single-l-mer record for that bucket.
//   w_s = minimizer score at s, w_before = at s - 1 (the previous chunk's last position when s = 0),
//   w_after = at e + 1 (ignored when e = 15)
template <typename Emit>
BK_HD void bk_piece_records(bk_u32 c2, bk_u32 c1, bk_u32 c0, bk_u32 s, bk_u32 e, bk_u32 w_s, bk_u32 w_before, bk_u32 w_after,
                            bk_u32 vl16, bk_u32 k, BkGeom g, Emit emit)
{
    const bk_u32 b = bk_bucket_of(w_s, g);
    bk_u32 flags = 0, lf = 0, rf = 0;
    if ((vl16 >> s) & 1u) {   // the l-mer ending at s exists: it is this piece's first l-mer
        lf = 1;
        const bk_u32 bp = bk_bucket_of(w_before, g);
        if (bp != b) {
            flags |= BK_HDR_LFF;
            if (s == 0)   // the previous chunk's piece could not look ahead: ship the l-mer to its bucket from here
                emit(bp, bk_make_record(c2, c1, c0, 32u - k, k + 1u, BK_HDR_RFF));
        }
    }
    if (e < 15u && ((vl16 >> (e + 1u)) & 1u)) {   // the l-mer ending at e+1: ours as well when its suffix vertex is foreign
        if (bk_bucket_of(w_after, g) != b) {
            rf = 1;
            flags |= BK_HDR_RFF;
        }
    }
    if (e - s + lf + rf == 0) return;   // a lone k-mer with no l-mer around it is no vertex of the graph
    const bk_u32 q0 = 32u + s - (k - 1u) - lf, q1 = 32u + e + rf;
    emit(b, bk_make_record(c2, c1, c0, q0, q1 - q0 + 1u, flags));
}

// piece boundaries of a lane: bit i of `starts` / `ends` = a piece starts / ends at position i (the j-th start pairs
// with the j-th end).  A piece = consecutive valid k-mers of one read with the same minimizer score, inside the chunk.
BK_HD void bk_piece_masks(bk_u32 eq16, bk_u32 vk16, bk_u32 vl16, bk_u32 &starts, bk_u32 &ends)
{
    const bk_u32 cont = vl16 & eq16 & 0xfffeu;   // position i continues the piece of i-1: same read, same minimizer (never across the lane start)
    starts = vk16 & ~cont;
    ends = vk16 & ~(cont >> 1);
}

template <typename Emit>
BK_HD void bk_lane_pieces(bk_u32 c2, bk_u32 c1, bk_u32 c0, const bk_u32 *win, bk_u32 win_prev, bk_u32 eq16, bk_u32 vk16,
                          bk_u32 vl16, bk_u32 k, BkGeom g, Emit emit)
{
    if (!vk16) return;
    bk_u32 starts, ends;
    bk_piece_masks(eq16, vk16, vl16, starts, ends);
    while (starts) {
        const bk_u32 s = (bk_u32)bk_ffs(starts) - 1u, e = (bk_u32)bk_ffs(ends) - 1u;
        starts &= starts - 1u;
        ends &= ends - 1u;
        bk_piece_records(c2, c1, c0, s, e, win[s], s ? win[s - 1] : win_prev, e < 15u ? win[e + 1u] : 0u, vl16, k, g, emit);
    }
}

// ---- sliding-window minimum -----------------------------------------------------------------------------
// sa[0..19] = scores of the 20 positions before the chunk, sa[20..35] = the chunk's own 16.
// win[i] = min of the W scores ending at position i  (W = k - m + 1 m-mers per k-mer, 1 <= W <= 20).
// van Herk / Gil-Werman: prefix and suffix minima inside blocks of W, two lookups per window; every index
// is a compile-time constant after unrolling.
template <int W>
BK_HD void bk_window_min(const bk_u32 (&sa)[36], bk_u32 (&win)[16])
{
    constexpr int O = 20 - (W - 1), N = W - 1 + 16;
    bk_u32 pm[N], sm[N];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int t = 0; t < N; t++) {
        const bk_u32 v = sa[O + t];
        pm[t] = (t % W == 0) ? v : (pm[t - 1] < v ? pm[t - 1] : v);
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int t = N - 1; t >= 0; t--) {
        const bk_u32 v = sa[O + t];
        sm[t] = (t % W == W - 1 || t == N - 1) ? v : (sm[t + 1] < v ? sm[t + 1] : v);
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 16; i++) win[i] = sm[i] < pm[i + W - 1] ? sm[i] : pm[i + W - 1];
}
// any W at run time (odd k: tests and small inputs)
BK_HD void bk_window_min_any(const bk_u32 (&sa)[36], bk_u32 W, bk_u32 (&win)[16])
{
    for (int i = 0; i < 16; i++) {
        bk_u32 v = sa[20 + i];
        for (bk_u32 j = 1; j < W; j++) v = sa[20 + i - j] < v ? sa[20 + i - j] : v;
        win[i] = v;
    }
}
