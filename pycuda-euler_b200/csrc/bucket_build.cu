// bucket_build.cu -- pass 2 of the bucketed hot path (bucket.cuh): one CTA per bucket builds the bucket's part
// of the de Bruijn graph entirely in shared memory and writes every artefact sequentially.
//
// Replaces, per bucket, the host dictionary fill of eulercuda.readLmersKmersCuda (eulercuda.py:141-178), the
// hash table of pygpuhash.create_hash_table_device (pygpuhash.py:262-315) and D1-D6 of
// pydebruijn.construct_debruijn_graph_device (pydebruijn.py:516-620):
//   A  count the canonical l-mers spelled by the bucket's records.  Every warp streams its own slice of the
//      records (two batches of 32 in registers, handed to the lanes by shuffle), every lane rolls the forward and
//      reverse-complement l-mer of its record one base per step, and a lane that finishes takes the warp's next
//      record.  Shared-memory tables with ONE key per slot and linear probing: a key that sits in its home slot costs
//      one 8-byte load and one shared atomicAdd; first occurrences and displaced keys go through a per-warp work queue
//      to the atomicCAS path, 32 at a time.  The count word also carries the "prefix / suffix vertex is ours" bits.
//   B  insert the owned end vertices of every distinct l-mer into a second shared-memory table and add the
//      multiplicity to the vertex's leaving / entering total; note which of its eight degree slots are non-zero;
//   C  totals of (edge records, vertices, edge multiplicities) over the slots;
//   D  decoupled look-back over the buckets in the order in which they finish counting (the output ticket is taken
//      when a bucket's totals are ready, so no bucket ever waits for another one's work): global bases;
//   E  vertex artefacts (the non-zero degree slots of a vertex are look-ups in the bucket's own l-mer table)
//      and edge artefacts, written in slot order with warp-row scans: consecutive lanes write consecutive ids.
//      The suffix vertex of an edge whose suffix lives in another bucket is resolved afterwards by two light passes
//      over the records through a small global table keyed by the canonical l-mer (bkt_boundary_publish_kernel,
//      bkt_fixup_kernel).
// Every sparse phase (B, E1, E2: ~30 % of the slots are occupied) compacts its work through the warp queue as well.
// A bucket several times the mean size would fill its table: the first pass gives it up after BB_PROBE_LIMIT probes
// and lists it, and a second launch of the same kernel rebuilds the listed buckets with the largest tables a block
// can hold (BKT_MAX_CAP slots), appending their artefacts -- ids are in completion order anyway.
//
// Warp-uniformity is a correctness matter here (shuffles and votes inside data-dependent loops): nothing that steers
// those loops may be read from memory another warp writes.  tests/host/simt_build_check.cpp runs this file under a
// SIMT emulator that aborts when the lanes of a warp meet at different collectives.
#include "bucket.cuh"
#ifdef EULER_SIMT_EMU   // tests/host/simt_build_check.cpp compiles the kernels of this file for the CPU (tests/host/simt_emu.h)
#include "simt_emu.h"
#else
#include "kernels.h"
#endif

#define BB_THREADS 256
#ifndef BB_PROBE_LIMIT
#define BB_PROBE_LIMIT 128u   // probes after which the first pass hands a bucket to the second
#endif
#define BB_WARPS (BB_THREADS / 32)
#ifndef BB_STEPS
#define BB_STEPS 2   // l-mers a lane rolls between two refills
#endif
#ifndef BB_MINB
#define BB_MINB 4
#endif
// -DBKT_TIMING: thread 0 of every block adds the clock cycles of each phase to stats[16 + phase] (development aid)
#ifdef BKT_TIMING
#define BB_TICK(i)                                                                      \
    do {                                                                                \
        if (threadIdx.x == 0) {                                                         \
            const long long now_ = clock64();                                           \
            atomicAdd((unsigned long long *)(a.stats + 16 + (i)), (unsigned long long)(now_ - tick_)); \
            tick_ = now_;                                                               \
        }                                                                               \
    } while (0)
#else
#define BB_TICK(i) do { } while (0)
#endif

struct BkBuildArgs {
    const uint4 *records;
    const u32 *counts;   // [nb * nranks] records in region (bucket, source rank)
    u32 nb, nranks, rcap, l;
    u32 cap;             // slots of each shared-memory table (a multiple of 256)
    u64 *lkeys; u32 *lvals, *loffs, *ev1, *ev2; u64 ucap;
    u64 *vkeys; u32 *lcount, *ecount, *lstart, *estart; euler_vertex *ev; u64 vcap;
    u32 *flag; u64 *agg_uv, *agg_e, *inc_uv, *inc_e; u32 *ticket;
    u32 *redo;           // [0] number of buckets that did not fit the first pass's tables, [1 + i] their ids
    u32 second;          // the second pass: block i rebuilds bucket redo[1 + i] with the largest tables
    u64 *bkeys; u32 *bvals; u64 bcap;
    u64 *stats;
};

// home slot of a key in a shared-memory table of `cap` slots (linear probing, one key per slot: in shared memory a probe
// is one 8-byte load, and the 32-byte buckets that pay off against DRAM sectors only cost compares and bank conflicts)
__device__ __forceinline__ u32 bb_home(u64 key, u32 cap)
{
    u32 h = (u32)key * 0x9E3779B1u + (u32)(key >> 32) * 0x85EBCA77u;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 13;
    return __umulhi(h, cap);
}
#ifndef EULER_SIMT_EMU   // (the emulator header brings its own)
__device__ __forceinline__ u32 ld_vol_u32(const u32 *p)
{
    u32 v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ u64 ld_vol_u64(const u64 *p)
{
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_vol_u32(u32 *p, u32 v) { asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_vol_u64(u64 *p, u64 v) { asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
#endif

// Slots are never freed, so the first EMPTY slot ends a search.  One exit per function: with early returns the compiler
// duplicated the caller's tail per return point and the warp ran it once per group of lanes.
// insert: returns the slot, 0xffffffff when `limit` probes did not find the key or a free slot (a table that is full,
// or so full that linear probing has degenerated: the bucket then goes to the second pass); first = this call claimed
// the slot.
__device__ __forceinline__ u32 sm_insert(u64 *keys, u32 cap, u64 key, bool &first, u32 limit)
{
    u32 h = bb_home(key, cap), slot = 0xffffffffu;
    first = false;
    for (u32 probe = 0; probe < limit; probe++) {
        u64 cur = keys[h];
        if (cur == EULER_EMPTY_KEY) {
            cur = atomicCAS(keys + h, EULER_EMPTY_KEY, key);
            first = cur == EULER_EMPTY_KEY;
            if (first) cur = key;
        }
        if (cur == key) {
            slot = h;
            break;
        }
        h = h + 1u == cap ? 0u : h + 1u;
    }
    return slot;
}
__device__ __forceinline__ u32 sm_find(const u64 *keys, u32 cap, u64 key)
{
    u32 h = bb_home(key, cap), slot = 0xffffffffu;
    for (u32 probe = 0; probe < cap; probe++) {
        const u64 cur = keys[h];
        if (cur == key || cur == EULER_EMPTY_KEY) {
            slot = cur == key ? h : 0xffffffffu;
            break;
        }
        h = h + 1u == cap ? 0u : h + 1u;
    }
    return slot;
}

__device__ __forceinline__ u32 warp_incl_u32(u32 v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ u64 warp_incl_u64(u64 v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ u64 warp_sum_u64(u64 v)
{
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// both-strand multiplicity of the l-mer x (reverse complement r) in the bucket's table (0 when absent)
__device__ __forceinline__ u32 bb_bs_count(const u64 *lt_keys, const u32 *lt_cnt, u32 cap, u64 x, u64 r)
{
    const u32 slot = sm_find(lt_keys, cap, x < r ? x : r);
    const u32 n = slot == 0xffffffffu ? 0u : (lt_cnt[slot] & 0x3fffffffu);
    return x == r ? 2u * n : n;
}

// what one l-mer slot contributes: edge records (0..2) and their multiplicity total
struct LtSlot {
    u64 c;
    u32 n, recs;
    bool own_p, own_s, pal;
    u64 edges;
};
__device__ __forceinline__ LtSlot bb_lt_slot(const u64 *lt_keys, const u32 *lt_cnt, u32 slot, u32 l)
{
    LtSlot s;
    s.c = lt_keys[slot];
    s.n = 0; s.recs = 0; s.own_p = s.own_s = s.pal = false; s.edges = 0;
    if (s.c == EULER_EMPTY_KEY) return s;
    const u32 w = lt_cnt[slot];
    s.n = w & 0x3fffffffu;
    s.own_p = (w >> 30) & 1u;
    s.own_s = (w >> 31) & 1u;
    s.pal = !(l & 1u) && s.c == bk_revcomp(s.c, l);   // only an even length can be its own reverse complement
    if (s.pal) { s.recs = s.own_p ? 1u : 0u; s.edges = s.own_p ? 2ull * s.n : 0ull; }
    else { s.recs = (s.own_p ? 1u : 0u) + (s.own_s ? 1u : 0u); s.edges = (u64)s.n * s.recs; }
    return s;
}

// Per-warp work queue in shared memory (64 entries of 16 bytes).  The per-slot phases walk table slots that are ~45 %
// occupied, and the count phase has a rare slow path; instead of letting the idle lanes ride along, a lane with work
// pushes it and the warp pops 32 entries at a time: every lane busy on the expensive code.
struct WarpQueue {
    ulonglong2 *q;   // this warp's 64 entries
    u32 n;           // warp-uniform
    __device__ __forceinline__ void push(bool pred, ulonglong2 v, unsigned lt_mask)
    {
        const unsigned m = __ballot_sync(0xffffffffu, pred);
        if (pred) q[n + __popc(m & lt_mask)] = v;
        n += __popc(m);
        __syncwarp();
    }
    // pops 32 entries (the caller checked n >= 32)
    __device__ __forceinline__ ulonglong2 pop32(int lane)
    {
        n -= 32u;
        const ulonglong2 v = q[n + lane];
        __syncwarp();
        return v;
    }
    // the rest at the end: valid for lanes < old n
    __device__ __forceinline__ ulonglong2 rest(int lane, bool &valid)
    {
        valid = (u32)lane < n;
        const ulonglong2 v = valid ? q[lane] : make_ulonglong2(0, 0);
        n = 0;
        __syncwarp();
        return v;
    }
};

// L > 0 fixes the l-mer length at compile time (the benchmark lengths 32 and 22): masks and shift counts of the
// rolling step become constants.  L == 0 reads it from the arguments.
template <int L>
__global__ void __launch_bounds__(BB_THREADS, BB_MINB) bkt_build_kernel(const BkBuildArgs a)
{
#ifdef EULER_SIMT_EMU
    unsigned char *bb_smem = simt::dyn_smem();
#else
    extern __shared__ __align__(16) unsigned char bb_smem[];
#endif
    const u32 cap = a.cap;
    u64 *lt_keys = (u64 *)bb_smem;
    u64 *vt_keys = lt_keys + cap;
    u32 *lt_cnt = (u32 *)(vt_keys + cap);
    u32 *vt_a = lt_cnt + cap;   // leaving total of the canonical strand, later the vertex id
    u32 *vt_b = vt_a + cap;     // entering total of the canonical strand
    u32 *vt_m = vt_b + cap;     // one byte per vertex slot: which of its 8 degree slots are non-zero (bits 0-3 leaving, 4-7 entering)
    __shared__ ulonglong2 s_queue[BB_WARPS][64];
    __shared__ u64 s_wtot[BB_WARPS][4];
    __shared__ u64 s_base[2];
    __shared__ u32 s_bucket, s_fail;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const u32 l = L > 0 ? (u32)L : a.l, k = l - 1;
    const u64 lmask = l >= 32 ? ~0ull : ((1ull << (2 * l)) - 1ull), kmask = lmask >> 2;
    const u32 top = 2 * (l - 1);
    WarpQueue wq;
    wq.q = s_queue[warp];
    wq.n = 0;
    // A bucket several times the mean size fills its table to the point where every insert walks hundreds of slots and
    // one block holds up the whole grid (measured: 4.9 ms instead of 1.9 ms on the rank that owned such a bucket).
    // The first pass therefore gives up on a bucket after BB_PROBE_LIMIT probes, the second pass after 8x as many
    // (its table is 4.8x larger than the default: a bucket that still fails there is repartitioned by the host).
    const u32 plimit1 = a.second ? 8u * BB_PROBE_LIMIT : BB_PROBE_LIMIT;   // the second pass tolerates a fuller table, not an endless walk per failing insert
    const u32 plimit = cap < plimit1 ? cap : plimit1;

#ifdef BKT_TIMING
    long long tick_ = clock64();
#endif
    if (tid == 0) {   // which bucket this block builds
        if (a.second) {
            const u32 nredo = a.redo[0] < BKT_REDO_CAP ? a.redo[0] : BKT_REDO_CAP;
            if (blockIdx.x == 0) a.stats[7] = a.redo[0];   // how many buckets the first pass handed over
            s_bucket = blockIdx.x < nredo ? a.redo[1 + blockIdx.x] : 0xffffffffu;
        } else {
            s_bucket = atomicAdd(a.ticket, 1u);
        }
        s_fail = 0;
    }
    __syncthreads();
    const u32 b = s_bucket;
    if (b >= a.nb) return;   // second pass: nothing (more) to redo
    {
        const ulonglong2 e2 = make_ulonglong2(EULER_EMPTY_KEY, EULER_EMPTY_KEY);
        for (u32 i = tid; i < cap; i += BB_THREADS) reinterpret_cast<ulonglong2 *>(lt_keys)[i] = e2;   // lt_keys and vt_keys are adjacent
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (u32 i = tid; i < 13 * cap / 16; i += BB_THREADS) reinterpret_cast<uint4 *>(lt_cnt)[i] = z;   // counts, vertex words, masks
    }
    __syncthreads();

    BB_TICK(0);
    // ---- A: count the l-mers of the bucket's records ---------------------------------------------------------
    bool fail = false;
    u32 max_region = 0;
    // slow path of the count: a key that is not in its home bucket yet (first occurrence, or its home was full)
    auto count_slow = [&](ulonglong2 e, bool valid) {
        if (valid) {
            bool first;
            const u32 slot = sm_insert(lt_keys, cap, e.x, first, plimit);
            if (slot == 0xffffffffu) fail = true;   // noted per lane and published after the loop: nothing warp-uniform may depend on it
            else atomicAdd(lt_cnt + slot, first ? (1u | ((u32)e.y << 30)) : 1u);
        }
    };
    for (u32 src = 0; src < a.nranks; src++) {
        u32 cnt = a.counts[(u64)b * a.nranks + src];
        max_region = cnt > max_region ? cnt : max_region;
        if (cnt > a.rcap) cnt = a.rcap;   // overflowed region: the run is repeated with a larger capacity anyway
        const uint4 *region = a.records + ((u64)b * a.nranks + src) * a.rcap;
        const u32 per = (cnt + BB_WARPS - 1) / BB_WARPS;
        u32 fetch = warp * per;
        const u32 end = fetch + per < cnt ? fetch + per : cnt;
        uint4 cur = make_uint4(0, 0, 0, 0), nxt = cur;
        u32 ncur = 0, nnxt = 0, used = 0;
        if (fetch < end) {
            ncur = end - fetch < 32u ? end - fetch : 32u;
            if ((u32)lane < ncur) cur = ld_stream_v4(region + fetch + lane);
            fetch += ncur;
        }
        if (fetch < end) {
            nnxt = end - fetch < 32u ? end - fetch : 32u;
            if ((u32)lane < nnxt) nxt = ld_stream_v4(region + fetch + lane);
            fetch += nnxt;
        }
        u64 f = 0, rc = 0, rem = 0;
        u32 left = 0, hdr = 0;
        bool firstl = false;
        while (true) {
            const bool need = left == 0;
            const unsigned nm = __ballot_sync(0xffffffffu, need);
            if (nm && ncur) {   // hand the next records of the current batch to the lanes that are done
                const u32 pos = used + __popc(nm & lt_mask);
                uint4 r;
                r.x = __shfl_sync(0xffffffffu, cur.x, pos & 31u);
                r.y = __shfl_sync(0xffffffffu, cur.y, pos & 31u);
                r.z = __shfl_sync(0xffffffffu, cur.z, pos & 31u);
                r.w = __shfl_sync(0xffffffffu, cur.w, pos & 31u);
                if (need && pos < ncur) {
                    hdr = r.x;
                    const u32 nbases = hdr & 63u;
                    left = nbases > k ? nbases - k : 0u;
                    const u64 hi = ((u64)r.y << 32) | r.z, lo = (u64)r.w << 32;
                    const u64 p = hi >> (64 - 2 * k);            // the first k bases: the state before the first l-mer
                    f = p;
                    rc = bk_revcomp(p, k) << 2;
                    rem = (hi << (2 * k)) | (lo >> (64 - 2 * k));   // the bases that follow, next one in bits 63:62
                    firstl = true;
                }
                used += __popc(nm);
                if (used >= ncur) {   // batch consumed: the prefetched one becomes current, the one after is requested
                    cur = nxt; ncur = nnxt; used = 0; nnxt = 0;
                    if (fetch < end) {
                        nnxt = end - fetch < 32u ? end - fetch : 32u;
                        if ((u32)lane < nnxt) nxt = ld_stream_v4(region + fetch + lane);
                        fetch += nnxt;
                    }
                }
            }
            if (!__any_sync(0xffffffffu, left != 0)) {
                if (ncur == 0) break;
                continue;
            }
#pragma unroll
            for (int st = 0; st < BB_STEPS; st++) {
                const bool act = left != 0;
                bool miss = false;
                ulonglong2 e = make_ulonglong2(0, 0);
                if (act) {
                    const u32 cc = (u32)(rem >> 62);
                    rem <<= 2;
                    f = ((f << 2) | cc) & lmask;
                    rc = (rc >> 2) | ((u64)(3u - cc) << top);
                    const bool flip = rc < f;
                    const u64 c = flip ? rc : f;
                    const u32 h = bb_home(c, cap);
                    if (lt_keys[h] == c) {
                        atomicAdd(lt_cnt + h, 1u);   // the common case: the key sits in its home slot
                    } else {   // first occurrence, or displaced from its home: the slow path, with the ownership of the end
                               // vertices in the orientation the read spells (bit 0: prefix(c), bit 1: suffix(c))
                        const u32 own_pf = (firstl && (hdr & BK_HDR_LFF)) ? 0u : 1u;
                        const u32 own_sf = (left == 1u && (hdr & BK_HDR_RFF)) ? 0u : 1u;
                        e = make_ulonglong2(c, flip ? (own_sf | (own_pf << 1)) : (own_pf | (own_sf << 1)));
                        miss = true;
                    }
                    firstl = false;
                    left--;
                }
                wq.push(miss, e, lt_mask);
                if (wq.n >= 32u) count_slow(wq.pop32(lane), true);
            }
        }
    }
    {
        bool valid;
        const ulonglong2 e = wq.rest(lane, valid);
        count_slow(e, valid);
    }
    if (fail) s_fail = 1;
    __syncthreads();
    BB_TICK(1);

    // per-slot passes: warp w owns slots [w * spw, (w + 1) * spw), a row = 32 consecutive slots
    const u32 spw = cap / BB_WARPS, rows = spw / 32u, wbase = warp * spw;

    // ---- B: owned end vertices of every distinct l-mer ---------------------------------------------------------
    if (!s_fail) {
        auto vertex_sides = [&](ulonglong2 e, bool valid) {
            if (!valid) return;
            const u32 slot = (u32)e.x;
            const u64 c = lt_keys[slot];
            const u32 w = lt_cnt[slot], n = w & 0x3fffffffu;
            const bool own_p = (w >> 30) & 1u, own_s = (w >> 31) & 1u;
            const bool pal = !(l & 1u) && c == bk_revcomp(c, l);
            const u32 m0 = pal ? 2u * n : n;
            bool first;
            // degree slots of the CANONICAL strand v0 of a vertex: lcount[v0][t] (mask bit t) and ecount[v0][t] (bit 4 + t);
            // the other strand mirrors them: lcount[rc v0][t] = ecount[v0][3 - t], ecount[rc v0][t] = lcount[v0][3 - t]
            if (own_p) {   // strand c leaves prefix(c) with m0, last base t
                const u64 p = c >> 2, rp = bk_revcomp(p, k);
                const u32 t = (u32)c & 3u;
                const u32 vs = sm_insert(vt_keys, cap, p < rp ? p : rp, first, plimit);
                if (vs == 0xffffffffu) fail = true;
                else {
                    atomicAdd((p <= rp) ? vt_a + vs : vt_b + vs, m0);   // p is the canonical strand (or a palindrome): its leaving total
                    const u32 bits = p < rp ? (1u << t) : (p > rp ? (16u << (3u - t)) : ((1u << t) | (16u << (3u - t))));
                    atomicOr(vt_m + (vs >> 2), bits << (8u * (vs & 3u)));
                }
            }
            if (own_s && !pal) {   // strand c enters suffix(c) with n, first base t (a palindromic l-mer is covered by its prefix side)
                const u64 s = c & kmask, rs = bk_revcomp(s, k);
                const u32 t = (u32)(c >> (2 * k)) & 3u;
                const u32 vs = sm_insert(vt_keys, cap, s < rs ? s : rs, first, plimit);
                if (vs == 0xffffffffu) fail = true;
                else {
                    if (s == rs) atomicAdd(vt_a + vs, n);            // palindromic vertex: one strand, leaving total == entering total
                    else atomicAdd((s < rs) ? vt_b + vs : vt_a + vs, n);   // canonical strand: entering; else the mirror = leaving of the canonical strand
                    const u32 bits = s < rs ? (16u << t) : (s > rs ? (1u << (3u - t)) : ((16u << t) | (1u << (3u - t))));
                    atomicOr(vt_m + (vs >> 2), bits << (8u * (vs & 3u)));
                }
            }
        };
        for (u32 row = 0; row < rows; row++) {
            const u32 slot = wbase + row * 32u + lane;
            wq.push(lt_keys[slot] != EULER_EMPTY_KEY, make_ulonglong2(slot, 0), lt_mask);
            if (wq.n >= 32u) vertex_sides(wq.pop32(lane), true);
        }
        bool valid;
        const ulonglong2 e = wq.rest(lane, valid);
        vertex_sides(e, valid);
        if (fail) s_fail = 1;
    }
    __syncthreads();
    BB_TICK(2);
    const bool failed = s_fail != 0;

    // ---- C: totals per warp ----------------------------------------------------------------------------------------
    {
        u64 t_rec = 0, t_edges = 0, t_v = 0, t_w = 0;
        if (!failed) {
            for (u32 row = 0; row < rows; row++) {
                const u32 slot = wbase + row * 32u + lane;
                const LtSlot s = bb_lt_slot(lt_keys, lt_cnt, slot, l);
                t_rec += s.recs;
                t_edges += s.edges;
                const u64 v = vt_keys[slot];
                if (v != EULER_EMPTY_KEY) {
                    const bool palv = !(k & 1u) && v == bk_revcomp(v, k);
                    t_v += palv ? 1u : 2u;
                    t_w += palv ? (u64)vt_a[slot] : (u64)vt_a[slot] + vt_b[slot];
                }
            }
        }
        t_rec = warp_sum_u64(t_rec); t_edges = warp_sum_u64(t_edges); t_v = warp_sum_u64(t_v); t_w = warp_sum_u64(t_w);
        if (lane == 0) { s_wtot[warp][0] = t_rec; s_wtot[warp][1] = t_edges; s_wtot[warp][2] = t_v; s_wtot[warp][3] = t_w; }
    }
    __syncthreads();
    u64 off_rec = 0, off_edges = 0, off_v = 0, off_w = 0, tot_rec = 0, tot_e = 0, tot_v = 0, tot_w = 0;
#pragma unroll
    for (int w = 0; w < BB_WARPS; w++) {
        const u64 r0 = s_wtot[w][0], r1 = s_wtot[w][1], r2 = s_wtot[w][2], r3 = s_wtot[w][3];
        if (w < warp) { off_rec += r0; off_edges += r1; off_v += r2; off_w += r3; }
        tot_rec += r0; tot_e += r1; tot_v += r2; tot_w += r3;
    }
    const u64 tot_uv = (tot_v << 32) | tot_rec;
    BB_TICK(3);

    // ---- D: look-back over the buckets, in the order in which they get HERE ------------------------------------------
    // The output ticket is taken now, not at the start: every lower ticket belongs to a block that has already
    // reached this point, so its totals are published (or are a few instructions away) and nobody ever waits for
    // another block's counting.  The artefacts therefore come out in completion order of the buckets.
    if (warp == 0 && failed) {
        // A bucket that does not fit the tables takes no output ticket (the chain never sees it): it is listed and
        // rebuilt by the second pass with the largest tables, its artefacts appended after everybody else's.
        if (lane == 0) {
            bool lost = true;
            if (!a.second) {
                const u32 i = atomicAdd(a.redo, 1u);
                if (i < BKT_REDO_CAP) { a.redo[1 + i] = b; lost = false; }
            }
            if (lost) atomicOr((unsigned long long *)(a.stats + 2), (unsigned long long)BKT_FLAG_TABLE);
            atomicMax((unsigned long long *)(a.stats + 6), (unsigned long long)max_region);
        }
    } else if (warp == 0) {
        u32 ticket = 0;
        if (lane == 0) ticket = atomicAdd(a.ticket + 1, 1u);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        u64 pre_uv = 0, pre_e = 0;
        if (ticket == 0) {
            if (lane == 0) {
                st_vol_u64(a.inc_uv, tot_uv);
                st_vol_u64(a.inc_e, tot_e);
                __threadfence();
                st_vol_u32(a.flag, 2u);
            }
        } else {
            if (lane == 0) {
                st_vol_u64(a.agg_uv + ticket, tot_uv);
                st_vol_u64(a.agg_e + ticket, tot_e);
                __threadfence();
                st_vol_u32(a.flag + ticket, 1u);
            }
            long long look = (long long)ticket - 1;
            while (true) {
                const long long idx = look - lane;
                u32 fl = 2u;
                u64 vuv = 0, ve = 0;
                if (idx >= 0) {
                    u32 spins = 0;
                    do {
                        fl = ld_vol_u32(a.flag + idx);
                        if (fl == 0u) __nanosleep(64);   // leave the issue slots to the blocks that still have work
                        if (fl == 0u && ++spins > (1u << 22)) {   // a predecessor never published: give up instead of hanging the GPU
                            atomicOr((unsigned long long *)(a.stats + 2), (unsigned long long)BKT_FLAG_INTERNAL);
                            fl = 3u;
                        }
                    } while (fl == 0u);
                    __threadfence();
                    if (fl != 3u) {
                        vuv = ld_vol_u64((fl == 2u ? a.inc_uv : a.agg_uv) + idx);
                        ve = ld_vol_u64((fl == 2u ? a.inc_e : a.agg_e) + idx);
                    } else {
                        fl = 2u;   // stop the look-back here
                    }
                }
                const unsigned inc_mask = __ballot_sync(0xffffffffu, fl == 2u);
                if (inc_mask) {
                    const int firsti = __ffs(inc_mask) - 1;
                    if (lane > firsti) { vuv = 0; ve = 0; }
                }
                vuv = warp_sum_u64(vuv);
                ve = warp_sum_u64(ve);
                pre_uv += vuv;
                pre_e += ve;
                if (inc_mask) break;
                look -= 32;
            }
            if (lane == 0) {
                st_vol_u64(a.inc_uv + ticket, pre_uv + tot_uv);
                st_vol_u64(a.inc_e + ticket, pre_e + tot_e);
                __threadfence();
                st_vol_u32(a.flag + ticket, 2u);
            }
        }
        if (lane == 0) {
            s_base[0] = pre_uv;
            s_base[1] = pre_e;
            // grand totals: the inclusive totals grow with the ticket, the last one to arrive leaves the maximum
            atomicMax((unsigned long long *)(a.stats + 3), (unsigned long long)((pre_uv + tot_uv) & 0xffffffffull));
            atomicMax((unsigned long long *)(a.stats + 4), (unsigned long long)((pre_uv + tot_uv) >> 32));
            atomicMax((unsigned long long *)(a.stats + 5), (unsigned long long)(pre_e + tot_e));
            atomicMax((unsigned long long *)(a.stats + 6), (unsigned long long)max_region);
            if (tot_w != tot_e) atomicOr((unsigned long long *)(a.stats + 2), (unsigned long long)BKT_FLAG_INTERNAL);
        }
    }
    __syncthreads();
    BB_TICK(4);
    if (failed) return;
    const u64 ubase = s_base[0] & 0xffffffffull, vbase = s_base[0] >> 32, ebase = s_base[1];
    if (ubase + tot_rec > a.ucap || vbase + tot_v > a.vcap) {
        if (tid == 0) atomicOr((unsigned long long *)(a.stats + 2), (unsigned long long)BKT_FLAG_OUTPUT);
        return;
    }

    // ---- E1: vertex artefacts -----------------------------------------------------------------------------------------
    {
        bool bad = false;
        // entry: x = slot | vid << 32, y = wpos (the edge-offset prefix of the vertex)
        auto vertex_out = [&](ulonglong2 e, bool valid) {
            if (!valid) return;
            const u32 slot = (u32)e.x & 0xffffu;
            const u64 vid = e.x >> 32;
            const u64 v = vt_keys[slot], rv = bk_revcomp(v, k);
            const bool palv = v == rv;
            const u32 L0 = vt_a[slot], E0 = vt_b[slot];
            u32 lc[4], ec[4];
            const u32 mask = (vt_m[slot >> 2] >> (8u * (slot & 3u))) & 0xffu;   // only the non-zero degree slots are looked up
#pragma unroll
            for (u32 t = 0; t < 4; t++) {   // rc(v t) = comp(t) rc(v), rc(t v) = rc(v) comp(t): no bit reversal per neighbour
                lc[t] = (mask >> t) & 1u ? bb_bs_count(lt_keys, lt_cnt, cap, (v << 2) | t, ((u64)(3u - t) << (2 * k)) | rv) : 0u;
                ec[t] = (mask >> (4u + t)) & 1u ? bb_bs_count(lt_keys, lt_cnt, cap, ((u64)t << (2 * k)) | v, (rv << 2) | (3u - t)) : 0u;
            }
            const u32 ls = lc[0] + lc[1] + lc[2] + lc[3], es = ec[0] + ec[1] + ec[2] + ec[3];
            if (ls != L0 || es != (palv ? L0 : E0)) bad = true;
            const u32 P = (u32)e.y;
            vt_a[slot] = (u32)vid;   // the id of the canonical strand, for the edge pass
            a.vkeys[vid] = v;
            reinterpret_cast<uint4 *>(a.lcount)[vid] = make_uint4(lc[0], lc[1], lc[2], lc[3]);
            reinterpret_cast<uint4 *>(a.ecount)[vid] = make_uint4(ec[0], ec[1], ec[2], ec[3]);
            reinterpret_cast<uint4 *>(a.lstart)[vid] = make_uint4(P, P + lc[0], P + lc[0] + lc[1], P + lc[0] + lc[1] + lc[2]);
            reinterpret_cast<uint4 *>(a.estart)[vid] = make_uint4(P, P + ec[0], P + ec[0] + ec[1], P + ec[0] + ec[1] + ec[2]);
            euler_vertex x;
            x.vid = v; x.ep = P; x.ecount = es; x.lp = P; x.lcount = ls;
            a.ev[vid] = x;
            if (!palv) {   // the reverse strand: lcount[rc v][b] = ecount[v][3-b], ecount[rc v][a] = lcount[v][3-a]
                const u64 id1 = vid + 1;
                const u32 PL = P + ls, PE = P + es;
                a.vkeys[id1] = rv;
                reinterpret_cast<uint4 *>(a.lcount)[id1] = make_uint4(ec[3], ec[2], ec[1], ec[0]);
                reinterpret_cast<uint4 *>(a.ecount)[id1] = make_uint4(lc[3], lc[2], lc[1], lc[0]);
                reinterpret_cast<uint4 *>(a.lstart)[id1] = make_uint4(PL, PL + ec[3], PL + ec[3] + ec[2], PL + ec[3] + ec[2] + ec[1]);
                reinterpret_cast<uint4 *>(a.estart)[id1] = make_uint4(PE, PE + lc[3], PE + lc[3] + lc[2], PE + lc[3] + lc[2] + lc[1]);
                x.vid = rv; x.ep = PE; x.ecount = ls; x.lp = PL; x.lcount = es;
                a.ev[id1] = x;
            }
        };
        u64 vcarry = vbase + off_v, wcarry = ebase + off_w;
        for (u32 row = 0; row < rows; row++) {
            const u32 slot = wbase + row * 32u + lane;
            const u64 v = vt_keys[slot];
            const bool occ = v != EULER_EMPTY_KEY;
            const bool palv = occ && !(k & 1u) && v == bk_revcomp(v, k);
            const u32 L0 = occ ? vt_a[slot] : 0u, E0 = occ ? vt_b[slot] : 0u;
            const u32 nstr = occ ? (palv ? 1u : 2u) : 0u;
            const u64 wsum = occ ? (palv ? (u64)L0 : (u64)L0 + E0) : 0ull;
            const u32 vinc = warp_incl_u32(nstr, lane);
            const u64 winc = warp_incl_u64(wsum, lane);
            const u64 vid = vcarry + vinc - nstr, wpos = wcarry + winc - wsum;
            vcarry += __shfl_sync(0xffffffffu, vinc, 31);
            wcarry += __shfl_sync(0xffffffffu, winc, 31);
            wq.push(occ, make_ulonglong2((u64)slot | (vid << 32), wpos), lt_mask);
            if (wq.n >= 32u) vertex_out(wq.pop32(lane), true);
        }
        bool valid;
        const ulonglong2 e = wq.rest(lane, valid);
        vertex_out(e, valid);
        if (bad) atomicOr((unsigned long long *)(a.stats + 2), (unsigned long long)BKT_FLAG_INTERNAL);
    }
    __syncthreads();
    BB_TICK(5);

    // ---- E2: edge artefacts -------------------------------------------------------------------------------------------
    {
        bool bfail = false;   // a vertex that phase B inserted is missing: cannot happen
        // entry: x = slot | first record index << 32, y = first edge offset
        auto edge_out = [&](ulonglong2 e, bool valid) {
            if (!valid) return;
            const u32 slot = (u32)e.x & 0xffffu;
            u64 rec = e.x >> 32, eo = e.y;
            const LtSlot s = bb_lt_slot(lt_keys, lt_cnt, slot, l);
            const u64 c = s.c, r = bk_revcomp(c, l);
            const u32 n = s.n, m0 = s.pal ? 2u * n : n;
            const u64 p = c >> 2, sf = c & kmask, rp = r & kmask, rs = r >> 2;   // rc(prefix c) = suffix(rc c), rc(suffix c) = prefix(rc c)
            u32 id_p = EULER_NO_ID, id_rp = EULER_NO_ID, id_s = EULER_NO_ID, id_rs = EULER_NO_ID;
            // No exit between the look-ups and the stores: with a `return` on the cannot-happen miss the lanes left the
            // probe loops one group at a time and ran the ten stores below 4.5 lanes wide (ncu source page, round 2);
            // this way the warp reconverges after each loop.  A miss is reported and writes ids of vertex 0.
            if (s.own_p) {
                const u32 vs = sm_find(vt_keys, cap, p < rp ? p : rp);
                if (vs == 0xffffffffu) bfail = true;   // cannot happen: inserted in B
                const u32 i0 = vs == 0xffffffffu ? 0u : vt_a[vs];
                id_p = p <= rp ? i0 : i0 + 1u;
                id_rp = rp <= p ? i0 : i0 + 1u;
            }
            if (s.own_s) {
                const u32 vs = sm_find(vt_keys, cap, sf < rs ? sf : rs);
                if (vs == 0xffffffffu) bfail = true;
                const u32 i0 = vs == 0xffffffffu ? 0u : vt_a[vs];
                id_s = sf <= rs ? i0 : i0 + 1u;
                id_rs = rs <= sf ? i0 : i0 + 1u;
            }
            if (s.own_p) {   // strand c is homed with its prefix vertex
                a.lkeys[rec] = c; a.lvals[rec] = m0; a.loffs[rec] = (u32)eo; a.ev1[rec] = id_p; a.ev2[rec] = id_s;
                rec++;
                eo += m0;
            }
            if (s.own_s && !s.pal) {   // strand rc(c) runs from rc(suffix c) to rc(prefix c)
                a.lkeys[rec] = r; a.lvals[rec] = n; a.loffs[rec] = (u32)eo; a.ev1[rec] = id_rs; a.ev2[rec] = id_rp;
            }
        };
        u64 rcarry = ubase + off_rec, ecarry = ebase + off_edges;
        for (u32 row = 0; row < rows; row++) {
            const u32 slot = wbase + row * 32u + lane;
            const LtSlot s = bb_lt_slot(lt_keys, lt_cnt, slot, l);
            const u32 rinc = warp_incl_u32(s.recs, lane);
            const u64 einc = warp_incl_u64(s.edges, lane);
            const u64 rec = rcarry + rinc - s.recs, eo = ecarry + einc - s.edges;
            rcarry += __shfl_sync(0xffffffffu, rinc, 31);
            ecarry += __shfl_sync(0xffffffffu, einc, 31);
            wq.push(s.recs != 0, make_ulonglong2((u64)slot | (rec << 32), eo), lt_mask);
            if (wq.n >= 32u) edge_out(wq.pop32(lane), true);
        }
        bool valid;
        const ulonglong2 e = wq.rest(lane, valid);
        edge_out(e, valid);
        if (bfail) atomicOr((unsigned long long *)(a.stats + 2), (unsigned long long)BKT_FLAG_INTERNAL);
    }
    __syncthreads();
    BB_TICK(6);
}

// ---- edges that cross buckets ------------------------------------------------------------------------------------------
// After the build, the record of strand x has ev2 == NO_ID when suffix(x) belongs to another bucket.  That bucket
// holds the record of strand rc(x), whose ev1 is id(rc(suffix x)) -- the partner strand of the vertex we are
// looking for -- and it is looking for the partner of OUR ev1.  Two passes over the records through a small global
// table keyed by the canonical l-mer (value slot 0: id(rc prefix c), written from the record of strand c; slot 1:
// id(suffix c), written from the record of strand rc(c)); no global atomic or dependent global access inside the
// per-bucket kernel.  An l-mer whose other end lives on another RANK finds nothing and keeps NO_ID.
__device__ __forceinline__ u64 bnd_hash(u64 c, u64 bmask) { return (((c ^ (c >> 29)) * 0x9E3779B97F4A7C15ull) >> 20) & bmask; }
__global__ void __launch_bounds__(256) bkt_boundary_publish_kernel(const u64 *__restrict__ lkeys, const u32 *__restrict__ ev1,
                                                                   const u32 *__restrict__ ev2, const u64 *__restrict__ d_u, u64 ucap, u32 l,
                                                                   u64 *__restrict__ bkeys, u32 *__restrict__ bvals, u64 bcap, u64 *__restrict__ stats)
{
    const u64 n = *d_u < ucap ? *d_u : ucap;
    const u64 bmask = bcap - 1;
    const u32 k = l - 1;
    bool bfail = false;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        if (ev2[i] != EULER_NO_ID) continue;
        const u64 x = lkeys[i], r = bk_revcomp(x, l);
        const u64 c = x < r ? x : r;
        const u64 p = x >> 2, rp = bk_revcomp(p, k);
        const u32 id = ev1[i];
        const u32 partner = p < rp ? id + 1u : (p > rp ? id - 1u : id);   // id(rc(prefix x)): the two strands of a vertex have adjacent ids
        u64 h = bnd_hash(c, bmask);
        u32 probe = 0;
        for (; probe < 4096; probe++) {
            const u64 cur = bkeys[h];
            if (cur == c) break;
            if (cur == EULER_EMPTY_KEY) {
                const u64 old = atomicCAS((unsigned long long *)(bkeys + h), EULER_EMPTY_KEY, c);
                if (old == EULER_EMPTY_KEY || old == c) break;
            }
            h = (h + 1) & bmask;
        }
        if (probe == 4096) bfail = true;
        else bvals[2 * h + (x == c ? 0 : 1)] = partner;
    }
    if (bfail) atomicOr((unsigned long long *)(stats + 2), (unsigned long long)BKT_FLAG_BOUNDARY);
}
__global__ void __launch_bounds__(256) bkt_fixup_kernel(const u64 *__restrict__ lkeys, u32 *__restrict__ ev2, const u64 *__restrict__ d_u,
                                                        u64 ucap, u32 l, const u64 *__restrict__ bkeys, const u32 *__restrict__ bvals,
                                                        u64 bcap)
{
    const u64 n = *d_u < ucap ? *d_u : ucap;
    const u64 bmask = bcap - 1;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        if (ev2[i] != EULER_NO_ID) continue;
        const u64 x = lkeys[i], r = bk_revcomp(x, l);
        const u64 c = x < r ? x : r;
        u64 h = bnd_hash(c, bmask);
        for (u32 probe = 0; probe < 4096; probe++) {
            const u64 cur = bkeys[h];
            if (cur == c) {
                ev2[i] = bvals[2 * h + (x == c ? 1 : 0)];   // strand c wants id(suffix c); strand rc(c) wants id(rc prefix c)
                break;
            }
            if (cur == EULER_EMPTY_KEY) break;   // the other end lives on another rank
            h = (h + 1) & bmask;
        }
    }
}

#ifndef EULER_SIMT_EMU   // host side (launches, canonical-id helpers): not part of the CPU emulation
size_t bkt_build_smem(u32 cap) { return (size_t)29 * cap; }   // two key arrays (8 B) + count + two vertex words (4 B) + mask byte

int bkt_build(euler_ctx *ctx, const BktBuild &B)
{
    if (!B.nb) return EULER_OK;
    if (B.cap < BB_THREADS || B.cap % 256) return euler_fail(ctx, EULER_ERR_ARG, "bucket table capacity must be a multiple of 256");
    if (B.bcap & (B.bcap - 1)) return euler_fail(ctx, EULER_ERR_ARG, "boundary table capacity must be a power of two");
    if (B.cap > BKT_MAX_CAP) return euler_fail(ctx, EULER_ERR_ARG, "bucket table capacity above %u", BKT_MAX_CAP);
    const size_t smem = bkt_build_smem(B.cap), smem_max = bkt_build_smem(BKT_MAX_CAP);
    static bool smem_set[64] = {false};   // the attribute belongs to the device
    if (ctx->device < 0 || ctx->device >= 64 || !smem_set[ctx->device]) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(bkt_build_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        CUDA_TRY(ctx, cudaFuncSetAttribute(bkt_build_kernel<22>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        CUDA_TRY(ctx, cudaFuncSetAttribute(bkt_build_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        if (ctx->device >= 0 && ctx->device < 64) smem_set[ctx->device] = true;
    }
    // state: flag u32[nb] | work ticket, output ticket u32[2] | redo count, redo list u32[1 + BKT_REDO_CAP] |
    //        (16-byte aligned) agg_uv, agg_e, inc_uv, inc_e u64[nb]
    u32 *flag = (u32 *)B.state;
    u32 *ticket = flag + B.nb;
    const size_t zero_bytes = ((size_t)B.nb + 2 + 1 + BKT_REDO_CAP) * 4;
    u64 *w64 = (u64 *)((char *)B.state + (zero_bytes + 15) / 16 * 16);
    CUDA_TRY(ctx, cudaMemsetAsync(B.state, 0, zero_bytes, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(B.bkeys, 0xFF, B.bcap * 8, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(B.bvals, 0xFF, B.bcap * 8, ctx->stream));
    BkBuildArgs a;
    a.records = (const uint4 *)B.records; a.counts = B.counts; a.nb = B.nb; a.nranks = B.nranks; a.rcap = B.rcap; a.l = B.l;
    a.cap = B.cap;
    a.lkeys = B.lkeys; a.lvals = B.lvals; a.loffs = B.loffs; a.ev1 = B.ev1; a.ev2 = B.ev2; a.ucap = B.ucap;
    a.vkeys = B.vkeys; a.lcount = B.lcount; a.ecount = B.ecount; a.lstart = B.lstart; a.estart = B.estart; a.ev = B.ev; a.vcap = B.vcap;
    a.flag = flag; a.ticket = ticket; a.agg_uv = w64; a.agg_e = w64 + B.nb; a.inc_uv = w64 + 2ull * B.nb; a.inc_e = w64 + 3ull * B.nb;
    a.bkeys = B.bkeys; a.bvals = B.bvals; a.bcap = B.bcap; a.stats = B.stats;
    a.redo = ticket + 2; a.second = 0;
    if (B.l == 32) bkt_build_kernel<32><<<B.nb, BB_THREADS, smem, ctx->stream>>>(a);
    else if (B.l == 22) bkt_build_kernel<22><<<B.nb, BB_THREADS, smem, ctx->stream>>>(a);
    else bkt_build_kernel<0><<<B.nb, BB_THREADS, smem, ctx->stream>>>(a);
    CUDA_TRY(ctx, cudaGetLastError());
    {
        // second pass: the few buckets that overflowed the tables above (skewed minimizers) are rebuilt with the largest
        // tables one block can hold, one block per SM; with nothing listed every block leaves at once.  (Always launched:
        // a first pass that already runs with the largest tables lists its failures all the same.)
        a.second = 1; a.cap = BKT_MAX_CAP;
        if (B.l == 32) bkt_build_kernel<32><<<BKT_REDO_CAP, BB_THREADS, smem_max, ctx->stream>>>(a);
        else if (B.l == 22) bkt_build_kernel<22><<<BKT_REDO_CAP, BB_THREADS, smem_max, ctx->stream>>>(a);
        else bkt_build_kernel<0><<<BKT_REDO_CAP, BB_THREADS, smem_max, ctx->stream>>>(a);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    const unsigned g = (unsigned)ctx->num_sms * 8;
    bkt_boundary_publish_kernel<<<g, 256, 0, ctx->stream>>>(B.lkeys, B.ev1, B.ev2, B.stats + 3, B.ucap, B.l, B.bkeys, B.bvals, B.bcap, B.stats);
    bkt_fixup_kernel<<<g, 256, 0, ctx->stream>>>(B.lkeys, B.ev2, B.stats + 3, B.ucap, B.l, B.bkeys, B.bvals, B.bcap);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

size_t bkt_state_bytes(u32 nb) { return (((size_t)nb + 2 + 1 + BKT_REDO_CAP) * 4 + 15) / 16 * 16 + (size_t)nb * 32 + 16; }

// ---- canonical ids (EULER_RUN_CANONICAL_IDS): bucket order -> ascending key order -------------------------------------
// The bucketed build numbers vertices and edge records in bucket order.  Ids = rank in ascending key order
//  are a permutation of that: sort the keys with their old index as payload, then gather.
__global__ void __launch_bounds__(256) bkt_iota_kernel(u32 *__restrict__ v, u64 n)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (u32)i;
}
__global__ void __launch_bounds__(256) bkt_invert_kernel(const u32 *__restrict__ perm, u64 n, u32 *__restrict__ inv)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[perm[i]] = (u32)i;
}
__global__ void __launch_bounds__(256) bkt_gather_rows_kernel(const u32 *__restrict__ perm, u64 n, const uint4 *__restrict__ a_in,
                                                              const uint4 *__restrict__ b_in, uint4 *__restrict__ a_out,
                                                              uint4 *__restrict__ b_out)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 o = perm[i];
    a_out[i] = a_in[o];
    b_out[i] = b_in[o];
}
__global__ void __launch_bounds__(256) bkt_gather_edges_kernel(const u32 *__restrict__ perm, u64 n, const u32 *__restrict__ newid,
                                                               const u32 *__restrict__ lvals, const u32 *__restrict__ ev1,
                                                               const u32 *__restrict__ ev2, u32 *__restrict__ lvals_out,
                                                               u32 *__restrict__ ev1_out, u32 *__restrict__ ev2_out)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 o = perm[i];
    lvals_out[i] = lvals[o];
    const u32 a = ev1[o], b = ev2[o];
    ev1_out[i] = a == EULER_NO_ID ? a : newid[a];
    ev2_out[i] = b == EULER_NO_ID ? b : newid[b];
}

int bkt_iota(euler_ctx *ctx, u32 *v, u64 n)
{
    if (!n) return EULER_OK;
    bkt_iota_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(v, n);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}
int bkt_invert_perm(euler_ctx *ctx, const u32 *perm, u64 n, u32 *inv)
{
    if (!n) return EULER_OK;
    bkt_invert_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(perm, n, inv);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}
int bkt_gather_rows(euler_ctx *ctx, const u32 *perm, u64 n, const u32 *a_in, const u32 *b_in, u32 *a_out, u32 *b_out)
{
    if (!n) return EULER_OK;
    bkt_gather_rows_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(perm, n, (const uint4 *)a_in, (const uint4 *)b_in, (uint4 *)a_out,
                                                                     (uint4 *)b_out);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}
int bkt_gather_edges(euler_ctx *ctx, const u32 *perm, u64 n, const u32 *newid, const u32 *lvals, const u32 *ev1, const u32 *ev2,
                     u32 *lvals_out, u32 *ev1_out, u32 *ev2_out)
{
    if (!n) return EULER_OK;
    bkt_gather_edges_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(perm, n, newid, lvals, ev1, ev2, lvals_out, ev1_out, ev2_out);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}
#endif   // EULER_SIMT_EMU
