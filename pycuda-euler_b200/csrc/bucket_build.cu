// bucket_build.cu -- pass 2 of the bucketed hot path (bucket.cuh): one CTA per bucket builds the bucket's part
// of the de Bruijn graph entirely in shared memory and writes every artefact sequentially.
//
// Replaces, per bucket, the host dictionary fill of eulercuda.readLmersKmersCuda (eulercuda.py:141-178), the
// hash table of pygpuhash.create_hash_table_device (pygpuhash.py:262-315) and D1-D6 of
// pydebruijn.construct_debruijn_graph_device (pydebruijn.py:516-620):
//   A  count the canonical l-mers spelled by the bucket's records.  Every warp streams its own slice of the
//      records (two batches of 32 in registers, handed to the lanes by shuffle), every lane rolls the forward and
//      reverse-complement l-mer of its record one base per step, and a lane that finishes takes the warp's next
//      record.  Shared-memory table of 4-slot buckets (two 128-bit loads per probe, atomicCAS claim); the count
//      word also carries the "prefix / suffix vertex is ours" bits.
//   B  insert the owned end vertices of every distinct l-mer into a second shared-memory table and add the
//      multiplicity to the vertex's leaving / entering total;
//   C  totals of (edge records, vertices, edge multiplicities) over the slots;
//   D  decoupled look-back over the buckets in ticket order (buckets are taken largest first, so a bucket's
//      predecessors are done when it gets here): global bases;
//   E  vertex artefacts (the eight degree slots of a vertex are eight lookups in the bucket's own l-mer table)
//      and edge artefacts, written in slot order with warp-row scans: consecutive lanes write consecutive ids.
//      The suffix vertex of an edge whose suffix lives in another bucket is published / resolved through a
//      small global table keyed by the canonical l-mer (bkt_fixup_kernel).
#include "bucket.cuh"
#include "kernels.h"

#define BB_THREADS 256
#define BB_WARPS (BB_THREADS / 32)
#define BB_STEPS 2   // l-mers a lane rolls between two refills
#ifndef BB_MINB
#define BB_MINB 4
#endif
#define BB_BINS 1024

struct BkBuildArgs {
    const uint4 *records;
    const u32 *counts;   // [nb * nranks] records in region (bucket, source rank)
    const u32 *order;    // ticket -> bucket (largest first)
    u32 nb, nranks, rcap, l;
    u32 cap;             // slots of each shared-memory table (a multiple of 256)
    u64 *lkeys; u32 *lvals, *loffs, *ev1, *ev2; u64 ucap;
    u64 *vkeys; u32 *lcount, *ecount, *lstart, *estart; euler_vertex *ev; u64 vcap;
    u32 *flag; u64 *agg_uv, *agg_e, *inc_uv, *inc_e; u32 *ticket;
    u64 *bkeys; u32 *bvals; u64 bcap;
    u64 *stats;
};

// home bucket (4 slots) of a key in a shared-memory table of nbk buckets
__device__ __forceinline__ u32 bb_home(u64 key, u32 nbk)
{
    u32 h = (u32)key * 0x9E3779B1u + (u32)(key >> 32) * 0x85EBCA77u;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    return __umulhi(h, nbk);
}
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

// Slots of a bucket fill in order and are never freed, so the first EMPTY slot ends a search.
// insert: returns the slot, 0xffffffff when the table is full; first = this call claimed the slot.
__device__ __forceinline__ u32 sm_insert(u64 *keys, u32 nbk, u64 key, bool &first)
{
    u32 hb = bb_home(key, nbk);
    first = false;
    for (u32 probe = 0; probe < nbk; probe++) {
        u64 *bk = keys + 4u * hb;
        const ulonglong2 q0 = *reinterpret_cast<const ulonglong2 *>(bk), q1 = *reinterpret_cast<const ulonglong2 *>(bk + 2);
        if (q0.x == key) return 4u * hb;
        if (q0.y == key) return 4u * hb + 1u;
        if (q1.x == key) return 4u * hb + 2u;
        if (q1.y == key) return 4u * hb + 3u;
        u32 fe = q0.x == EULER_EMPTY_KEY ? 0u : (q0.y == EULER_EMPTY_KEY ? 1u : (q1.x == EULER_EMPTY_KEY ? 2u : (q1.y == EULER_EMPTY_KEY ? 3u : 4u)));
        for (; fe < 4u; fe++) {   // claim the first empty slot; a slot lost to another key sends us to the next one
            const u64 old = atomicCAS(bk + fe, EULER_EMPTY_KEY, key);
            if (old == EULER_EMPTY_KEY) { first = true; return 4u * hb + fe; }
            if (old == key) return 4u * hb + fe;
        }
        hb = hb + 1u == nbk ? 0u : hb + 1u;
    }
    return 0xffffffffu;
}
__device__ __forceinline__ u32 sm_find(const u64 *keys, u32 nbk, u64 key)
{
    u32 hb = bb_home(key, nbk);
    for (u32 probe = 0; probe < nbk; probe++) {
        const u64 *bk = keys + 4u * hb;
        const ulonglong2 q0 = *reinterpret_cast<const ulonglong2 *>(bk), q1 = *reinterpret_cast<const ulonglong2 *>(bk + 2);
        if (q0.x == key) return 4u * hb;
        if (q0.y == key) return 4u * hb + 1u;
        if (q1.x == key) return 4u * hb + 2u;
        if (q1.y == key) return 4u * hb + 3u;
        if (q0.x == EULER_EMPTY_KEY || q0.y == EULER_EMPTY_KEY || q1.x == EULER_EMPTY_KEY || q1.y == EULER_EMPTY_KEY) return 0xffffffffu;
        hb = hb + 1u == nbk ? 0u : hb + 1u;
    }
    return 0xffffffffu;
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

// both-strand multiplicity of the l-mer x in the bucket's table (0 when absent)
__device__ __forceinline__ u32 bb_bs_count(const u64 *lt_keys, const u32 *lt_cnt, u32 nbk, u64 x, u32 l)
{
    const u64 r = bk_revcomp(x, l);
    const u32 slot = sm_find(lt_keys, nbk, x < r ? x : r);
    if (slot == 0xffffffffu) return 0u;
    const u32 n = lt_cnt[slot] & 0x3fffffffu;
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
    s.pal = s.c == bk_revcomp(s.c, l);
    if (s.pal) { s.recs = s.own_p ? 1u : 0u; s.edges = s.own_p ? 2ull * s.n : 0ull; }
    else { s.recs = (s.own_p ? 1u : 0u) + (s.own_s ? 1u : 0u); s.edges = (u64)s.n * s.recs; }
    return s;
}

__global__ void __launch_bounds__(BB_THREADS, BB_MINB) bkt_build_kernel(const BkBuildArgs a)
{
    extern __shared__ __align__(16) unsigned char bb_smem[];
    const u32 cap = a.cap, nbk = cap / 4;
    u64 *lt_keys = (u64 *)bb_smem;
    u64 *vt_keys = lt_keys + cap;
    u32 *lt_cnt = (u32 *)(vt_keys + cap);
    u32 *vt_a = lt_cnt + cap;   // leaving total of the canonical strand, later the vertex id
    u32 *vt_b = vt_a + cap;     // entering total of the canonical strand
    __shared__ u64 s_wtot[BB_WARPS][4];
    __shared__ u64 s_base[2];
    __shared__ u32 s_bucket, s_fail;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const u32 l = a.l, k = l - 1;
    const u64 lmask = l >= 32 ? ~0ull : ((1ull << (2 * l)) - 1ull), kmask = lmask >> 2;
    const u32 top = 2 * (l - 1);

    if (tid == 0) { s_bucket = atomicAdd(a.ticket, 1u); s_fail = 0; }
    {
        const ulonglong2 e2 = make_ulonglong2(EULER_EMPTY_KEY, EULER_EMPTY_KEY);
        for (u32 i = tid; i < cap; i += BB_THREADS) {   // lt_keys and vt_keys are adjacent: 2 * cap keys = cap pairs
            reinterpret_cast<ulonglong2 *>(lt_keys)[i] = e2;
        }
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (u32 i = tid; i < 3 * cap / 4; i += BB_THREADS) reinterpret_cast<uint4 *>(lt_cnt)[i] = z;
    }
    __syncthreads();
    const u32 ticket = s_bucket;
    if (ticket >= a.nb) return;   // never: the grid is nb blocks
    const u32 b = a.order ? a.order[ticket] : ticket;

    // ---- A: count the l-mers of the bucket's records ---------------------------------------------------------
    bool fail = false;
    u32 max_region = 0;
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
                if (left) {
                    const u32 cc = (u32)(rem >> 62);
                    rem <<= 2;
                    f = ((f << 2) | cc) & lmask;
                    rc = (rc >> 2) | ((u64)(3u - cc) << top);
                    const bool flip = rc < f;
                    const u64 c = flip ? rc : f;
                    // ownership of the end vertices, in the orientation the read spells
                    const u32 own_pf = (firstl && (hdr & BK_HDR_LFF)) ? 0u : 1u;
                    const u32 own_sf = (left == 1u && (hdr & BK_HDR_RFF)) ? 0u : 1u;
                    const u32 own = flip ? (own_sf | (own_pf << 1)) : (own_pf | (own_sf << 1));   // bit 0: prefix(c), bit 1: suffix(c)
                    bool first;
                    const u32 slot = sm_insert(lt_keys, nbk, c, first);
                    if (slot == 0xffffffffu) fail = true;
                    else atomicAdd(lt_cnt + slot, first ? (1u | (own << 30)) : 1u);
                    firstl = false;
                    left--;
                }
            }
        }
    }
    if (fail) s_fail = 1;
    __syncthreads();

    // per-slot passes: warp w owns slots [w * spw, (w + 1) * spw), a row = 32 consecutive slots
    const u32 spw = cap / BB_WARPS, rows = spw / 32u, wbase = warp * spw;

    // ---- B: owned end vertices of every distinct l-mer ---------------------------------------------------------
    if (!s_fail) {
        for (u32 row = 0; row < rows; row++) {
            const u32 slot = wbase + row * 32u + lane;
            const u64 c = lt_keys[slot];
            if (c == EULER_EMPTY_KEY) continue;
            const u32 w = lt_cnt[slot], n = w & 0x3fffffffu;
            const bool own_p = (w >> 30) & 1u, own_s = (w >> 31) & 1u;
            const bool pal = c == bk_revcomp(c, l);
            const u32 m0 = pal ? 2u * n : n;
            bool first;
            if (own_p) {   // strand c leaves prefix(c) with m0
                const u64 p = c >> 2, rp = bk_revcomp(p, k);
                const u32 vs = sm_insert(vt_keys, nbk, p < rp ? p : rp, first);
                if (vs == 0xffffffffu) fail = true;
                else atomicAdd((p <= rp) ? vt_a + vs : vt_b + vs, m0);   // p is the canonical strand (or a palindrome): its leaving total
            }
            if (own_s && !pal) {   // strand c enters suffix(c) with n (a palindromic l-mer is covered by its prefix side)
                const u64 s = c & kmask, rs = bk_revcomp(s, k);
                const u32 vs = sm_insert(vt_keys, nbk, s < rs ? s : rs, first);
                if (vs == 0xffffffffu) fail = true;
                else if (s == rs) atomicAdd(vt_a + vs, n);            // palindromic vertex: one strand, leaving total == entering total
                else atomicAdd((s < rs) ? vt_b + vs : vt_a + vs, n);   // canonical strand: entering; else the mirror = leaving of the canonical strand
            }
        }
        if (fail) s_fail = 1;
    }
    __syncthreads();
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
                    const bool palv = v == bk_revcomp(v, k);
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

    // ---- D: look-back over the buckets (ticket order) --------------------------------------------------------------
    if (warp == 0) {
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
                        if (fl == 0u && ++spins > (1u << 26)) {   // a predecessor never published: give up instead of hanging the GPU
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
            if (ticket == a.nb - 1) {   // grand totals
                a.stats[3] = (pre_uv + tot_uv) & 0xffffffffull;
                a.stats[4] = (pre_uv + tot_uv) >> 32;
                a.stats[5] = pre_e + tot_e;
            }
            atomicMax((unsigned long long *)(a.stats + 6), (unsigned long long)max_region);
            if (failed) atomicOr((unsigned long long *)(a.stats + 2), (unsigned long long)BKT_FLAG_TABLE);
            if (tot_w != tot_e) atomicOr((unsigned long long *)(a.stats + 2), (unsigned long long)BKT_FLAG_INTERNAL);
        }
    }
    __syncthreads();
    if (failed) return;
    const u64 ubase = s_base[0] & 0xffffffffull, vbase = s_base[0] >> 32, ebase = s_base[1];
    if (ubase + tot_rec > a.ucap || vbase + tot_v > a.vcap) {
        if (tid == 0) atomicOr((unsigned long long *)(a.stats + 2), (unsigned long long)BKT_FLAG_OUTPUT);
        return;
    }

    // ---- E1: vertex artefacts -----------------------------------------------------------------------------------------
    {
        u64 vcarry = vbase + off_v, wcarry = ebase + off_w;
        bool bad = false;
        for (u32 row = 0; row < rows; row++) {
            const u32 slot = wbase + row * 32u + lane;
            const u64 v = vt_keys[slot];
            const bool occ = v != EULER_EMPTY_KEY;
            const u64 rv = occ ? bk_revcomp(v, k) : 0ull;
            const bool palv = v == rv;
            const u32 L0 = occ ? vt_a[slot] : 0u, E0 = occ ? vt_b[slot] : 0u;
            const u32 nstr = occ ? (palv ? 1u : 2u) : 0u;
            const u64 wsum = occ ? (palv ? (u64)L0 : (u64)L0 + E0) : 0ull;
            const u32 vinc = warp_incl_u32(nstr, lane);
            const u64 winc = warp_incl_u64(wsum, lane);
            const u64 vid = vcarry + vinc - nstr, wpos = wcarry + winc - wsum;
            vcarry += __shfl_sync(0xffffffffu, vinc, 31);
            wcarry += __shfl_sync(0xffffffffu, winc, 31);
            if (!occ) continue;
            u32 lc[4], ec[4];
#pragma unroll
            for (u32 t = 0; t < 4; t++) {
                lc[t] = bb_bs_count(lt_keys, lt_cnt, nbk, (v << 2) | t, l);
                ec[t] = bb_bs_count(lt_keys, lt_cnt, nbk, ((u64)t << (2 * k)) | v, l);
            }
            const u32 ls = lc[0] + lc[1] + lc[2] + lc[3], es = ec[0] + ec[1] + ec[2] + ec[3];
            if (ls != L0 || es != (palv ? L0 : E0)) bad = true;
            const u32 P = (u32)wpos;
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
        }
        if (bad) atomicOr((unsigned long long *)(a.stats + 2), (unsigned long long)BKT_FLAG_INTERNAL);
    }
    __syncthreads();

    // ---- E2: edge artefacts -------------------------------------------------------------------------------------------
    {
        u64 rcarry = ubase + off_rec, ecarry = ebase + off_edges;
        bool bfail = false;
        for (u32 row = 0; row < rows; row++) {
            const u32 slot = wbase + row * 32u + lane;
            const LtSlot s = bb_lt_slot(lt_keys, lt_cnt, slot, l);
            const u32 rinc = warp_incl_u32(s.recs, lane);
            const u64 einc = warp_incl_u64(s.edges, lane);
            u64 rec = rcarry + rinc - s.recs, eo = ecarry + einc - s.edges;
            rcarry += __shfl_sync(0xffffffffu, rinc, 31);
            ecarry += __shfl_sync(0xffffffffu, einc, 31);
            if (!s.recs) continue;
            const u64 c = s.c, r = bk_revcomp(c, l);
            const u32 n = s.n, m0 = s.pal ? 2u * n : n;
            const u64 p = c >> 2, sf = c & kmask, rp = bk_revcomp(p, k), rs = bk_revcomp(sf, k);
            u32 id_p = EULER_NO_ID, id_rp = EULER_NO_ID, id_s = EULER_NO_ID, id_rs = EULER_NO_ID;
            if (s.own_p) {
                const u32 vs = sm_find(vt_keys, nbk, p < rp ? p : rp);
                if (vs == 0xffffffffu) { bfail = true; continue; }   // cannot happen: inserted in B
                const u32 i0 = vt_a[vs];
                id_p = p <= rp ? i0 : i0 + 1u;
                id_rp = rp <= p ? i0 : i0 + 1u;
            }
            if (s.own_s) {
                const u32 vs = sm_find(vt_keys, nbk, sf < rs ? sf : rs);
                if (vs == 0xffffffffu) { bfail = true; continue; }
                const u32 i0 = vt_a[vs];
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
            if (s.own_p != s.own_s) {   // the other end vertex lives in another bucket: publish our side's id under the canonical l-mer
                const u64 bmask = a.bcap - 1;
                u64 h = ((c ^ (c >> 29)) * 0x9E3779B97F4A7C15ull >> 20) & bmask;
                u32 probe = 0;
                for (; probe < 4096; probe++) {
                    const u64 cur = ld_vol_u64(a.bkeys + h);
                    if (cur == c) break;
                    if (cur == EULER_EMPTY_KEY) {
                        const u64 old = atomicCAS((unsigned long long *)(a.bkeys + h), EULER_EMPTY_KEY, c);
                        if (old == EULER_EMPTY_KEY || old == c) break;
                    }
                    h = (h + 1) & bmask;
                }
                if (probe == 4096) bfail = true;
                else a.bvals[2 * h + (s.own_p ? 0 : 1)] = s.own_p ? id_rp : id_s;   // [0]: id(rc prefix) from the prefix owner, [1]: id(suffix) from the suffix owner
            }
        }
        if (bfail) atomicOr((unsigned long long *)(a.stats + 2), (unsigned long long)BKT_FLAG_BOUNDARY);
    }
}

// resolve the suffix vertex of the edges that cross buckets (ev2 == NO_ID after the build)
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
        u64 h = ((c ^ (c >> 29)) * 0x9E3779B97F4A7C15ull >> 20) & bmask;
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

// ---- bucket order: largest first -----------------------------------------------------------------------------------
// A bucket can pass the look-back only when every earlier ticket has published its totals, so tickets are handed
// out in descending size: whoever waits, waits for buckets that started earlier AND are no smaller.  Sizes are
// binned (BB_BINS classes of the record count), which is all the order has to be.
__device__ __forceinline__ u32 bb_size_bin(const u32 *counts, u32 b, u32 nranks, u32 rcap)
{
    u64 total = 0;
    for (u32 s = 0; s < nranks; s++) {
        const u32 c = counts[(u64)b * nranks + s];
        total += c < rcap ? c : rcap;
    }
    const u64 full = (u64)rcap * nranks;
    const u32 bin = (u32)(total * (BB_BINS - 1) / (full ? full : 1));
    return (BB_BINS - 1) - (bin > BB_BINS - 1 ? BB_BINS - 1 : bin);   // bin 0 = the largest buckets
}
__global__ void __launch_bounds__(256) bkt_size_hist_kernel(const u32 *__restrict__ counts, u32 nb, u32 nranks, u32 rcap, u32 *__restrict__ hist)
{
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) atomicAdd(hist + bb_size_bin(counts, b, nranks, rcap), 1u);
}
__global__ void __launch_bounds__(256) bkt_order_kernel(const u32 *__restrict__ counts, u32 nb, u32 nranks, u32 rcap, const u32 *__restrict__ hist,
                                                        u32 *__restrict__ fill, u32 *__restrict__ order)
{
    __shared__ u32 s_base[BB_BINS];
    __shared__ u32 s_part[256];
    // exclusive scan of the BB_BINS histogram, recomputed by every block (4 bins per thread)
    const int tid = threadIdx.x;
    u32 v[BB_BINS / 256], sum = 0;
#pragma unroll
    for (int i = 0; i < BB_BINS / 256; i++) { v[i] = hist[tid * (BB_BINS / 256) + i]; sum += v[i]; }
    s_part[tid] = sum;
    __syncthreads();
    u32 off = 0;
    for (int i = 0; i < tid; i++) off += s_part[i];
#pragma unroll
    for (int i = 0; i < BB_BINS / 256; i++) { s_base[tid * (BB_BINS / 256) + i] = off; off += v[i]; }
    __syncthreads();
    const u32 b = blockIdx.x * blockDim.x + tid;
    if (b < nb) {
        const u32 bin = bb_size_bin(counts, b, nranks, rcap);
        order[s_base[bin] + atomicAdd(fill + bin, 1u)] = b;
    }
}

size_t bkt_build_smem(u32 cap) { return (size_t)28 * cap; }

int bkt_build(euler_ctx *ctx, const BktBuild &B)
{
    if (!B.nb) return EULER_OK;
    if (B.cap < BB_THREADS || B.cap % 256) return euler_fail(ctx, EULER_ERR_ARG, "bucket table capacity must be a multiple of 256");
    if (B.bcap & (B.bcap - 1)) return euler_fail(ctx, EULER_ERR_ARG, "boundary table capacity must be a power of two");
    const size_t smem = bkt_build_smem(B.cap);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(bkt_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    // state: flag u32[nb] | ticket u32 | hist u32[BINS] | fill u32[BINS] | (16-byte aligned) agg_uv, agg_e, inc_uv, inc_e u64[nb] | order u32[nb]
    u32 *flag = (u32 *)B.state;
    u32 *ticket = flag + B.nb;
    u32 *hist = ticket + 1, *fill = hist + BB_BINS;
    const size_t zero_bytes = ((size_t)B.nb + 1 + 2 * BB_BINS) * 4;
    u64 *w64 = (u64 *)((char *)B.state + (zero_bytes + 15) / 16 * 16);
    u32 *order = (u32 *)(w64 + 4ull * B.nb);
    CUDA_TRY(ctx, cudaMemsetAsync(B.state, 0, zero_bytes, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(B.bkeys, 0xFF, B.bcap * 8, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(B.bvals, 0xFF, B.bcap * 8, ctx->stream));
    bkt_size_hist_kernel<<<grid_for(B.nb, 256), 256, 0, ctx->stream>>>(B.counts, B.nb, B.nranks, B.rcap, hist);
    bkt_order_kernel<<<grid_for(B.nb, 256), 256, 0, ctx->stream>>>(B.counts, B.nb, B.nranks, B.rcap, hist, fill, order);
    CUDA_TRY(ctx, cudaGetLastError());
    BkBuildArgs a;
    a.records = (const uint4 *)B.records; a.counts = B.counts; a.order = order; a.nb = B.nb; a.nranks = B.nranks; a.rcap = B.rcap; a.l = B.l;
    a.cap = B.cap;
    a.lkeys = B.lkeys; a.lvals = B.lvals; a.loffs = B.loffs; a.ev1 = B.ev1; a.ev2 = B.ev2; a.ucap = B.ucap;
    a.vkeys = B.vkeys; a.lcount = B.lcount; a.ecount = B.ecount; a.lstart = B.lstart; a.estart = B.estart; a.ev = B.ev; a.vcap = B.vcap;
    a.flag = flag; a.ticket = ticket; a.agg_uv = w64; a.agg_e = w64 + B.nb; a.inc_uv = w64 + 2ull * B.nb; a.inc_e = w64 + 3ull * B.nb;
    a.bkeys = B.bkeys; a.bvals = B.bvals; a.bcap = B.bcap; a.stats = B.stats;
    bkt_build_kernel<<<B.nb, BB_THREADS, smem, ctx->stream>>>(a);
    CUDA_TRY(ctx, cudaGetLastError());
    const unsigned g = (unsigned)ctx->num_sms * 8;
    bkt_fixup_kernel<<<g, 256, 0, ctx->stream>>>(B.lkeys, B.ev2, B.stats + 3, B.ucap, B.l, B.bkeys, B.bvals, B.bcap);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

size_t bkt_state_bytes(u32 nb) { return (((size_t)nb + 1 + 2 * BB_BINS) * 4 + 15) / 16 * 16 + (size_t)nb * 32 + (size_t)nb * 4 + 16; }

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
