// bucket_build.cu -- pass 2 of the bucketed hot path (bucket.cuh): one CTA per bucket builds the bucket's part
// of the de Bruijn graph entirely in shared memory and writes every artefact sequentially.
//
// Replaces, per bucket, the host dictionary fill of eulercuda.readLmersKmersCuda (eulercuda.py:141-178), the
// hash table of pygpuhash.create_hash_table_device (pygpuhash.py:262-315) and D1-D6 of
// pydebruijn.construct_debruijn_graph_device (pydebruijn.py:516-620):
//   A  count the canonical l-mers spelled by the bucket's records (shared-memory table, atomicCAS claim,
//      count + "prefix / suffix vertex is ours" bits in one word);
//   B  insert the owned end vertices of every distinct l-mer into a second shared-memory table and add the
//      multiplicity to the vertex's leaving / entering total;
//   C  block scans: edge records and edge offsets over the l-mer slots, ids and degree totals over the vertex slots;
//   D  decoupled look-back over the buckets (ticket order): global bases of (records, vertices, edges);
//   E  vertex artefacts (the eight degree slots of a vertex are eight lookups in the bucket's own l-mer table)
//      and edge artefacts.  The suffix vertex of an edge whose suffix lives in another bucket is published /
//      resolved through a small global table keyed by the canonical l-mer (bkt_fixup_kernel).
#include "bucket.cuh"
#include "kernels.h"

#define BB_THREADS 256
#define BB_WARPS (BB_THREADS / 32)
#define BB_RC 512   // records staged per chunk (8 KB)

struct BkBuildArgs {
    const uint4 *records;
    const u32 *counts;   // [nb * nranks] records in region (bucket, source rank)
    u32 nb, nranks, rcap, l;
    u32 log_capl, log_capv;
    u64 *lkeys; u32 *lvals, *loffs, *ev1, *ev2; u64 ucap;
    u64 *vkeys; u32 *lcount, *ecount, *lstart, *estart; euler_vertex *ev; u64 vcap;
    u32 *flag; u64 *agg_uv, *agg_e, *inc_uv, *inc_e; u32 *ticket;
    u64 *bkeys; u32 *bvals; u64 bcap;
    u64 *stats;
};

__device__ __forceinline__ u32 bb_hash(u64 key, u32 log_cap)
{
    return (u32)(((key ^ (key >> 31)) * 0x9E3779B97F4A7C15ull) >> (64 - log_cap));
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

// insert `key` into a shared-memory open-addressing table; returns the slot, or 0xffffffff when the table is full.
// first = this call claimed the slot.
__device__ __forceinline__ u32 sm_insert(u64 *keys, u32 log_cap, u64 key, bool &first)
{
    const u32 mask = (1u << log_cap) - 1u;
    u32 h = bb_hash(key, log_cap);
    first = false;
    for (u32 probe = 0; probe <= mask; probe++) {
        const u64 cur = keys[h];
        if (cur == key) return h;
        if (cur == EULER_EMPTY_KEY) {
            const u64 old = atomicCAS(keys + h, EULER_EMPTY_KEY, key);
            if (old == EULER_EMPTY_KEY) { first = true; return h; }
            if (old == key) return h;
        }
        h = (h + 1u) & mask;
    }
    return 0xffffffffu;
}
__device__ __forceinline__ u32 sm_find(const u64 *keys, u32 log_cap, u64 key)
{
    const u32 mask = (1u << log_cap) - 1u;
    u32 h = bb_hash(key, log_cap);
    for (u32 probe = 0; probe <= mask; probe++) {
        const u64 cur = keys[h];
        if (cur == key) return h;
        if (cur == EULER_EMPTY_KEY) return 0xffffffffu;
        h = (h + 1u) & mask;
    }
    return 0xffffffffu;
}

// block-wide exclusive scan of one u64 per thread (packed pairs allowed); returns the exclusive prefix, total in *tot
__device__ __forceinline__ u64 bb_block_scan(u64 v, u64 *s_warp /* BB_WARPS + 1 */, u64 *tot)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    __syncthreads();   // s_warp may still be read from the previous scan
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    u64 off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < BB_WARPS; w++) {
        const u64 t = s_warp[w];
        if (w < warp) off += t;
        total += t;
    }
    *tot = total;
    return off + inc - v;
}

// both-strand multiplicity of the l-mer x in the bucket's table (0 when absent)
__device__ __forceinline__ u32 bb_bs_count(const u64 *lt_keys, const u32 *lt_cnt, u32 log_capl, u64 x, u32 l)
{
    const u64 r = bk_revcomp(x, l);
    const u32 slot = sm_find(lt_keys, log_capl, x < r ? x : r);
    if (slot == 0xffffffffu) return 0u;
    const u32 n = lt_cnt[slot] & 0x3fffffffu;
    return x == r ? 2u * n : n;
}

__global__ void __launch_bounds__(BB_THREADS) bkt_build_kernel(const BkBuildArgs a)
{
    extern __shared__ __align__(16) unsigned char bb_smem[];
    const u32 capl = 1u << a.log_capl, capv = 1u << a.log_capv;
    u64 *lt_keys = (u64 *)bb_smem;
    u64 *vt_keys = lt_keys + capl;
    uint4 *s_recs = (uint4 *)(vt_keys + capv);
    u32 *lt_cnt = (u32 *)(s_recs + BB_RC);
    u32 *vt_a = lt_cnt + capl;   // leaving total of the canonical strand, later the vertex id
    u32 *vt_b = vt_a + capv;     // entering total of the canonical strand
    __shared__ u64 s_warp[BB_WARPS + 1];
    __shared__ u64 s_base[2];
    __shared__ u32 s_bucket, s_fail;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const u32 l = a.l, k = l - 1;
    const u64 kmask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);

    if (tid == 0) { s_bucket = atomicAdd(a.ticket, 1u); s_fail = 0; }
    for (u32 i = tid; i < capl; i += BB_THREADS) { lt_keys[i] = EULER_EMPTY_KEY; lt_cnt[i] = 0; }
    for (u32 i = tid; i < capv; i += BB_THREADS) { vt_keys[i] = EULER_EMPTY_KEY; vt_a[i] = 0; vt_b[i] = 0; }
    __syncthreads();
    const u32 b = s_bucket;
    if (b >= a.nb) return;   // never: the grid is nb blocks

    // ---- A: count the l-mers of the bucket's records ---------------------------------------------------------
    bool fail = false;
    u32 max_region = 0;
    for (u32 src = 0; src < a.nranks; src++) {
        u32 cnt = a.counts[(u64)b * a.nranks + src];
        max_region = cnt > max_region ? cnt : max_region;
        if (cnt > a.rcap) cnt = a.rcap;   // overflowed region: the run is repeated with a larger capacity anyway
        const uint4 *region = a.records + ((u64)b * a.nranks + src) * a.rcap;
        for (u32 c0 = 0; c0 < cnt; c0 += BB_RC) {
            const u32 nrc = cnt - c0 < BB_RC ? cnt - c0 : BB_RC;
            for (u32 i = tid; i < nrc; i += BB_THREADS) s_recs[i] = ld_stream_v4(region + c0 + i);
            __syncthreads();
            // every warp owns a contiguous slice of the chunk; a lane that has finished its record takes the warp's next one
            const u32 per = (nrc + BB_WARPS - 1) / BB_WARPS;
            u32 next = warp * per;
            const u32 end = next + per < nrc ? next + per : nrc;
            u64 hi = 0, lo = 0;
            u32 hdr = 0, j = 0, nl = 0;
            while (true) {
                const bool need = j >= nl;
                const unsigned nm = __ballot_sync(0xffffffffu, need);
                if (need) {
                    const u32 idx = next + __popc(nm & lt_mask);
                    j = 0;
                    nl = 0;
                    if (idx < end) {
                        const uint4 rec = s_recs[idx];
                        hdr = rec.x;
                        const u32 nb = hdr & 63u;
                        nl = nb > k ? nb - k : 0u;
                        hi = ((u64)rec.y << 32) | rec.z;
                        lo = (u64)rec.w << 32;
                    }
                }
                next += __popc(nm);
                const bool active = j < nl;
                if (!__any_sync(0xffffffffu, active)) break;
                if (active) {
                    const u32 sh = 2 * j;
                    const u64 f = sh ? ((hi << sh) | (lo >> (64 - sh))) : hi;
                    const u64 x = f >> (64 - 2 * l);
                    const u64 r = bk_revcomp(x, l);
                    const bool flip = r < x;
                    const u64 c = flip ? r : x;
                    // ownership of the end vertices, in the orientation the read spells
                    const u32 own_pf = (j == 0 && (hdr & BK_HDR_LFF)) ? 0u : 1u;
                    const u32 own_sf = (j + 1 == nl && (hdr & BK_HDR_RFF)) ? 0u : 1u;
                    const u32 own = flip ? (own_sf | (own_pf << 1)) : (own_pf | (own_sf << 1));   // bit 0: prefix(c), bit 1: suffix(c)
                    bool first;
                    const u32 slot = sm_insert(lt_keys, a.log_capl, c, first);
                    if (slot == 0xffffffffu) fail = true;
                    else atomicAdd(lt_cnt + slot, first ? (1u | (own << 30)) : 1u);
                    j++;
                }
            }
            __syncthreads();
        }
    }
    if (fail) s_fail = 1;
    __syncthreads();

    // ---- B: owned end vertices of every distinct l-mer ---------------------------------------------------------
    for (u32 slot = tid; slot < capl; slot += BB_THREADS) {
        const u64 c = lt_keys[slot];
        if (c == EULER_EMPTY_KEY) continue;
        const u32 w = lt_cnt[slot], n = w & 0x3fffffffu;
        const bool own_p = (w >> 30) & 1u, own_s = (w >> 31) & 1u;
        const bool pal = c == bk_revcomp(c, l);
        const u32 m0 = pal ? 2u * n : n;
        bool first;
        if (own_p) {   // strand c leaves prefix(c) with m0
            const u64 p = c >> 2, rp = bk_revcomp(p, k);
            const u32 vs = sm_insert(vt_keys, a.log_capv, p < rp ? p : rp, first);
            if (vs == 0xffffffffu) fail = true;
            else atomicAdd((p <= rp) ? vt_a + vs : vt_b + vs, m0);   // p is the canonical strand (or a palindrome): its leaving total
        }
        if (own_s && !pal) {   // strand c enters suffix(c) with n (a palindromic l-mer is covered by its prefix side)
            const u64 s = c & kmask, rs = bk_revcomp(s, k);
            const u32 vs = sm_insert(vt_keys, a.log_capv, s < rs ? s : rs, first);
            if (vs == 0xffffffffu) fail = true;
            else if (s == rs) atomicAdd(vt_a + vs, n);          // palindromic vertex: one strand, leaving total == entering total
            else atomicAdd((s < rs) ? vt_b + vs : vt_a + vs, n);   // canonical strand: entering; else the mirror = leaving of the canonical strand
        }
    }
    if (fail) s_fail = 1;
    __syncthreads();
    const bool failed = s_fail != 0;

    // ---- C: scans (every thread owns a contiguous range of slots) ------------------------------------------------
    const u32 lpt = capl / BB_THREADS, vpt = capv / BB_THREADS;
    u64 my_rec = 0, my_edges = 0, my_v = 0, my_w = 0;
    if (!failed) {
        for (u32 slot = tid * lpt; slot < (tid + 1) * lpt; slot++) {
            const u64 c = lt_keys[slot];
            if (c == EULER_EMPTY_KEY) continue;
            const u32 w = lt_cnt[slot], n = w & 0x3fffffffu;
            const u32 own_p = (w >> 30) & 1u, own_s = (w >> 31) & 1u;
            const bool pal = c == bk_revcomp(c, l);
            if (pal) { my_rec += own_p; my_edges += own_p ? 2ull * n : 0ull; }
            else { my_rec += own_p + own_s; my_edges += (u64)n * (own_p + own_s); }
        }
        for (u32 slot = tid * vpt; slot < (tid + 1) * vpt; slot++) {
            const u64 v = vt_keys[slot];
            if (v == EULER_EMPTY_KEY) continue;
            const bool palv = v == bk_revcomp(v, k);
            my_v += palv ? 1u : 2u;
            my_w += palv ? (u64)vt_a[slot] : (u64)vt_a[slot] + vt_b[slot];
        }
    }
    u64 tot_uv, tot_e, tot_w;
    const u64 ex_uv = bb_block_scan((my_v << 32) | my_rec, s_warp, &tot_uv);
    const u64 ex_e = bb_block_scan(my_edges, s_warp, &tot_e);
    const u64 ex_w = bb_block_scan(my_w, s_warp, &tot_w);

    // ---- D: look-back over the buckets ------------------------------------------------------------------------------
    if (warp == 0) {
        u64 pre_uv = 0, pre_e = 0;
        if (b == 0) {
            if (lane == 0) {
                st_vol_u64(a.inc_uv, tot_uv);
                st_vol_u64(a.inc_e, tot_e);
                __threadfence();
                st_vol_u32(a.flag, 2u);
            }
        } else {
            if (lane == 0) {
                st_vol_u64(a.agg_uv + b, tot_uv);
                st_vol_u64(a.agg_e + b, tot_e);
                __threadfence();
                st_vol_u32(a.flag + b, 1u);
            }
            long long look = (long long)b - 1;
            while (true) {
                const long long idx = look - lane;
                u32 f = 2u;
                u64 vuv = 0, ve = 0;
                if (idx >= 0) {
                    u32 spins = 0;
                    do {
                        f = ld_vol_u32(a.flag + idx);
                        if (f == 0u && ++spins > (1u << 26)) {   // a predecessor never published: give up instead of hanging the GPU
                            atomicOr((unsigned long long *)(a.stats + 2), (unsigned long long)BKT_FLAG_INTERNAL);
                            f = 3u;
                        }
                    } while (f == 0u);
                    __threadfence();
                    if (f != 3u) {
                        vuv = ld_vol_u64((f == 2u ? a.inc_uv : a.agg_uv) + idx);
                        ve = ld_vol_u64((f == 2u ? a.inc_e : a.agg_e) + idx);
                    } else {
                        f = 2u;   // stop the look-back here
                    }
                }
                const unsigned inc_mask = __ballot_sync(0xffffffffu, f == 2u);
                if (inc_mask) {
                    const int firsti = __ffs(inc_mask) - 1;
                    if (lane > firsti) { vuv = 0; ve = 0; }
                }
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    vuv += __shfl_xor_sync(0xffffffffu, vuv, d);
                    ve += __shfl_xor_sync(0xffffffffu, ve, d);
                }
                pre_uv += vuv;
                pre_e += ve;
                if (inc_mask) break;
                look -= 32;
            }
            if (lane == 0) {
                st_vol_u64(a.inc_uv + b, pre_uv + tot_uv);
                st_vol_u64(a.inc_e + b, pre_e + tot_e);
                __threadfence();
                st_vol_u32(a.flag + b, 2u);
            }
        }
        if (lane == 0) {
            s_base[0] = pre_uv;
            s_base[1] = pre_e;
            if (b == a.nb - 1) {   // grand totals
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
    if (ubase + (tot_uv & 0xffffffffull) > a.ucap || vbase + (tot_uv >> 32) > a.vcap) {
        if (tid == 0) atomicOr((unsigned long long *)(a.stats + 2), (unsigned long long)BKT_FLAG_OUTPUT);
        return;
    }

    // ---- E1: vertex artefacts -----------------------------------------------------------------------------------------
    {
        u64 vid = vbase + (ex_uv >> 32), wpos = ebase + ex_w;
        bool bad = false;
        for (u32 slot = tid * vpt; slot < (tid + 1) * vpt; slot++) {
            const u64 v = vt_keys[slot];
            if (v == EULER_EMPTY_KEY) continue;
            const u64 rv = bk_revcomp(v, k);
            const bool palv = v == rv;
            const u32 L0 = vt_a[slot], E0 = vt_b[slot];
            u32 lc[4], ec[4];
#pragma unroll
            for (u32 t = 0; t < 4; t++) {
                lc[t] = bb_bs_count(lt_keys, lt_cnt, a.log_capl, (v << 2) | t, l);
                ec[t] = bb_bs_count(lt_keys, lt_cnt, a.log_capl, ((u64)t << (2 * k)) | v, l);
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
                vid += 2;
                wpos += (u64)ls + es;
            } else {
                vid += 1;
                wpos += ls;
            }
        }
        if (bad) atomicOr((unsigned long long *)(a.stats + 2), (unsigned long long)BKT_FLAG_INTERNAL);
    }
    __syncthreads();

    // ---- E2: edge artefacts -------------------------------------------------------------------------------------------
    {
        u64 rec = ubase + (ex_uv & 0xffffffffull), eo = ebase + ex_e;
        bool bfail = false;
        for (u32 slot = tid * lpt; slot < (tid + 1) * lpt; slot++) {
            const u64 c = lt_keys[slot];
            if (c == EULER_EMPTY_KEY) continue;
            const u32 w = lt_cnt[slot], n = w & 0x3fffffffu;
            const bool own_p = (w >> 30) & 1u, own_s = (w >> 31) & 1u;
            const u64 r = bk_revcomp(c, l);
            const bool pal = c == r;
            const u32 m0 = pal ? 2u * n : n;
            const u64 p = c >> 2, s = c & kmask, rp = bk_revcomp(p, k), rs = bk_revcomp(s, k);
            u32 id_p = EULER_NO_ID, id_rp = EULER_NO_ID, id_s = EULER_NO_ID, id_rs = EULER_NO_ID;
            if (own_p) {
                const u32 vs = sm_find(vt_keys, a.log_capv, p < rp ? p : rp);
                if (vs == 0xffffffffu) { bfail = true; continue; }   // cannot happen: inserted in B
                const u32 i0 = vt_a[vs];
                id_p = p <= rp ? i0 : i0 + 1u;
                id_rp = rp <= p ? i0 : i0 + 1u;
            }
            if (own_s) {
                const u32 vs = sm_find(vt_keys, a.log_capv, s < rs ? s : rs);
                if (vs == 0xffffffffu) { bfail = true; continue; }
                const u32 i0 = vt_a[vs];
                id_s = s <= rs ? i0 : i0 + 1u;
                id_rs = rs <= s ? i0 : i0 + 1u;
            }
            if (own_p) {   // strand c is homed with its prefix vertex
                a.lkeys[rec] = c; a.lvals[rec] = m0; a.loffs[rec] = (u32)eo; a.ev1[rec] = id_p; a.ev2[rec] = id_s;
                rec++;
                eo += m0;
            }
            if (own_s && !pal) {   // strand rc(c) runs from rc(suffix c) to rc(prefix c)
                a.lkeys[rec] = r; a.lvals[rec] = n; a.loffs[rec] = (u32)eo; a.ev1[rec] = id_rs; a.ev2[rec] = id_rp;
                rec++;
                eo += n;
            }
            if (own_p != own_s) {   // the other end vertex lives in another bucket: publish our side's id under the canonical l-mer
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
                else a.bvals[2 * h + (own_p ? 0 : 1)] = own_p ? id_rp : id_s;   // [0]: id(rc prefix) from the prefix owner, [1]: id(suffix) from the suffix owner
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

size_t bkt_build_smem(u32 log_capl, u32 log_capv)
{
    return ((size_t)12 << log_capl) + ((size_t)16 << log_capv) + (size_t)BB_RC * 16;
}

int bkt_build(euler_ctx *ctx, const BktBuild &B)
{
    if (!B.nb) return EULER_OK;
    if ((1u << B.log_capl) < BB_THREADS || (1u << B.log_capv) < BB_THREADS) return euler_fail(ctx, EULER_ERR_ARG, "bucket tables smaller than the block");
    if (B.bcap & (B.bcap - 1)) return euler_fail(ctx, EULER_ERR_ARG, "boundary table capacity must be a power of two");
    const size_t smem = bkt_build_smem(B.log_capl, B.log_capv);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(bkt_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    // look-back state: flag u32[nb] | ticket u32 | agg_uv, agg_e, inc_uv, inc_e u64[nb]
    u32 *flag = (u32 *)B.state;
    u32 *ticket = flag + B.nb;
    u64 *w64 = (u64 *)((char *)B.state + (((size_t)B.nb + 1) * 4 + 15) / 16 * 16);
    CUDA_TRY(ctx, cudaMemsetAsync(B.state, 0, ((size_t)B.nb + 1) * 4, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(B.bkeys, 0xFF, B.bcap * 8, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(B.bvals, 0xFF, B.bcap * 8, ctx->stream));
    BkBuildArgs a;
    a.records = (const uint4 *)B.records; a.counts = B.counts; a.nb = B.nb; a.nranks = B.nranks; a.rcap = B.rcap; a.l = B.l;
    a.log_capl = B.log_capl; a.log_capv = B.log_capv;
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

size_t bkt_state_bytes(u32 nb) { return (((size_t)nb + 1) * 4 + 15) / 16 * 16 + (size_t)nb * 32; }

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
