// graph.cu -- vertex table, id assignment and de Bruijn graph construction (D1-D6).
// Replaces pygpuhash.py H1-H4, pydebruijn.py getHashValue/debruijnCount/setupVertices/setupEdges.
#include "kernels.h"
#include "scan.cuh"

#define GB 256  // block size for the element-wise kernels here

int graph_table_clear(euler_ctx *ctx, u64 *keys, u32 *vals, u64 cap)
{
    CUDA_TRY(ctx, cudaMemsetAsync(keys, 0xFF, cap * sizeof(u64), ctx->stream));
    if (vals) CUDA_TRY(ctx, cudaMemsetAsync(vals, 0, cap * sizeof(u32), ctx->stream));
    return EULER_OK;
}

__device__ __forceinline__ u64 canon64(u64 x, u32 len)
{
    const u64 r = revcomp64(x, len);
    return x < r ? x : r;
}

// ---- vertex table build ------------------------------------------------------------------------
__global__ void __launch_bounds__(GB) vertex_insert_kernel(const u64 *__restrict__ lt_keys, u64 lt_cap, u32 l,
                                                            u64 *__restrict__ vt_keys, u64 vt_cap, TableHash vth, u64 *flags)
{
    const u64 slot = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= lt_cap) return;
    const u64 key = lt_keys[slot];
    if (key == EULER_EMPTY_KEY) return;
    const u32 k = l - 1;
    const u64 kmask = key_mask_d(k);
    const u64 max_probe = vt_cap / EULER_BUCKET < 4096 ? vt_cap / EULER_BUCKET : 4096;
    const u32 nb = (u32)(vt_cap / EULER_BUCKET);
    const u64 cp = canon64(key >> 2, k), cs = canon64(key & kmask, k);
    u32 sa = 0, sp = 0, ss = 0;
    if (vth.span_nb) min_scores(key, l, vth.m, sa, sp, ss);   // prefix / suffix minimizers from one pass over the l-mer
    const u64 a = table_insert_at(vt_keys, vt_cap, cp, home_from_score(cp, sp, nb, vth), max_probe);
    const u64 b = table_insert_at(vt_keys, vt_cap, cs, home_from_score(cs, ss, nb, vth), max_probe);
    if (a == EULER_NO_SLOT || b == EULER_NO_SLOT) atomicOr((unsigned long long *)flags, 2ull);
}

int graph_vertex_insert(euler_ctx *ctx, const u64 *lt_keys, u64 lt_cap, u32 l, u64 *vt_keys, u64 vt_cap, TableHash vth,
                        u64 *d_flags)
{
    vertex_insert_kernel<<<grid_for(lt_cap, GB), GB, 0, ctx->stream>>>(lt_keys, lt_cap, l, vt_keys, vt_cap, vth, d_flags);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- strand-weight scan over slots -------------------------------------------------------------
struct SlotWeight {
    const u64 *keys;
    u32 len;
    __device__ __forceinline__ u32 operator()(u64 i) const
    {
        const u64 k = keys[i];
        if (k == EULER_EMPTY_KEY) return 0u;
        return k == revcomp64(k, len) ? 1u : 2u;
    }
};

int graph_slot_scan(euler_ctx *ctx, const u64 *keys, u64 cap, u32 len, u32 *d_base, u64 *d_total)
{
    return scan_exclusive(ctx, SlotWeight{keys, len}, cap, d_base, d_total);
}

__global__ void __launch_bounds__(GB) compact_lmers_kernel(const u64 *__restrict__ lt_keys, const u32 *__restrict__ lt_cnt,
                                                            const u32 *__restrict__ base, u64 cap, u32 l,
                                                            u64 *__restrict__ lkeys, u32 *__restrict__ lvals)
{
    const u64 slot = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= cap) return;
    const u64 c = lt_keys[slot];
    if (c == EULER_EMPTY_KEY) return;
    const u32 n = lt_cnt[slot];
    const u32 idx = base[slot];
    const u64 r = revcomp64(c, l);
    if (c == r) {  // palindrome: both strands hit the same key (eulercuda.py:151-161)
        lkeys[idx] = c;
        lvals[idx] = 2u * n;
    } else {
        lkeys[idx] = c;
        lvals[idx] = n;
        lkeys[idx + 1] = r;
        lvals[idx + 1] = n;
    }
}

int graph_compact_lmers(euler_ctx *ctx, const u64 *lt_keys, const u32 *lt_cnt, const u32 *lt_base, u64 lt_cap, u32 l,
                        u64 *lkeys, u32 *lvals)
{
    compact_lmers_kernel<<<grid_for(lt_cap, GB), GB, 0, ctx->stream>>>(lt_keys, lt_cnt, lt_base, lt_cap, l, lkeys, lvals);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

__global__ void __launch_bounds__(GB) bucket_bases_kernel(const u32 *__restrict__ id0, u64 nb, u32 *__restrict__ bbase)
{
    const u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nb) bbase[b] = id0[b * EULER_BUCKET];
}
int graph_bucket_bases(euler_ctx *ctx, const u32 *id0, u64 cap, u32 *bbase)
{
    const u64 nb = cap / EULER_BUCKET;
    if (!nb) return EULER_OK;
    bucket_bases_kernel<<<grid_for(nb, GB), GB, 0, ctx->stream>>>(id0, nb, bbase);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

__global__ void __launch_bounds__(GB) compact_vertices_kernel(const u64 *__restrict__ vt_keys, const u32 *__restrict__ base,
                                                               u64 cap, u32 k, u64 *__restrict__ vkeys)
{
    const u64 slot = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= cap) return;
    const u64 c = vt_keys[slot];
    if (c == EULER_EMPTY_KEY) return;
    const u32 idx = base[slot];
    const u64 r = revcomp64(c, k);
    vkeys[idx] = c;
    if (c != r) vkeys[idx + 1] = r;
}

int graph_compact_vertices(euler_ctx *ctx, const u64 *vt_keys, const u32 *vt_base, u64 vt_cap, u32 k, u64 *vkeys)
{
    compact_vertices_kernel<<<grid_for(vt_cap, GB), GB, 0, ctx->stream>>>(vt_keys, vt_base, vt_cap, k, vkeys);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

__global__ void __launch_bounds__(GB) assign_sorted_ids_kernel(const u64 *__restrict__ vkeys, u64 nv,
                                                                const u64 *__restrict__ vt_keys, u64 vt_cap, u32 k,
                                                                TableHash vth, u32 *__restrict__ id0, u32 *__restrict__ id1)
{
    const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nv) return;
    const u64 x = vkeys[r];
    const u64 rc = revcomp64(x, k);
    const u64 c = x < rc ? x : rc;
    const u64 slot = table_find_at(vt_keys, vt_cap, c, table_home(c, k, (u32)(vt_cap / EULER_BUCKET), vth));
    if (slot == EULER_NO_SLOT) return;
    if (x == c) id0[slot] = (u32)r;
    if (x == rc || x != c) id1[slot] = (u32)r;
}

int graph_assign_sorted_ids(euler_ctx *ctx, const u64 *vkeys, u64 nv, const u64 *vt_keys, u64 vt_cap, u32 k, TableHash vth,
                            u32 *id0, u32 *id1)
{
    if (!nv) return EULER_OK;
    assign_sorted_ids_kernel<<<grid_for(nv, GB), GB, 0, ctx->stream>>>(vkeys, nv, vt_keys, vt_cap, k, vth, id0, id1);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- D1: degree slots + compressed edges -------------------------------------------------------
__device__ __forceinline__ u32 vt_lookup(const VertexTable &vt, u64 v)
{
    const u64 r = revcomp64(v, vt.k);
    const u64 c = v < r ? v : r;
    const u64 slot = table_find_at(vt.keys, vt.cap, c, table_home(c, vt.k, (u32)(vt.cap / EULER_BUCKET), vt.th));
    if (slot == EULER_NO_SLOT) return EULER_NO_ID;
    if (v == c) return vt.id0[slot];
    return vt.id1 ? vt.id1[slot] : vt.id0[slot] + 1u;
}
__device__ __forceinline__ u32 pt_lookup(const PlainTable &pt, u64 v)
{
    // getHashValue pydebruijn.py:57-88: value or MAX_INT
    const u64 slot = table_find(pt.keys, pt.cap, v);
    return slot == EULER_NO_SLOT ? EULER_NO_ID : pt.vals[slot];
}

template <typename Table, typename Lookup>
__global__ void __launch_bounds__(GB) degree_slots_kernel(const u64 *__restrict__ lkeys, const u32 *__restrict__ lvals,
                                                           u64 nl, u32 l, Table tab, Lookup lookup, u64 size,
                                                           u32 *__restrict__ lcount, u32 *__restrict__ ecount,
                                                           u32 *__restrict__ ev1, u32 *__restrict__ ev2)
{
    // debruijnCount pydebruijn.py:107-141
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nl) return;
    const u64 x = lkeys[i];
    const u32 m = lvals[i];
    const u32 k = l - 1;
    const u64 kmask = key_mask_d(k);
    const u32 pid = lookup(tab, x >> 2);
    const u32 sid = lookup(tab, x & kmask);
    const u32 to = (u32)(x & 3), from = (u32)((x >> (2 * k)) & 3);
    const u64 to_index = ((u64)pid << 2) + to, from_index = ((u64)sid << 2) + from;
    if (pid != EULER_NO_ID && to_index < size) lcount[to_index] = m;
    if (sid != EULER_NO_ID && from_index < size) ecount[from_index] = m;
    if (ev1) { ev1[i] = pid; ev2[i] = sid; }
}

struct VtLookupFn { __device__ __forceinline__ u32 operator()(const VertexTable &t, u64 v) const { return vt_lookup(t, v); } };
struct PtLookupFn { __device__ __forceinline__ u32 operator()(const PlainTable &t, u64 v) const { return pt_lookup(t, v); } };

int graph_degree_slots(euler_ctx *ctx, const u64 *lkeys, const u32 *lvals, u64 nl, u32 l, const VertexTable &vt,
                       u32 *lcount, u32 *ecount, u32 *ev1, u32 *ev2)
{
    if (!nl) return EULER_OK;
    degree_slots_kernel<<<grid_for(nl, GB), GB, 0, ctx->stream>>>(lkeys, lvals, nl, l, vt, VtLookupFn(), ~0ull, lcount,
                                                                  ecount, ev1, ev2);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

int graph_degree_slots_plain(euler_ctx *ctx, const u64 *lkeys, const u32 *lvals, u64 nl, u32 l, const PlainTable &pt,
                             u64 nv, u32 *lcount, u32 *ecount)
{
    if (!nl) return EULER_OK;
    degree_slots_kernel<<<grid_for(nl, GB), GB, 0, ctx->stream>>>(lkeys, lvals, nl, l, pt, PtLookupFn(), 4 * nv, lcount,
                                                                  ecount, (u32 *)nullptr, (u32 *)nullptr);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- D5: vertices ------------------------------------------------------------------------------
__device__ __forceinline__ void fill_vertex(euler_vertex *ev, u64 id, u64 key, const u32 *lcount, const u32 *lstart,
                                            const u32 *ecount, const u32 *estart)
{
    // setupVertices pydebruijn.py:280-294
    const uint4 lc = *reinterpret_cast<const uint4 *>(lcount + 4 * id);
    const uint4 ec = *reinterpret_cast<const uint4 *>(ecount + 4 * id);
    euler_vertex v;
    v.vid = key;
    v.lp = lstart[4 * id];
    v.lcount = lc.x + lc.y + lc.z + lc.w;
    v.ep = estart[4 * id];
    v.ecount = ec.x + ec.y + ec.z + ec.w;
    ev[id] = v;
}

__global__ void __launch_bounds__(GB) setup_vertices_kernel(const u64 *__restrict__ vkeys, u64 nv,
                                                             const u32 *__restrict__ lcount, const u32 *__restrict__ lstart,
                                                             const u32 *__restrict__ ecount, const u32 *__restrict__ estart,
                                                             euler_vertex *__restrict__ ev)
{
    const u64 id = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nv) return;
    fill_vertex(ev, id, vkeys[id], lcount, lstart, ecount, estart);
}

int graph_setup_vertices(euler_ctx *ctx, const u64 *vkeys, u64 nv, const u32 *lcount, const u32 *lstart,
                         const u32 *ecount, const u32 *estart, euler_vertex *ev)
{
    if (!nv) return EULER_OK;
    setup_vertices_kernel<<<grid_for(nv, GB), GB, 0, ctx->stream>>>(vkeys, nv, lcount, lstart, ecount, estart, ev);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

__global__ void __launch_bounds__(GB) setup_vertices_plain_kernel(const u64 *__restrict__ kmer_keys, u64 nk, PlainTable pt,
                                                                   u64 nv, const u32 *__restrict__ lcount,
                                                                   const u32 *__restrict__ lstart,
                                                                   const u32 *__restrict__ ecount,
                                                                   const u32 *__restrict__ estart,
                                                                   euler_vertex *__restrict__ ev)
{
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nk) return;
    const u64 key = kmer_keys[t];
    const u32 index = pt_lookup(pt, key);
    if (index != EULER_NO_ID && index < nv) fill_vertex(ev, index, key, lcount, lstart, ecount, estart);
}

int graph_setup_vertices_plain(euler_ctx *ctx, const u64 *kmer_keys, u64 nk, const PlainTable &pt, u64 nv,
                               const u32 *lcount, const u32 *lstart, const u32 *ecount, const u32 *estart,
                               euler_vertex *ev)
{
    if (!nk) return EULER_OK;
    setup_vertices_plain_kernel<<<grid_for(nk, GB), GB, 0, ctx->stream>>>(kmer_keys, nk, pt, nv, lcount, lstart, ecount,
                                                                          estart, ev);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- D6: expanded edges ------------------------------------------------------------------------
// One warp per group of 32 distinct l-mers; the multiplicity loop is spread over the lanes so
// that the ee[] / l[] / e[] writes of one l-mer are contiguous (setupEdges pydebruijn.py:426-475).
template <bool PLAIN>
__global__ void __launch_bounds__(GB) setup_edges_kernel(const u64 *__restrict__ lkeys, const u32 *__restrict__ lvals,
                                                          const u32 *__restrict__ loffs, u64 nl, u32 l,
                                                          const u32 *__restrict__ ev1, const u32 *__restrict__ ev2,
                                                          PlainTable pt, const u32 *__restrict__ lstart,
                                                          const u32 *__restrict__ estart, u32 ecount,
                                                          euler_edge *__restrict__ ee, u32 *__restrict__ lev,
                                                          u32 *__restrict__ ent, const unsigned char *__restrict__ tf)
{
    // tf != NULL (128-bit keys, wide.cu): first / last base codes come precomputed, lkeys is not read
    const int lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 i = warp * 32 + lane;
    u32 m = 0, off = 0, pid = EULER_NO_ID, sid = EULER_NO_ID, lo = 0, eo = 0;
    if (i < nl) {
        const u32 k = l - 1;
        const u64 x = tf ? 0ull : lkeys[i];
        const u32 to = tf ? (tf[i] & 3u) : (u32)(x & 3), from = tf ? (tf[i] >> 2) : (u32)((x >> (2 * k)) & 3);
        m = lvals[i];
        off = loffs[i];
        if (PLAIN) {
            pid = pt_lookup(pt, x >> 2);
            sid = pt_lookup(pt, x & key_mask_d(k));
        } else {
            pid = ev1[i];
            sid = ev2[i];
        }
        if (pid == EULER_NO_ID || sid == EULER_NO_ID) m = 0;
        else {
            lo = lstart[((u64)pid << 2) + to];
            eo = estart[((u64)sid << 2) + from];
        }
    }
    for (int src = 0; src < 32; src++) {
        const u32 mm = __shfl_sync(0xffffffffu, m, src);
        if (!mm) continue;
        const u32 o = __shfl_sync(0xffffffffu, off, src);
        const u32 p = __shfl_sync(0xffffffffu, pid, src);
        const u32 s = __shfl_sync(0xffffffffu, sid, src);
        const u32 a = __shfl_sync(0xffffffffu, lo, src);
        const u32 b = __shfl_sync(0xffffffffu, eo, src);
        for (u32 t = lane; t < mm; t += 32) {
            const u32 id = o + t;
            if (id >= ecount) break;
            euler_edge e;
            e.eid = id; e.v1 = p; e.v2 = s; e.s = ecount; e.pad = 0;
            ee[id] = e;
            lev[a + t] = id;
            ent[b + t] = id;
        }
    }
}

int graph_setup_edges(euler_ctx *ctx, const u64 *lkeys, const u32 *lvals, const u32 *loffs, u64 nl, u32 l,
                      const u32 *ev1, const u32 *ev2, const u32 *lstart, const u32 *estart, u32 ecount,
                      euler_edge *ee, u32 *lev, u32 *ent, const unsigned char *tf)
{
    if (!nl) return EULER_OK;
    PlainTable none = {nullptr, nullptr, 0};
    setup_edges_kernel<false><<<grid_for(nl, GB), GB, 0, ctx->stream>>>(lkeys, lvals, loffs, nl, l, ev1, ev2, none, lstart,
                                                                        estart, ecount, ee, lev, ent, tf);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

int graph_setup_edges_plain(euler_ctx *ctx, const u64 *lkeys, const u32 *lvals, const u32 *loffs, u64 nl, u32 l,
                            const PlainTable &pt, const u32 *lstart, const u32 *estart, u32 ecount, euler_edge *ee,
                            u32 *lev, u32 *ent)
{
    if (!nl) return EULER_OK;
    setup_edges_kernel<true><<<grid_for(nl, GB), GB, 0, ctx->stream>>>(lkeys, lvals, loffs, nl, l, nullptr, nullptr, pt,
                                                                       lstart, estart, ecount, ee, lev, ent, nullptr);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- 64-bit sum of a u32 array (true edge total when it does not fit the 32-bit scan field) -------
__global__ void __launch_bounds__(GB) sum_u32_kernel(const u32 *__restrict__ v, u64 n, u64 *__restrict__ out)
{
    u64 acc = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) acc += v[i];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd((unsigned long long *)out, (unsigned long long)acc);
}
int graph_sum_u32(euler_ctx *ctx, const u32 *v, u64 n, u64 *d_out)
{
    CUDA_TRY(ctx, cudaMemsetAsync(d_out, 0, sizeof(u64), ctx->stream));
    if (!n) return EULER_OK;
    u64 grid = (u64)ctx->num_sms * 8;
    const u64 need = (n + GB - 1) / GB;
    if (grid > need) grid = need;
    sum_u32_kernel<<<(unsigned)grid, GB, 0, ctx->stream>>>(v, n, d_out);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- plain table (module-level gpuhash API) ----------------------------------------------------
__global__ void __launch_bounds__(GB) plain_build_kernel(const u64 *__restrict__ keys, const u32 *__restrict__ vals, u64 n,
                                                          u64 *__restrict__ TK, u32 *__restrict__ TV, u64 cap, u64 *flags)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 slot = table_insert(TK, cap, keys[i], cap / EULER_BUCKET);
    if (slot == EULER_NO_SLOT) { atomicOr((unsigned long long *)flags, 1ull); return; }
    TV[slot] = vals[i];
}

int graph_plain_build(euler_ctx *ctx, const u64 *keys, const u32 *vals, u64 n, u64 *TK, u32 *TV, u64 cap, u64 *d_flags)
{
    CUDA_TRY(ctx, cudaMemsetAsync(TK, 0xFF, cap * sizeof(u64), ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(TV, 0xFF, cap * sizeof(u32), ctx->stream));
    if (!n) return EULER_OK;
    plain_build_kernel<<<grid_for(n, GB), GB, 0, ctx->stream>>>(keys, vals, n, TK, TV, cap, d_flags);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

__global__ void __launch_bounds__(GB) plain_lookup_kernel(PlainTable pt, const u64 *__restrict__ q, u64 nq,
                                                           u32 *__restrict__ out)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    out[i] = pt_lookup(pt, q[i]);
}

int graph_plain_lookup(euler_ctx *ctx, const PlainTable &pt, const u64 *q, u64 nq, u32 *out)
{
    if (!nq) return EULER_OK;
    plain_lookup_kernel<<<grid_for(nq, GB), GB, 0, ctx->stream>>>(pt, q, nq, out);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- count filter (referenceAssembler.build limit, :37-39) -------------------------------------
struct KeepWeight {
    const u32 *vals;
    u32 limit;
    __device__ __forceinline__ u32 operator()(u64 i) const { return vals[i] > limit ? 1u : 0u; }
};

__global__ void __launch_bounds__(GB) filter_scatter_kernel(const u64 *__restrict__ keys, const u32 *__restrict__ vals, u64 n,
                                                             u32 limit, const u32 *__restrict__ base,
                                                             u64 *__restrict__ out_keys, u32 *__restrict__ out_vals)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 v = vals[i];
    if (v > limit) {
        out_keys[base[i]] = keys[i];
        out_vals[base[i]] = v;
    }
}

int graph_filter_counts(euler_ctx *ctx, const u64 *keys, const u32 *vals, u64 n, u32 limit, u32 *d_base, u64 *out_keys,
                        u32 *out_vals, u64 *d_total)
{
    EULER_TRY(scan_exclusive(ctx, KeepWeight{vals, limit}, n, d_base, d_total));
    if (!n) return EULER_OK;
    filter_scatter_kernel<<<grid_for(n, GB), GB, 0, ctx->stream>>>(keys, vals, n, limit, d_base, out_keys, out_vals);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ================================================================================================
// Fused fast path (slot-order ids): three passes instead of nine.
// ================================================================================================

// ---- pair scan over l-mer table slots: distinct both-strand l-mers (lo) and edge offsets (hi) ----
struct LtScanPolicy {
    typedef u64 T;
    const u64 *keys;
    const u32 *cnt;
    u32 l;
    u32 *base, *eoff;
    __device__ __forceinline__ u64 load(u64 i) const
    {
        const u64 k = keys[i];
        if (k == EULER_EMPTY_KEY) return 0ull;
        const u64 m = 2ull * cnt[i];                        // both strands together (n + n, or 2n on a palindrome)
        return (m << 32) | (k == revcomp64(k, l) ? 1ull : 2ull);
    }
    __device__ __forceinline__ void store(u64 i, u64 ex, u64, bool valid) const
    {
        if (valid) { base[i] = (u32)ex; eoff[i] = (u32)(ex >> 32); }
    }
};

int graph_lt_scan(euler_ctx *ctx, const u64 *lt_keys, const u32 *lt_cnt, u64 cap, u32 l, u32 *base, u32 *eoff,
                  u64 *d_total_packed)
{
    return scan_run(ctx, LtScanPolicy{lt_keys, lt_cnt, l, base, eoff}, cap, d_total_packed);
}

// ---- edge kernel: one thread per table slot ------------------------------------------------------
// Emits both strands of the slot's canonical l-mer: lmerKeys / lmerValues / lmerOffsets, the
// compressed edge (v1, v2) and the four degree slots (debruijnCount pydebruijn.py:107-141).
__global__ void __launch_bounds__(GB) edges_fused_kernel(const u64 *__restrict__ lt_keys, const u32 *__restrict__ lt_cnt,
                                                          const u32 *__restrict__ base, const u32 *__restrict__ eoff, u64 cap,
                                                          u32 l, VertexTable vt, u64 *__restrict__ lkeys, u32 *__restrict__ lvals,
                                                          u32 *__restrict__ loffs, u32 *__restrict__ ev1, u32 *__restrict__ ev2,
                                                          DegOut dg)
{
    const u64 slot = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= cap) return;
    const u64 c = lt_keys[slot];
    if (c == EULER_EMPTY_KEY) return;
    const u32 n = lt_cnt[slot];
    const u32 idx = base[slot], eo = eoff[slot];
    const u32 k = l - 1;
    const u64 kmask = key_mask_d(k);
    const u64 r = revcomp64(c, l);
    const bool pal = c == r;
    // vertex ids of prefix / suffix on both strands from two table probes
    const u64 p = c >> 2, s = c & kmask;
    const u64 rp = revcomp64(p, k), rs = revcomp64(s, k);
    const u64 cp = p < rp ? p : rp, cs = s < rs ? s : rs;
    const u32 vnb = (u32)(vt.cap / EULER_BUCKET);
    u32 sca = 0, scp = 0, scs = 0;
    if (vt.th.span_nb) min_scores(c, l, vt.th.m, sca, scp, scs);
    u32 p0, s0, p1, s1;
    if (vt.bbase) {   // slot-order ids from the per-bucket bases (no id0[] sector)
        p0 = table_find_id(vt.keys, vt.cap, cp, home_from_score(cp, scp, vnb, vt.th), vt.bbase, k);
        s0 = table_find_id(vt.keys, vt.cap, cs, home_from_score(cs, scs, vnb, vt.th), vt.bbase, k);
        if (p0 == EULER_NO_ID || s0 == EULER_NO_ID) return;  // cannot happen: both were inserted from this slot
        p1 = p == rp ? p0 : p0 + 1u;
        s1 = s == rs ? s0 : s0 + 1u;
    } else {
        const u64 sp = table_find_at(vt.keys, vt.cap, cp, home_from_score(cp, scp, vnb, vt.th));
        const u64 ss = table_find_at(vt.keys, vt.cap, cs, home_from_score(cs, scs, vnb, vt.th));
        if (sp == EULER_NO_SLOT || ss == EULER_NO_SLOT) return;
        p0 = vt.id0[sp]; s0 = vt.id0[ss];
        p1 = vt.id1 ? vt.id1[sp] : (p == rp ? p0 : p0 + 1u);
        s1 = vt.id1 ? vt.id1[ss] : (s == rs ? s0 : s0 + 1u);
    }
    const u32 id_p = (p == cp) ? p0 : p1, id_rp = (p == cp) ? p1 : p0;   // id(prefix c), id(rc prefix c)
    const u32 id_s = (s == cs) ? s0 : s1, id_rs = (s == cs) ? s1 : s0;
    const u32 m0 = pal ? 2u * n : n;
    lkeys[idx] = c; lvals[idx] = m0; loffs[idx] = eo; ev1[idx] = id_p; ev2[idx] = id_s;
    deg_put_l(dg, id_p, (u32)(c & 3), m0);
    deg_put_e(dg, id_s, id_rs, (u32)((c >> (2 * k)) & 3), m0);
    if (!pal) {  // reverse strand: rc(c) runs from rc(suffix c) to rc(prefix c)
        lkeys[idx + 1] = r; lvals[idx + 1] = n; loffs[idx + 1] = eo + n; ev1[idx + 1] = id_rs; ev2[idx + 1] = id_rp;
        deg_put_l(dg, id_rs, (u32)(r & 3), n);
        deg_put_e(dg, id_rp, id_p, (u32)((r >> (2 * k)) & 3), n);
    }
}

int graph_edges_fused(euler_ctx *ctx, const u64 *lt_keys, const u32 *lt_cnt, const u32 *base, const u32 *eoff, u64 cap,
                      u32 l, const VertexTable &vt, u64 *lkeys, u32 *lvals, u32 *loffs, u32 *ev1, u32 *ev2, u32 *lcount,
                      u32 *ecount, u32 *deg)
{
    if (deg && vt.id1) return euler_fail(ctx, EULER_ERR_ARG, "paired degree layout needs slot-order ids");
    edges_fused_kernel<<<grid_for(cap, GB), GB, 0, ctx->stream>>>(lt_keys, lt_cnt, base, eoff, cap, l, vt, lkeys, lvals, loffs,
                                                                  ev1, ev2, DegOut{lcount, ecount, deg});
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- vertex pass: pair scan of (lcount, ecount) that also writes the EulerVertex records ---------
// scans pydebruijn.py:560-567 + setupVertices :280-294 in one read of the degree slots.  The scanned
// item is a VERTEX: its four lcount and four ecount slots are one 16-byte load each, the four starts
// one 16-byte store each, so a warp row moves 512 B per array and the warp scan runs once per 128 slots.
#define VS_ROWS 4
#define VS_TILE (SCAN_THREADS * VS_ROWS)
// PAIRED: the degree slots come from the paired regions (common.cuh, DegOut): region(v) = [lcount(v) |
// ecount(partner(v))]; ecount(v) is fetched from the partner's region -- the neighbouring lane, except
// across a row boundary -- and the reference arrays lcount / ecount are written here, coalesced.
template <bool PAIRED>
__global__ void __launch_bounds__(SCAN_THREADS, 3) vertex_scan_kernel(const uint4 *__restrict__ lcount4, const uint4 *__restrict__ ecount4,
                                                                      const uint4 *__restrict__ deg4, uint4 *__restrict__ lcount_out,
                                                                      uint4 *__restrict__ ecount_out, u32 k,
                                                                      const u64 *__restrict__ vkeys, u64 nv,
                                                                      uint4 *__restrict__ lstart4, uint4 *__restrict__ estart4,
                                                                      euler_vertex *__restrict__ ev, ScanState *state, u64 *counter,
                                                                      u64 ntiles)
{
    __shared__ u64 s_tile;
    __shared__ u64 s_warp[SCAN_WARPS];
    __shared__ u64 s_prefix;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(counter, 1ull);
    __syncthreads();
    const u64 tile = s_tile;
    const u64 base = tile * (u64)VS_TILE + (u64)warp * (32 * VS_ROWS) + lane;
    uint4 lc[VS_ROWS], ec[VS_ROWS];
    u64 v[VS_ROWS];
    u64 carry = 0;
#pragma unroll
    for (int r = 0; r < VS_ROWS; r++) {
        const u64 idx = base + (u64)r * 32;
        if (!PAIRED) {
            lc[r] = idx < nv ? lcount4[idx] : make_uint4(0, 0, 0, 0);
            ec[r] = idx < nv ? ecount4[idx] : make_uint4(0, 0, 0, 0);
        } else {
            uint4 mine = make_uint4(0, 0, 0, 0);   // ecount(partner(idx)), parked in my region
            long long partner = (long long)idx;
            lc[r] = mine;
            if (idx < nv) {
                lc[r] = deg4[2 * idx];
                mine = deg4[2 * idx + 1];
                const u64 key = vkeys[idx], rk = revcomp64(key, k);
                partner += key < rk ? 1 : (key > rk ? -1 : 0);
            }
            const int src = lane + (int)(partner - (long long)idx);
            const bool in_row = src >= 0 && src < 32;
            const int s2 = in_row ? src : lane;
            ec[r].x = __shfl_sync(0xffffffffu, mine.x, s2);
            ec[r].y = __shfl_sync(0xffffffffu, mine.y, s2);
            ec[r].z = __shfl_sync(0xffffffffu, mine.z, s2);
            ec[r].w = __shfl_sync(0xffffffffu, mine.w, s2);
            if (!in_row && idx < nv) ec[r] = deg4[2 * (u64)partner + 1];
            if (idx < nv) { lcount_out[idx] = lc[r]; ecount_out[idx] = ec[r]; }
        }
        v[r] = ((u64)(ec[r].x + ec[r].y + ec[r].z + ec[r].w) << 32) | (u64)(lc[r].x + lc[r].y + lc[r].z + lc[r].w);
        carry += v[r];
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) carry += __shfl_xor_sync(0xffffffffu, carry, d);
    if (lane == 0) s_warp[warp] = carry;
    __syncthreads();
    u64 warp_off = 0, block_sum = 0;
#pragma unroll
    for (int w = 0; w < SCAN_WARPS; w++) {
        const u64 t = s_warp[w];
        if (w < warp) warp_off += t;
        block_sum += t;
    }
    if (warp == 0) {
        const u64 prefix = scan_lookback(state, tile, block_sum, lane, ntiles, (u64 *)nullptr);
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    u64 off = s_prefix + warp_off;
#pragma unroll
    for (int r = 0; r < VS_ROWS; r++) {
        const u64 idx = base + (u64)r * 32;
        u64 inc = v[r];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u64 t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        const u64 ex = off + inc - v[r];
        off += __shfl_sync(0xffffffffu, inc, 31);
        if (idx < nv) {
            const u32 l0 = (u32)ex, e0 = (u32)(ex >> 32);
            lstart4[idx] = make_uint4(l0, l0 + lc[r].x, l0 + lc[r].x + lc[r].y, l0 + lc[r].x + lc[r].y + lc[r].z);
            estart4[idx] = make_uint4(e0, e0 + ec[r].x, e0 + ec[r].x + ec[r].y, e0 + ec[r].x + ec[r].y + ec[r].z);
            euler_vertex x;
            x.vid = vkeys[idx];
            x.lp = l0; x.lcount = (u32)v[r];
            x.ep = e0; x.ecount = (u32)(v[r] >> 32);
            ev[idx] = x;
        }
    }
}

static int vertices_launch(euler_ctx *ctx, bool paired, const u32 *lcount, const u32 *ecount, const u32 *deg, u32 k, const u64 *vkeys,
                           u64 nv, u32 *lstart, u32 *estart, euler_vertex *ev)
{
    if (!nv) return EULER_OK;
    if ((((uintptr_t)lcount | (uintptr_t)ecount | (uintptr_t)lstart | (uintptr_t)estart | (uintptr_t)deg) & 15) != 0)
        return euler_fail(ctx, EULER_ERR_ARG, "degree-slot arrays must be 16-byte aligned");
    const u64 ntiles = (nv + VS_TILE - 1) / VS_TILE;
    ScanState *state = nullptr;
    u64 *counter = nullptr;
    EULER_TRY(scan_state_reserve(ctx, ntiles, &state, &counter));
    CUDA_TRY(ctx, cudaMemsetAsync(state, 0, (ntiles + 1) * sizeof(ScanState), ctx->stream));
    if (paired)
        vertex_scan_kernel<true><<<(unsigned)ntiles, SCAN_THREADS, 0, ctx->stream>>>(
            nullptr, nullptr, (const uint4 *)deg, (uint4 *)lcount, (uint4 *)ecount, k, vkeys, nv, (uint4 *)lstart, (uint4 *)estart, ev,
            state, counter, ntiles);
    else
        vertex_scan_kernel<false><<<(unsigned)ntiles, SCAN_THREADS, 0, ctx->stream>>>(
            (const uint4 *)lcount, (const uint4 *)ecount, nullptr, nullptr, nullptr, k, vkeys, nv, (uint4 *)lstart, (uint4 *)estart, ev,
            state, counter, ntiles);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

int graph_vertices_fused(euler_ctx *ctx, const u32 *lcount, const u32 *ecount, const u64 *vkeys, u64 nv, u32 *lstart,
                         u32 *estart, euler_vertex *ev)
{
    return vertices_launch(ctx, false, lcount, ecount, nullptr, 0, vkeys, nv, lstart, estart, ev);
}

// same from the paired regions `deg` (u32[8 nv]); also writes the reference arrays lcount / ecount
int graph_vertices_paired(euler_ctx *ctx, const u32 *deg, u32 k, const u64 *vkeys, u64 nv, u32 *lcount, u32 *ecount, u32 *lstart,
                          u32 *estart, euler_vertex *ev)
{
    return vertices_launch(ctx, true, lcount, ecount, deg, k, vkeys, nv, lstart, estart, ev);
}
