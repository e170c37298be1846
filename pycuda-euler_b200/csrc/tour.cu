// tour.cu -- Euler-tour stage: successor table, successor graph, connected components,
// circuit graph, spanning forest, swipe, contig starts, list-ranked contig emission.
// Replaces pyeulertour.py T1-T12, pycomponent.py C1-C10 and the host steps
// eulercuda.findSpanningTree (:267-306) / generatePartialContig (:329-404).
#include "kernels.h"
#include "scan.cuh"
#include "sort.cuh"
#include "tmp.cuh"

#define TB 256

// ---- T1 assignSuccessor pyeulertour.py:62-84 ---------------------------------------------------
// One thread per entering-list slot: slot i of vertex v (i - ep < min(ecount, lcount)) pairs
// entering edge e[i] with leaving edge l[lp + (i - ep)].
__global__ void __launch_bounds__(TB) assign_successor_kernel(const euler_vertex *__restrict__ ev, const u32 *__restrict__ lev,
                                                               const u32 *__restrict__ ent, u32 vcount,
                                                               euler_edge *__restrict__ ee, u32 ecount)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ecount) return;
    const u32 edge = ent[i];
    if (edge >= ecount) return;
    const u32 v = ee[edge].v2;
    if (v >= vcount) return;
    const euler_vertex x = ev[v];
    if (i < x.ep) return;
    const u32 r = i - x.ep;
    if (r >= x.ecount || r >= x.lcount) return;
    const u32 li = x.lp + r;
    if (li < ecount) ee[edge].s = lev[li];
}

int tour_assign_successor(euler_ctx *ctx, const euler_vertex *ev, const u32 *lev, const u32 *ent, u32 vcount,
                          euler_edge *ee, u32 ecount)
{
    if (!ecount) return EULER_OK;
    assign_successor_kernel<<<grid_for(ecount, TB), TB, 0, ctx->stream>>>(ev, lev, ent, vcount, ee, ecount);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

__global__ void __launch_bounds__(TB) reset_succ_kernel(euler_edge *__restrict__ ee, u32 ecount)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ecount) ee[t].s = ecount;
}
int tour_reset_successors(euler_ctx *ctx, euler_edge *ee, u32 ecount)
{
    if (!ecount) return EULER_OK;
    reset_succ_kernel<<<grid_for(ecount, TB), TB, 0, ctx->stream>>>(ee, ecount);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- T2/T3 successor graph pyeulertour.py:136-145,190-199 --------------------------------------
__global__ void __launch_bounds__(TB) succ_graph_p1_kernel(const euler_edge *__restrict__ ee, euler_succ_vertex *__restrict__ v,
                                                            u32 ecount)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ecount) return;
    euler_succ_vertex x;
    x.vid = (u32)ee[t].eid;
    x.n1 = ee[t].s;
    x.n2 = ecount;
    v[t] = x;
}
__global__ void __launch_bounds__(TB) succ_graph_p2_kernel(euler_succ_vertex *__restrict__ v, u32 ecount)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ecount) return;
    const u32 n1 = v[t].n1;
    if (n1 < ecount) v[n1].n2 = v[t].vid;
}

int tour_successor_graph(euler_ctx *ctx, const euler_edge *ee, u32 ecount, euler_succ_vertex *v)
{
    if (!ecount) return EULER_OK;
    succ_graph_p1_kernel<<<grid_for(ecount, TB), TB, 0, ctx->stream>>>(ee, v, ecount);
    succ_graph_p2_kernel<<<grid_for(ecount, TB), TB, 0, ctx->stream>>>(v, ecount);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- components: lock-free union-find, smaller id wins => label = min id of the component ------
// (the fix-point of pycomponent.py's atomicMin hooking :320,:330,:489,:496 + final jump :556-560)
__device__ __forceinline__ u32 ld_vol_u32(const u32 *p)
{
    u32 v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ u32 uf_find(u32 *p, u32 x)
{
    u32 cur = ld_vol_u32(p + x);
    if (cur != x) {
        u32 prev = x, next;
        while (cur > (next = ld_vol_u32(p + cur))) {
            p[prev] = next;  // path halving toward smaller ids only
            prev = cur;
            cur = next;
        }
    }
    return cur;
}
__device__ __forceinline__ void uf_union(u32 *p, u32 a, u32 b)
{
    u32 ra = uf_find(p, a), rb = uf_find(p, b);
    while (ra != rb) {
        if (ra < rb) { const u32 t = ra; ra = rb; rb = t; }  // ra > rb: hook ra under rb
        const u32 old = atomicCAS(p + ra, ra, rb);
        if (old == ra) break;
        ra = old;
    }
}
__global__ void __launch_bounds__(TB) iota_kernel(u32 *p, u32 n)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) p[t] = t;
}
__global__ void __launch_bounds__(TB) cc_hook_kernel(const euler_succ_vertex *__restrict__ v, u32 *D, u32 n)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const u32 n1 = v[t].n1, n2 = v[t].n2;
    if (n1 < n) uf_union(D, t, n1);
    if (n2 < n) uf_union(D, t, n2);
}
__global__ void __launch_bounds__(TB) cc_flatten_kernel(u32 *D, u32 n)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    u32 x = t, px;
    while ((px = ld_vol_u32(D + x)) != x) x = px;
    D[t] = x;
}

int tour_components(euler_ctx *ctx, const euler_succ_vertex *v, u32 n, u32 *D)
{
    if (!n) return EULER_OK;
    iota_kernel<<<grid_for(n, TB), TB, 0, ctx->stream>>>(D, n);
    cc_hook_kernel<<<grid_for(n, TB), TB, 0, ctx->stream>>>(v, D, n);
    cc_flatten_kernel<<<grid_for(n, TB), TB, 0, ctx->stream>>>(D, n);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- T4-T6 circuit-graph vertices pyeulertour.py:226-231,748-752,283-288 -----------------------
__global__ void __launch_bounds__(TB) mark_labels_kernel(const u32 *__restrict__ D, u32 *C, u32 n)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) C[D[t]] = 1u;
}
__global__ void __launch_bounds__(TB) gather_cv_kernel(const u32 *__restrict__ C, const u32 *__restrict__ offset, u32 n,
                                                        u32 *__restrict__ cv)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && C[t]) cv[offset[t]] = t;
}

int tour_circuit_vertices(euler_ctx *ctx, const u32 *D, u32 ecount, u32 *C, u32 *offset, u32 *cv, u64 *d_count)
{
    if (!ecount) {
        CUDA_TRY(ctx, cudaMemsetAsync(d_count, 0, sizeof(u64), ctx->stream));
        return EULER_OK;
    }
    CUDA_TRY(ctx, cudaMemsetAsync(C, 0, (size_t)ecount * 4, ctx->stream));
    mark_labels_kernel<<<grid_for(ecount, TB), TB, 0, ctx->stream>>>(D, C, ecount);
    EULER_TRY(scan_exclusive(ctx, ScanInU32{C}, ecount, offset, d_count));
    if (cv) gather_cv_kernel<<<grid_for(ecount, TB), TB, 0, ctx->stream>>>(C, offset, ecount, cv);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- T7-T9 circuit-graph edges pyeulertour.py:342-370,442-469 + host sort :792 -----------------
template <bool WRITE>
__global__ void __launch_bounds__(TB) circuit_edges_kernel(const euler_vertex *__restrict__ ev, const u32 *__restrict__ ent,
                                                            u32 vcount, const u32 *__restrict__ D, const u32 *__restrict__ cmap,
                                                            u32 ecount, u32 *__restrict__ cnt, const u32 *__restrict__ voff,
                                                            euler_circuit_edge *__restrict__ out)
{
    const u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= vcount) return;
    const euler_vertex x = ev[v];
    u32 n = 0;
    if (x.ecount > 0) {
        const u32 hi = x.ep + x.ecount - 1;
        u32 w = WRITE ? voff[v] : 0;
        for (u32 idx = x.ep; idx < hi && idx < ecount; idx++) {
            const u32 a = ent[idx], b = ent[idx + 1];
            if (a >= ecount || b >= ecount) continue;
            const u32 c1 = cmap[D[a]], c2 = cmap[D[b]];
            if (c1 == c2) continue;
            if (WRITE) {
                euler_circuit_edge ce;
                ce.ceid = 0;
                ce.c1 = c1 < c2 ? c1 : c2;
                ce.c2 = c1 < c2 ? c2 : c1;
                ce.e1 = a;
                ce.e2 = b;
                out[w++] = ce;
            }
            n++;
        }
    }
    if (!WRITE) cnt[v] = n;
}

__global__ void __launch_bounds__(TB) ce_key_e1_kernel(const euler_circuit_edge *__restrict__ ce, u64 n, u64 *__restrict__ key,
                                                        u32 *__restrict__ idx)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = ((u64)ce[i].e1 << 32) | ce[i].e2;
    idx[i] = (u32)i;
}
__global__ void __launch_bounds__(TB) ce_key_c_kernel(const euler_circuit_edge *__restrict__ ce, const u32 *__restrict__ idx,
                                                       u64 n, u64 *__restrict__ key)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const euler_circuit_edge c = ce[idx[i]];
    key[i] = ((u64)c.c1 << 32) | c.c2;
}
__global__ void __launch_bounds__(TB) ce_gather_kernel(const euler_circuit_edge *__restrict__ in, const u32 *__restrict__ idx,
                                                        u64 n, euler_circuit_edge *__restrict__ out)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[idx[i]];
}

int tour_circuit_edges(euler_ctx *ctx, const euler_vertex *ev, const euler_edge *, const u32 *ent, u32 vcount,
                       const u32 *D, const u32 *cmap, u32 ecount, euler_circuit_edge **out, u64 *count)
{
    *out = nullptr;
    *count = 0;
    if (!vcount || !ecount) return EULER_OK;
    DevTmp<u32> cnt(ctx, vcount), voff(ctx, vcount);
    DevTmp<u64> total(ctx, 1);
    TMP_CHECK(ctx, cnt); TMP_CHECK(ctx, voff); TMP_CHECK(ctx, total);
    circuit_edges_kernel<false><<<grid_for(vcount, TB), TB, 0, ctx->stream>>>(ev, ent, vcount, D, cmap, ecount, cnt, nullptr,
                                                                             nullptr);
    CUDA_TRY(ctx, cudaGetLastError());
    EULER_TRY(scan_exclusive(ctx, ScanInU32{cnt}, vcount, voff.get(), total.get()));
    u64 n = 0;
    EULER_TRY(read_u64(ctx, total, &n));
    if (!n) return EULER_OK;
    if (n >= 0xffffffffull) return euler_fail(ctx, EULER_ERR_RANGE, "circuit edge count %llu exceeds u32", n);
    DevBuf &result = ctx->cg_buf;
    EULER_TRY(dev_reserve(ctx, result, n * sizeof(euler_circuit_edge)));
    DevTmp<euler_circuit_edge> raw(ctx, n);
    DevTmp<u64> key(ctx, n), key_tmp(ctx, n);
    DevTmp<u32> idx(ctx, n), idx_tmp(ctx, n);
    const u32 nblocks = (u32)((n + RS_TILE - 1) / RS_TILE);
    DevTmp<u32> hist(ctx, (size_t)256 * nblocks);
    TMP_CHECK(ctx, raw); TMP_CHECK(ctx, key); TMP_CHECK(ctx, key_tmp); TMP_CHECK(ctx, idx); TMP_CHECK(ctx, idx_tmp);
    TMP_CHECK(ctx, hist);
    circuit_edges_kernel<true><<<grid_for(vcount, TB), TB, 0, ctx->stream>>>(ev, ent, vcount, D, cmap, ecount, nullptr, voff,
                                                                            raw);
    CUDA_TRY(ctx, cudaGetLastError());
    // np.sort(order=['c1','c2']) with ties broken by the remaining fields (ceid=0, e1, e2):
    // stable LSD: first by (e1,e2), then by (c1,c2).
    ce_key_e1_kernel<<<grid_for(n, TB), TB, 0, ctx->stream>>>(raw, n, key, idx);
    EULER_TRY(radix_sort_pairs(ctx, key, idx, n, 64, key_tmp, idx_tmp, hist));
    ce_key_c_kernel<<<grid_for(n, TB), TB, 0, ctx->stream>>>(raw, idx, n, key);
    EULER_TRY(radix_sort_pairs(ctx, key, idx, n, 64, key_tmp, idx_tmp, hist));
    ce_gather_kernel<<<grid_for(n, TB), TB, 0, ctx->stream>>>(raw, idx, n, (euler_circuit_edge *)result.p);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *out = (euler_circuit_edge *)result.p;
    *count = n;
    return EULER_OK;
}

// ---- spanning forest: Boruvka with weight = edge index (== Kruskal in index order, the unique
// minimum spanning forest under distinct weights).  Replaces eulercuda.findSpanningTree :267-306.
__global__ void __launch_bounds__(TB) fill_u32_kernel(u32 *p, u64 n, u32 val)
{
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) p[t] = val;
}
__global__ void __launch_bounds__(TB) boruvka_min_kernel(const euler_circuit_edge *__restrict__ cg, u64 m, const u32 *__restrict__ comp,
                                                          u32 *__restrict__ best, u32 *any)
{
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const u32 a = comp[cg[j].c1], b = comp[cg[j].c2];
    if (a == b) return;
    atomicMin(best + a, (u32)j);
    atomicMin(best + b, (u32)j);
    *any = 1u;
}
__global__ void __launch_bounds__(TB) boruvka_hook_kernel(const euler_circuit_edge *__restrict__ cg, const u32 *__restrict__ comp,
                                                           const u32 *__restrict__ best, u32 nv, u32 *__restrict__ next,
                                                           u32 *__restrict__ tree_flag)
{
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nv) return;
    if (comp[c] != c) return;  // not a component root
    const u32 j = best[c];
    if (j == 0xffffffffu) return;
    tree_flag[j] = 1u;
    const u32 a = comp[cg[j].c1], b = comp[cg[j].c2];
    const u32 other = (a == c) ? b : a;
    // mutual choice (both ends picked j): the smaller root stays root
    if (best[other] == j && c < other) return;
    next[c] = other;
}
__global__ void __launch_bounds__(TB) boruvka_jump_kernel(const u32 *__restrict__ next, u32 *__restrict__ comp, u32 nv)
{
    const u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    u32 x = comp[v], nx;
    while ((nx = next[x]) != x) x = nx;
    comp[v] = x;
}
struct FlagIn {
    const u32 *f;
    __device__ __forceinline__ u32 operator()(u64 i) const { return f[i] ? 1u : 0u; }
};
__global__ void __launch_bounds__(TB) compact_flag_idx_kernel(const u32 *__restrict__ flag, const u32 *__restrict__ pos, u64 n,
                                                               u32 *__restrict__ out)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) out[pos[i]] = (u32)i;
}

int tour_spanning_forest(euler_ctx *ctx, const euler_circuit_edge *cg, u64 m, u32 nv, u32 *tree, u32 *tree_count)
{
    *tree_count = 0;
    if (!m || !nv) return EULER_OK;
    DevTmp<u32> comp(ctx, nv), next(ctx, nv), best(ctx, nv), flag(ctx, m), pos(ctx, m), any(ctx, 1);
    DevTmp<u64> total(ctx, 1);
    TMP_CHECK(ctx, comp); TMP_CHECK(ctx, next); TMP_CHECK(ctx, best); TMP_CHECK(ctx, flag); TMP_CHECK(ctx, pos);
    TMP_CHECK(ctx, any); TMP_CHECK(ctx, total);
    iota_kernel<<<grid_for(nv, TB), TB, 0, ctx->stream>>>(comp, nv);
    iota_kernel<<<grid_for(nv, TB), TB, 0, ctx->stream>>>(next, nv);
    CUDA_TRY(ctx, cudaMemsetAsync(flag, 0, m * 4, ctx->stream));
    for (int round = 0; round < 64; round++) {
        fill_u32_kernel<<<grid_for(nv, TB), TB, 0, ctx->stream>>>(best, nv, 0xffffffffu);
        CUDA_TRY(ctx, cudaMemsetAsync(any, 0, 4, ctx->stream));
        boruvka_min_kernel<<<grid_for(m, TB), TB, 0, ctx->stream>>>(cg, m, comp, best, any);
        u32 h_any = 0;
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_pinned, any.get(), 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        h_any = *(u32 *)ctx->h_pinned;
        if (!h_any) break;
        boruvka_hook_kernel<<<grid_for(nv, TB), TB, 0, ctx->stream>>>(cg, comp, best, nv, next, flag);
        boruvka_jump_kernel<<<grid_for(nv, TB), TB, 0, ctx->stream>>>(next, comp, nv);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    EULER_TRY(scan_exclusive(ctx, FlagIn{flag}, m, pos.get(), total.get()));
    compact_flag_idx_kernel<<<grid_for(m, TB), TB, 0, ctx->stream>>>(flag, pos, m, tree);
    CUDA_TRY(ctx, cudaGetLastError());
    u64 n = 0;
    EULER_TRY(read_u64(ctx, total, &n));
    *tree_count = (u32)n;
    return EULER_OK;
}

// ---- T10 markSpanningEulerEdges pyeulertour.py:624-632 -----------------------------------------
__global__ void __launch_bounds__(TB) mark_spanning_kernel(const euler_circuit_edge *__restrict__ cg, const u32 *__restrict__ tree,
                                                            u32 tree_count, u32 ecount, u32 *__restrict__ mark)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= tree_count) return;
    const euler_circuit_edge c = cg[tree[t]];
    const u32 e = c.e1 < c.e2 ? c.e1 : c.e2;
    if (e < ecount) mark[e] = 1u;
}

int tour_mark_spanning(euler_ctx *ctx, const euler_circuit_edge *cg, const u32 *tree, u32 tree_count, u32 ecount, u32 *mark)
{
    CUDA_TRY(ctx, cudaMemsetAsync(mark, 0, (size_t)ecount * 4, ctx->stream));
    if (!tree_count) return EULER_OK;
    mark_spanning_kernel<<<grid_for(tree_count, TB), TB, 0, ctx->stream>>>(cg, tree, tree_count, ecount, mark);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- T11 executeSwipe pyeulertour.py:528-557 (semantics of the commented block :540-553) -------
__global__ void __launch_bounds__(TB) swipe_kernel(const euler_vertex *__restrict__ ev, const u32 *__restrict__ ent, u32 vcount,
                                                    euler_edge *__restrict__ ee, const u32 *__restrict__ mark, u32 ecount)
{
    const u32 v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= vcount) return;
    const euler_vertex x = ev[v];
    if (x.ecount == 0) return;
    u32 index = x.ep;
    const u32 maxIndex = index + x.ecount - 1;
    while (index < maxIndex && ee[ent[index]].eid < ecount) {
        if (mark[ee[ent[index]].eid] == 1u) {
            const u32 t = index;
            const u32 s = ee[ent[index]].s;
            while (index < maxIndex && mark[ee[ent[index]].eid] == 1u) {
                ee[ent[index]].s = ee[ent[index + 1]].s;
                index++;
            }
            if (t != index) ee[ent[index]].s = s;
        }
        index++;
    }
}

int tour_swipe(euler_ctx *ctx, const euler_vertex *ev, const u32 *ent, u32 vcount, euler_edge *ee, const u32 *mark,
               u32 ecount)
{
    if (!vcount || !ecount) return EULER_OK;
    swipe_kernel<<<grid_for(vcount, TB), TB, 0, ctx->stream>>>(ev, ent, vcount, ee, mark, ecount);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- T12 identifyContigStart pyeulertour.py:682-688 --------------------------------------------
__global__ void __launch_bounds__(TB) clear_start_kernel(const euler_edge *__restrict__ ee, u32 ecount, u32 *__restrict__ start)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ecount) return;
    const u32 s = ee[t].s;
    if (s < ecount) start[s] = 0u;
}

int tour_contig_starts(euler_ctx *ctx, const euler_edge *ee, u32 ecount, u32 *start)
{
    if (!ecount) return EULER_OK;
    fill_u32_kernel<<<grid_for(ecount, TB), TB, 0, ctx->stream>>>(start, ecount, 1u);
    clear_start_kernel<<<grid_for(ecount, TB), TB, 0, ctx->stream>>>(ee, ecount, start);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- contig emission: generatePartialContig host walk eulercuda.py:351-402 on device -----------
// chain node = Euler edge.  A contig is the first vertex's k-mer followed by the last base of v2
// of every edge of the chain (B12).  See chain.cuh.
#include "chain.cuh"

struct EulerChainModel {
    const euler_vertex *ev;
    const euler_edge *ee;
    u32 n;
    const u64 *vk_hi;
    static constexpr u32 HEAD_APPENDS = 1;
    __device__ __forceinline__ u32 succ(u32 i) const { return ee[i].s; }
    __device__ __forceinline__ u64 head_key(u32 i) const { return ev[ee[i].v1].vid; }
    __device__ __forceinline__ u64 head_key_hi(u32 i) const { return vk_hi ? vk_hi[ee[i].v1] : 0ull; }
    __device__ __forceinline__ char base(u32 i) const { return "ACGT"[ev[ee[i].v2].vid & 3]; }
    __device__ __forceinline__ bool emit(const u32 *, u32) const { return true; }
};

int tour_emit_contigs(euler_ctx *ctx, const euler_vertex *ev, u32 vcount, const euler_edge *ee, u32 n, u32 l,
                      char **d_out, u64 *out_bytes, u64 *ncontigs, const u64 *vk_hi)
{
    (void)vcount;
    EulerChainModel m = {ev, ee, n, vk_hi};
    return chain_emit(ctx, m, n, l - 1, d_out, out_bytes, ncontigs);
}
