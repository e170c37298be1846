/*
 * euler_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("oracle") of the pycuda-euler hot path: encode -> both-strand l-mer
 * multiset -> vertex table -> de Bruijn degree slots / scans / vertices / edges -> Euler
 * successor graph -> components -> circuit graph -> spanning forest -> swipe -> contig walk.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and there only as the checker / reported CPU baseline.  The product path
 * (pycuda-euler_b200/) never links, imports or calls it.
 *
 * Parity pinning: validated in this repo against (i) the reference's own CPU assembler
 * /root/reference/src/referenceassembler/referenceAssembler.py run unmodified on the
 * reference fixture tests/g200reads.fa (goldens under tests/golden/, generator script
 * tests/golden/make_golden.py) and (ii) the hand-derived kernel known answers of SURVEY §8c.
 * The PyCUDA kernels themselves cannot run here (no pycuda, no GPU) and are defective
 * (SURVEY §2.4); where they are the only source, this file restates their intent and says so.
 *
 * All citations are file:line under /root/reference/src/eulercuda/ unless a path is given.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* codeF, case-insensitive on the low 3 bits like pyencode.py:42,66 but with non-ACGT bytes
 * breaking the window (referenceAssembler.py:29) instead of aliasing to 'A'. */
static uint8_t orc_code[256];
static int orc_code_ready = 0;
static void orc_init_code(void)
{
    if (orc_code_ready) return;
    memset(orc_code, 4, sizeof(orc_code));
    orc_code['A'] = orc_code['a'] = 0;
    orc_code['C'] = orc_code['c'] = 1;
    orc_code['G'] = orc_code['g'] = 2;
    orc_code['T'] = orc_code['t'] = 3;
    orc_code_ready = 1;
}

#define KEY_T uint64_t
#define SFX 64
#include "oracle_impl.h"
#undef KEY_T
#undef SFX

#define KEY_T unsigned __int128
#define SFX 128
#include "oracle_impl.h"
#undef KEY_T
#undef SFX

typedef unsigned __int128 u128;

/* ---- device-ABI structs of the reference (SURVEY §2.3 "Structs") ---------------------- */
typedef struct { uint64_t vid; uint32_t ep, ecount, lp, lcount; } EulerVertex; /* pydebruijn.py:197-203 */
typedef struct { uint64_t eid; uint32_t v1, v2, s, pad; } EulerEdge;           /* pydebruijn.py:345-351 */
typedef struct { uint32_t vid, n1, n2; } SVertex;                              /* pyeulertour.py:127-131 */
typedef struct { uint32_t ceid, e1, e2, c1, c2; } CircuitEdge;                 /* pyeulertour.py:421-427 */

/* =======================================================================================
 *  1. encoder, module-level (one value per byte position of the flat buffer)
 * ===================================================================================== */

/*
 * Per-position encodings: out_fwd[t] = l-mer starting at byte t, out_rc[t] = its reverse
 * complement, valid[t] = 1 iff the window lies inside one read and is all-ACGT; 0/0/0 elsewhere.
 * encodeLmerDevice pyencode.py:45-76, encodeLmerComplementDevice :173-209 (intent).
 * len <= 32.
 */
int orc_encode_positions(const char *buf, const uint64_t *off, uint64_t nreads, uint32_t len,
                         uint64_t *out_fwd, uint64_t *out_rc, uint8_t *valid)
{
    orc_init_code();
    if (len < 1 || len > 32) return -1;
    const uint64_t B = off[nreads];
    memset(out_fwd, 0, B * sizeof(uint64_t));
    memset(out_rc, 0, B * sizeof(uint64_t));
    memset(valid, 0, B);
    const uint64_t mask = key_mask64(len);
    const uint32_t top = 2u * (len - 1u);
    for (uint64_t j = 0; j < nreads; j++) {
        uint64_t f = 0, r = 0, run = 0;
        for (uint64_t t = off[j]; t < off[j + 1]; t++) {
            const uint8_t c = orc_code[(uint8_t)buf[t]];
            if (c > 3) { run = 0; f = r = 0; continue; }
            f = ((f << 2) | c) & mask;
            r = (r >> 2) | ((uint64_t)(3u - c) << top);
            if (++run >= len) {
                const uint64_t s = t + 1 - len;
                out_fwd[s] = f; out_rc[s] = r; valid[s] = 1;
            }
        }
    }
    return 0;
}

/* computeKmerDevice pyencode.py:112-135: prefix/suffix (l-1)-mers; mask = 2(l-1) ones (eulercuda.py:112-113) */
void orc_compute_kmers(const uint64_t *lmers, uint64_t n, uint64_t kmer_mask, uint64_t *pk, uint64_t *sk)
{
    for (uint64_t i = 0; i < n; i++) {
        pk[i] = (lmers[i] & (kmer_mask << 2)) >> 2;
        sk[i] = lmers[i] & kmer_mask;
    }
}

/* hash_h pygpuhash.py:28-36 (compat restatement only; the product table is open addressing) */
uint32_t orc_hash_h(uint64_t key, uint32_t bucket_count)
{
    return (uint32_t)(((0x01010101ull + 0x12345678ull * key) % 1900813ull) % bucket_count);
}

uint64_t orc_revcomp64(uint64_t x, uint32_t len) { return key_rc64(x, len); }

/* =======================================================================================
 *  2. mer counting (any length <= 64)
 * ===================================================================================== */
typedef struct {
    uint64_t n;
    uint64_t *lo, *hi;  /* hi == NULL-equivalent zeros when len <= 32 */
    uint32_t *vals;
} orc_counts;

orc_counts *orc_count(const char *buf, const uint64_t *off, uint64_t nreads, uint32_t len)
{
    orc_init_code();
    if (len < 1 || len > 64) return NULL;
    orc_counts *c = (orc_counts *)calloc(1, sizeof(*c));
    if (len <= 32) {
        kv_t64 kv;
        if (count_mers64(buf, off, nreads, len, &kv)) { free(c); return NULL; }
        c->n = kv.n; c->lo = kv.keys; c->vals = kv.vals;
        c->hi = (uint64_t *)calloc(kv.n ? kv.n : 1, sizeof(uint64_t));
    } else {
        kv_t128 kv;
        if (count_mers128(buf, off, nreads, len, &kv)) { free(c); return NULL; }
        c->n = kv.n; c->vals = kv.vals;
        c->lo = (uint64_t *)malloc((kv.n ? kv.n : 1) * sizeof(uint64_t));
        c->hi = (uint64_t *)malloc((kv.n ? kv.n : 1) * sizeof(uint64_t));
        for (uint64_t i = 0; i < kv.n; i++) { c->lo[i] = (uint64_t)kv.keys[i]; c->hi[i] = (uint64_t)(kv.keys[i] >> 64); }
        free(kv.keys);
    }
    return c;
}
uint64_t orc_counts_n(const orc_counts *c) { return c->n; }
void orc_counts_copy(const orc_counts *c, uint64_t *lo, uint64_t *hi, uint32_t *vals)
{
    if (lo) memcpy(lo, c->lo, c->n * sizeof(uint64_t));
    if (hi) memcpy(hi, c->hi, c->n * sizeof(uint64_t));
    if (vals) memcpy(vals, c->vals, c->n * sizeof(uint32_t));
}
void orc_counts_free(orc_counts *c) { if (!c) return; free(c->lo); free(c->hi); free(c->vals); free(c); }

/* =======================================================================================
 *  3. de Bruijn graph build (D1-D6), l <= 32 (64-bit keys) and l <= 64 (128-bit keys)
 * ===================================================================================== */
typedef struct {
    uint32_t l;
    uint64_t nl, nv, ne;           /* distinct l-mers, vertices, edges (= sum of multiplicities, B6) */
    uint64_t *lk_lo, *lk_hi;       /* lmerKeys (ascending) */
    uint32_t *lvals;               /* lmerValues */
    uint64_t *loffs;               /* lmerOffsets = exscan(lmerValues)  pydebruijn.py:569-573 */
    uint64_t *vk_lo, *vk_hi;       /* kmerKeys (ascending), id = index */
    uint32_t *lcount, *ecount;     /* 4*nv each, pydebruijn.py:134,140 */
    uint64_t *lstart, *estart;     /* exscan, pydebruijn.py:560-567 */
    uint32_t *ev1, *ev2;           /* compressed edge endpoints per distinct l-mer */
    EulerVertex *ev;               /* nv (u32 fields; only valid when ne < 2^32) */
    EulerEdge *ee;                 /* ne, only when expanded */
    uint32_t *le, *ee_in;          /* l[] and e[] lists, only when expanded */
    int expanded;
} orc_graph;

#define GRAPH_BODY(KEY_T, SFX)                                                                     \
    {                                                                                              \
        CAT(kv_t, SFX) kv;                                                                         \
        if (CAT(count_mers, SFX)(buf, off, nreads, l, &kv)) { free(g); return NULL; }              \
        KEY_T *vk; uint64_t nv;                                                                    \
        if (CAT(vertex_set, SFX)(kv.keys, kv.n, l, &vk, &nv)) { free(g); return NULL; }            \
        const KEY_T kmask = CAT(key_mask, SFX)(l - 1);                                             \
        g->nl = kv.n; g->nv = nv; g->lvals = kv.vals;                                              \
        g->lk_lo = (uint64_t *)malloc((kv.n + 1) * 8); g->lk_hi = (uint64_t *)calloc(kv.n + 1, 8); \
        g->vk_lo = (uint64_t *)malloc((nv + 1) * 8); g->vk_hi = (uint64_t *)calloc(nv + 1, 8);     \
        g->loffs = (uint64_t *)malloc((kv.n + 1) * 8);                                             \
        g->lcount = (uint32_t *)calloc(4 * nv + 4, 4); g->ecount = (uint32_t *)calloc(4 * nv + 4, 4); \
        g->lstart = (uint64_t *)calloc(4 * nv + 4, 8); g->estart = (uint64_t *)calloc(4 * nv + 4, 8); \
        g->ev1 = (uint32_t *)malloc((kv.n + 1) * 4); g->ev2 = (uint32_t *)malloc((kv.n + 1) * 4);  \
        for (uint64_t i = 0; i < nv; i++) {                                                        \
            g->vk_lo[i] = (uint64_t)vk[i];                                                         \
            if (sizeof(KEY_T) > 8) g->vk_hi[i] = (uint64_t)((u128)vk[i] >> 64);                    \
        }                                                                                          \
        uint64_t acc = 0;                                                                          \
        /* debruijnCount pydebruijn.py:107-141 */                                                  \
        for (uint64_t i = 0; i < kv.n; i++) {                                                      \
            const KEY_T x = kv.keys[i];                                                            \
            g->lk_lo[i] = (uint64_t)x;                                                             \
            if (sizeof(KEY_T) > 8) g->lk_hi[i] = (uint64_t)((u128)x >> 64);                        \
            g->loffs[i] = acc; acc += kv.vals[i];                                                  \
            const uint32_t p = CAT(find_key, SFX)(vk, nv, (x >> 2) & kmask);                       \
            const uint32_t s = CAT(find_key, SFX)(vk, nv, x & kmask);                              \
            const uint32_t to = (uint32_t)(x & 3), from = (uint32_t)((x >> (2 * (l - 1))) & 3);    \
            g->ev1[i] = p; g->ev2[i] = s;                                                          \
            g->lcount[4 * (uint64_t)p + to] = kv.vals[i];                                          \
            g->ecount[4 * (uint64_t)s + from] = kv.vals[i];                                        \
        }                                                                                          \
        g->ne = acc;                                                                               \
        free(kv.keys); free(vk);                                                                   \
    }

orc_graph *orc_graph_build(const char *buf, const uint64_t *off, uint64_t nreads, uint32_t l, int expand)
{
    orc_init_code();
    if (l < 2 || l > 64) return NULL;
    orc_graph *g = (orc_graph *)calloc(1, sizeof(*g));
    g->l = l;
    if (l <= 32) GRAPH_BODY(uint64_t, 64) else GRAPH_BODY(u128, 128)

    const uint64_t nv = g->nv, nl = g->nl;
    /* scans pydebruijn.py:560-567 */
    uint64_t a = 0, b = 0;
    for (uint64_t i = 0; i < 4 * nv; i++) {
        g->lstart[i] = a; a += g->lcount[i];
        g->estart[i] = b; b += g->ecount[i];
    }
    /* setupVertices pydebruijn.py:280-294 */
    g->ev = (EulerVertex *)calloc(nv + 1, sizeof(EulerVertex));
    for (uint64_t v = 0; v < nv; v++) {
        EulerVertex *x = &g->ev[v];
        x->vid = g->vk_lo[v];
        x->lp = (uint32_t)g->lstart[4 * v];
        x->ep = (uint32_t)g->estart[4 * v];
        x->lcount = g->lcount[4 * v] + g->lcount[4 * v + 1] + g->lcount[4 * v + 2] + g->lcount[4 * v + 3];
        x->ecount = g->ecount[4 * v] + g->ecount[4 * v + 1] + g->ecount[4 * v + 2] + g->ecount[4 * v + 3];
    }
    g->expanded = 0;
    if (expand && g->ne < 0xffffffffull) {
        /* setupEdges pydebruijn.py:426-475 (B7 guard dropped; s = E "no successor", :467-468) */
        const uint64_t ne = g->ne;
        g->ee = (EulerEdge *)calloc(ne + 1, sizeof(EulerEdge));
        g->le = (uint32_t *)calloc(ne + 1, 4);
        g->ee_in = (uint32_t *)calloc(ne + 1, 4);
        for (uint64_t i = 0; i < nl; i++) {
            const uint32_t to = (uint32_t)(g->lk_lo[i] & 3);
            uint32_t from;
            if (2 * (l - 1) >= 64) from = (uint32_t)((g->lk_hi[i] >> (2 * (l - 1) - 64)) & 3);
            else from = (uint32_t)((g->lk_lo[i] >> (2 * (l - 1))) & 3);
            uint64_t lo = g->lstart[4 * (uint64_t)g->ev1[i] + to];
            uint64_t eo = g->estart[4 * (uint64_t)g->ev2[i] + from];
            uint64_t id = g->loffs[i];
            for (uint32_t m = 0; m < g->lvals[i]; m++, lo++, eo++, id++) {
                g->ee[id].eid = id; g->ee[id].v1 = g->ev1[i]; g->ee[id].v2 = g->ev2[i];
                g->ee[id].s = (uint32_t)ne; g->ee[id].pad = 0;
                g->le[lo] = (uint32_t)id;
                g->ee_in[eo] = (uint32_t)id;
            }
        }
        g->expanded = 1;
    }
    return g;
}

void orc_graph_counts(const orc_graph *g, uint64_t *nl, uint64_t *nv, uint64_t *ne, int *expanded)
{
    *nl = g->nl; *nv = g->nv; *ne = g->ne; *expanded = g->expanded;
}

/* which: 0 lk_lo 1 lk_hi 2 lvals 3 loffs 4 vk_lo 5 vk_hi 6 lcount 7 ecount 8 lstart 9 estart
 *        10 ev1 11 ev2 12 ev 13 ee 14 l 15 e */
int orc_graph_copy(const orc_graph *g, int which, void *dst)
{
    const uint64_t nl = g->nl, nv = g->nv, ne = g->ne;
    switch (which) {
    case 0: memcpy(dst, g->lk_lo, nl * 8); break;
    case 1: memcpy(dst, g->lk_hi, nl * 8); break;
    case 2: memcpy(dst, g->lvals, nl * 4); break;
    case 3: memcpy(dst, g->loffs, nl * 8); break;
    case 4: memcpy(dst, g->vk_lo, nv * 8); break;
    case 5: memcpy(dst, g->vk_hi, nv * 8); break;
    case 6: memcpy(dst, g->lcount, 4 * nv * 4); break;
    case 7: memcpy(dst, g->ecount, 4 * nv * 4); break;
    case 8: memcpy(dst, g->lstart, 4 * nv * 8); break;
    case 9: memcpy(dst, g->estart, 4 * nv * 8); break;
    case 10: memcpy(dst, g->ev1, nl * 4); break;
    case 11: memcpy(dst, g->ev2, nl * 4); break;
    case 12: memcpy(dst, g->ev, nv * sizeof(EulerVertex)); break;
    case 13: if (!g->expanded) return -1; memcpy(dst, g->ee, ne * sizeof(EulerEdge)); break;
    case 14: if (!g->expanded) return -1; memcpy(dst, g->le, ne * 4); break;
    case 15: if (!g->expanded) return -1; memcpy(dst, g->ee_in, ne * 4); break;
    default: return -1;
    }
    return 0;
}

void orc_graph_free(orc_graph *g)
{
    if (!g) return;
    free(g->lk_lo); free(g->lk_hi); free(g->lvals); free(g->loffs); free(g->vk_lo); free(g->vk_hi);
    free(g->lcount); free(g->ecount); free(g->lstart); free(g->estart); free(g->ev1); free(g->ev2);
    free(g->ev); free(g->ee); free(g->le); free(g->ee_in); free(g);
}

/* =======================================================================================
 *  4. Euler tour stage on plain arrays (key-width independent)
 * ===================================================================================== */

/* assignSuccessor pyeulertour.py:62-84: pair i-th entering with i-th leaving edge */
void orc_assign_successor(const EulerVertex *ev, const uint32_t *l, const uint32_t *e, uint32_t vcount,
                          EulerEdge *ee, uint32_t ecount)
{
    for (uint32_t v = 0; v < vcount; v++) {
        const uint32_t n = ev[v].ecount < ev[v].lcount ? ev[v].ecount : ev[v].lcount;
        for (uint32_t i = 0; i < n; i++) {
            const uint32_t ei = ev[v].ep + i, li = ev[v].lp + i;
            if (ei < ecount && li < ecount && e[ei] < ecount) ee[e[ei]].s = l[li];
        }
    }
}

/* constructSuccessorGraphP1/P2 pyeulertour.py:136-145,190-199 */
void orc_successor_graph(const EulerEdge *ee, SVertex *v, uint32_t ecount)
{
    for (uint32_t t = 0; t < ecount; t++) { v[t].vid = (uint32_t)ee[t].eid; v[t].n1 = ee[t].s; v[t].n2 = ecount; }
    for (uint32_t t = 0; t < ecount; t++) if (v[t].n1 < ecount) v[v[t].n1].n2 = v[t].vid;
}

static uint32_t uf_find(uint32_t *p, uint32_t x)
{
    while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; }
    return x;
}

/*
 * find_component_device pycomponent.py:676-732 taken to its fix-point (B8): every hook is an
 * atomicMin toward the smaller label (:320,:330,:489,:496) and the final jump (:556-560) flattens,
 * so at convergence D[i] = minimum node id of i's component.
 */
void orc_components(const SVertex *v, uint32_t *D, uint32_t n)
{
    for (uint32_t i = 0; i < n; i++) D[i] = i;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t nb[2] = { v[i].n1, v[i].n2 };
        for (int k = 0; k < 2; k++) {
            if (nb[k] >= n) continue;
            uint32_t a = uf_find(D, i), b = uf_find(D, nb[k]);
            if (a == b) continue;
            if (a < b) D[b] = a; else D[a] = b;
        }
    }
    for (uint32_t i = 0; i < n; i++) D[i] = uf_find(D, i);
}

/* calculateCircuitGraphVertexData :226-231 + scan :748-752 + constructCircuitGraphVertex :283-288.
 * C[c]=1 for every label c in use; offset = exscan(C); cv[offset[t]] = t. Returns circuit count. */
uint32_t orc_circuit_vertices(const uint32_t *D, uint32_t ecount, uint32_t *C, uint32_t *offset, uint32_t *cv)
{
    memset(C, 0, (size_t)ecount * 4);
    for (uint32_t t = 0; t < ecount; t++) C[D[t]] = 1;
    uint32_t acc = 0;
    for (uint32_t t = 0; t < ecount; t++) { offset[t] = acc; acc += C[t]; }
    if (cv) for (uint32_t t = 0; t < ecount; t++) if (C[t]) cv[offset[t]] = t;
    return acc;
}

static int cmp_cedge(const void *pa, const void *pb)
{
    const CircuitEdge *a = (const CircuitEdge *)pa, *b = (const CircuitEdge *)pb;
    if (a->c1 != b->c1) return a->c1 < b->c1 ? -1 : 1;
    if (a->c2 != b->c2) return a->c2 < b->c2 ? -1 : 1;
    if (a->ceid != b->ceid) return a->ceid < b->ceid ? -1 : 1;
    if (a->e1 != b->e1) return a->e1 < b->e1 ? -1 : 1;
    if (a->e2 != b->e2) return a->e2 < b->e2 ? -1 : 1;
    return 0;
}

/*
 * calculateCircuitGraphEdgeData :342-370 + scan :774-777 + assignCircuitGraphEdgeData :442-469 +
 * host sort :792 (np.sort order=['c1','c2'], remaining fields break ties in dtype order).
 * Returns the number of circuit edges; `out` must hold `cap` entries (returns needed count if > cap).
 */
uint64_t orc_circuit_edges(const EulerVertex *ev, const uint32_t *e, uint32_t vcount, const uint32_t *D,
                           const uint32_t *map, uint32_t ecount, CircuitEdge *out, uint64_t cap)
{
    uint64_t n = 0;
    for (uint32_t v = 0; v < vcount; v++) {
        if (ev[v].ecount == 0) continue;
        const uint32_t hi = ev[v].ep + ev[v].ecount - 1;
        for (uint32_t idx = ev[v].ep; idx < hi && idx < ecount; idx++) {
            if (e[idx] >= ecount || e[idx + 1] >= ecount) continue;
            const uint32_t c1 = map[D[e[idx]]], c2 = map[D[e[idx + 1]]];
            if (c1 == c2) continue;
            if (n < cap) {
                out[n].ceid = 0;
                out[n].c1 = c1 < c2 ? c1 : c2; out[n].c2 = c1 < c2 ? c2 : c1;
                out[n].e1 = e[idx]; out[n].e2 = e[idx + 1];
            }
            n++;
        }
    }
    if (n <= cap) qsort(out, n, sizeof(CircuitEdge), cmp_cedge);
    return n;
}

/*
 * findSpanningTree eulercuda.py:267-306 (graph_tool Kruskal, unit weights) with B10 fixed:
 * returns the list of circuit-edge indices of the spanning forest, Kruskal in index order
 * (== the unique minimum spanning forest under weight = edge index).
 */
uint32_t orc_spanning_forest(const CircuitEdge *cg, uint64_t cg_count, uint32_t cg_vcount, uint32_t *tree)
{
    uint32_t *p = (uint32_t *)malloc(((size_t)cg_vcount + 1) * 4);
    for (uint32_t i = 0; i < cg_vcount; i++) p[i] = i;
    uint32_t nt = 0;
    for (uint64_t j = 0; j < cg_count; j++) {
        uint32_t a = uf_find(p, cg[j].c1), b = uf_find(p, cg[j].c2);
        if (a == b) continue;
        if (a < b) p[b] = a; else p[a] = b;
        tree[nt++] = (uint32_t)j;
    }
    free(p);
    return nt;
}

/* markSpanningEulerEdges pyeulertour.py:624-632 (mark starts all-zero; the port's np.ones at :661 is a defect) */
void orc_mark_spanning(const CircuitEdge *cg, const uint32_t *tree, uint32_t tree_count, uint32_t *mark)
{
    for (uint32_t t = 0; t < tree_count; t++) {
        const CircuitEdge *c = &cg[tree[t]];
        mark[c->e1 < c->e2 ? c->e1 : c->e2] = 1;
    }
}

/* executeSwipe pyeulertour.py:528-557, semantics from the commented-out body :540-553 (B9) */
void orc_swipe(const EulerVertex *ev, const uint32_t *e, uint32_t vcount, EulerEdge *ee, const uint32_t *mark,
               uint32_t ecount)
{
    for (uint32_t v = 0; v < vcount; v++) {
        if (ev[v].ecount == 0) continue;
        uint32_t index = ev[v].ep;
        const uint32_t maxIndex = index + ev[v].ecount - 1;
        while (index < maxIndex && ee[e[index]].eid < ecount) {
            if (mark[ee[e[index]].eid] == 1) {
                const uint32_t t = index;
                const uint32_t s = ee[e[index]].s;
                while (index < maxIndex && mark[ee[e[index]].eid] == 1) {
                    ee[e[index]].s = ee[e[index + 1]].s;
                    index++;
                }
                if (t != index) ee[e[index]].s = s;
            }
            index++;
        }
    }
}

/* identifyContigStart pyeulertour.py:682-688: start[i]=1 unless i is somebody's successor */
void orc_contig_starts(const EulerEdge *ee, uint32_t ecount, uint32_t *start)
{
    for (uint32_t t = 0; t < ecount; t++) start[t] = 1;
    for (uint32_t t = 0; t < ecount; t++) if (ee[t].s < ecount) start[ee[t].s] = 0;
}

static void put_kmer(char *dst, uint64_t lo, uint64_t hi, uint32_t k)
{
    /* getString eulercuda.py:315-321: MSB-first, A0 C1 G2 T3 */
    u128 x = ((u128)hi << 64) | lo;
    for (uint32_t i = 0; i < k; i++) { dst[k - 1 - i] = "ACGT"[(uint32_t)(x & 3)]; x >>= 2; }
}

/*
 * generatePartialContig host walk eulercuda.py:351-402 with B12 fixed: a contig is the first
 * vertex's k-mer followed by the last base of each following vertex (the C original advanced by
 * l-2, see the residue at :365,:370).  Contigs are written '\n'-separated into `out` (cap bytes);
 * returns bytes needed; *ncontigs receives the count.  vk_hi may be NULL (k <= 32).
 */
uint64_t orc_walk_contigs(const uint64_t *vk_lo, const uint64_t *vk_hi, const EulerEdge *ee, uint32_t ecount,
                          uint32_t l, char *out, uint64_t cap, uint64_t *ncontigs)
{
    const uint32_t k = l - 1;
    uint32_t *start = (uint32_t *)malloc(((size_t)ecount + 1) * 4);
    uint8_t *visited = (uint8_t *)calloc((size_t)ecount + 1, 1);
    orc_contig_starts(ee, ecount, start);
    uint64_t w = 0, nc = 0;
    char kbuf[64];
#define EMIT_FULL(vidx) do { put_kmer(kbuf, vk_lo[vidx], vk_hi ? vk_hi[vidx] : 0, k); \
        for (uint32_t q_ = 0; q_ < k; q_++) { if (w < cap) out[w] = kbuf[q_]; w++; } } while (0)
#define EMIT_LAST(vidx) do { if (w < cap) out[w] = "ACGT"[(uint32_t)(vk_lo[vidx] & 3)]; w++; } while (0)
    for (int pass = 0; pass < 2; pass++) {
        for (uint32_t i = 0; i < ecount; i++) {
            if (visited[i]) continue;
            if (pass == 0 && !start[i]) continue;
            EMIT_FULL(ee[i].v1);
            uint32_t next = i;
            while (ee[next].s < ecount && !visited[ee[next].s]) {
                visited[next] = 1;
                next = ee[next].s;
                EMIT_LAST(ee[next].v1);
            }
            if (!visited[next]) { EMIT_LAST(ee[next].v2); visited[next] = 1; }
            if (w < cap) out[w] = '\n';
            w++; nc++;
        }
    }
#undef EMIT_FULL
#undef EMIT_LAST
    free(start); free(visited);
    *ncontigs = nc;
    return w;
}

/* =======================================================================================
 *  5. synthetic reads, identical to the device generator
 * ===================================================================================== */
static inline uint64_t splitmix64_at(uint64_t seed, uint64_t ctr)
{
    uint64_t z = seed + (ctr + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint32_t genome_base(uint64_t i)
{
    return (uint32_t)((splitmix64_at(0x5EED0001ull, i >> 5) >> (2 * (i & 31))) & 3);
}
/* reads [first, first+count) of the (G, L, err_ppm) data set, L bytes each, no separators */
void orc_synth_reads(uint64_t G, uint32_t L, uint32_t err_ppm, uint64_t first, uint64_t count, char *out)
{
    const uint64_t thr = ((uint64_t)err_ppm << 32) / 1000000ull;
    #pragma omp parallel for schedule(static)
    for (int64_t jj = 0; jj < (int64_t)count; jj++) {
        const uint64_t j = first + (uint64_t)jj;
        const uint64_t h = splitmix64_at(0x5EED0002ull, j);
        const uint64_t st = (h >> 1) % (G - L + 1);
        const int rcs = (int)(h & 1);
        char *o = out + (uint64_t)jj * L;
        for (uint32_t p = 0; p < L; p++) {
            uint32_t b = rcs ? 3u - genome_base(st + L - 1 - p) : genome_base(st + p);
            if (thr) {
                const uint64_t e = splitmix64_at(0x5EED0003ull, j * (uint64_t)L + p);
                if ((e & 0xffffffffull) < thr) b = (b + 1 + (uint32_t)((e >> 32) % 3)) & 3;
            }
            o[p] = "ACGT"[b];
        }
    }
}
