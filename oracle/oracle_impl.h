/*
 * oracle_impl.h -- TEST INFRASTRUCTURE ONLY (CPU oracle). Not part of the product path.
 *
 * Width-generic body of the CPU restatement.  Included twice by euler_oracle.c with
 *   KEY_T  = uint64_t            SFX = 64   (l-mer length <= 32)
 *   KEY_T  = unsigned __int128   SFX = 128  (l-mer length <= 64)
 *
 * Each function cites the reference file:line (under /root/reference/src/eulercuda/ unless
 * noted) whose *intended* semantics it restates, with the defect fixes of SURVEY.md §2.4.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

/* ---- 2-bit packing: pyencode.py:42 (codeF), :64-71 (MSB-first pack) ------------------- */

static inline KEY_T FN(key_mask)(uint32_t len)
{
    const uint32_t bits = 2u * len;
    if (bits >= 8u * sizeof(KEY_T)) return ~(KEY_T)0;
    return (((KEY_T)1) << bits) - 1;
}

/* reverse complement of a len-mer: referenceAssembler.py:7-10 (twin); pyencode.py:169 (codeR) */
static inline KEY_T FN(key_rc)(KEY_T x, uint32_t len)
{
    KEY_T r = 0;
    for (uint32_t i = 0; i < len; i++) {
        r = (r << 2) | (KEY_T)(3u - (uint32_t)(x & 3u));
        x >>= 2;
    }
    return r;
}

typedef struct {
    KEY_T   *keys;   /* sorted ascending, distinct */
    uint32_t *vals;  /* multiplicity */
    uint64_t n;
} FN(kv_t);

static int FN(cmp_key)(const void *a, const void *b)
{
    const KEY_T x = *(const KEY_T *)a, y = *(const KEY_T *)b;
    return (x > y) - (x < y);
}

/* LSD radix sort (8-bit digits) of n keys; falls back to qsort for tiny n. */
static void FN(sort_keys)(KEY_T *a, uint64_t n, uint32_t len)
{
    if (n < 2048) { qsort(a, n, sizeof(KEY_T), FN(cmp_key)); return; }
    KEY_T *tmp = (KEY_T *)malloc(n * sizeof(KEY_T));
    const uint32_t nbytes = (2u * len + 7u) / 8u;
    KEY_T *src = a, *dst = tmp;
    for (uint32_t d = 0; d < nbytes; d++) {
        uint64_t hist[257];
        memset(hist, 0, sizeof(hist));
        const uint32_t sh = 8u * d;
        for (uint64_t i = 0; i < n; i++) hist[1 + (uint32_t)((src[i] >> sh) & 0xff)]++;
        for (int i = 0; i < 256; i++) hist[i + 1] += hist[i];
        for (uint64_t i = 0; i < n; i++) dst[hist[(uint32_t)((src[i] >> sh) & 0xff)]++] = src[i];
        KEY_T *t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, n * sizeof(KEY_T));
    free(tmp);
}

/*
 * Number of valid windows of length `len` in one read: windows are per read (SURVEY B1 fix,
 * eulercuda.py:91,138) and never contain a non-ACGT byte (referenceAssembler.py:29 split('N')).
 */
static uint64_t FN(count_windows)(const char *s, uint64_t n, uint32_t len)
{
    uint64_t run = 0, cnt = 0;
    for (uint64_t i = 0; i < n; i++) {
        if (orc_code[(uint8_t)s[i]] > 3) run = 0; else run++;
        if (run >= len) cnt++;
    }
    return cnt;
}

/*
 * Emit forward and reverse-complement encodings of every valid window of one read.
 * Restates encodeLmerDevice (pyencode.py:45-76) + the *intended* encodeLmerComplementDevice
 * (pyencode.py:173-209, defect B2 -> twin()).
 * out receives 2 keys per window (fwd, rc) when both_strands, else 1.
 */
static uint64_t FN(emit_windows)(const char *s, uint64_t n, uint32_t len, int both_strands, KEY_T *out)
{
    const KEY_T mask = FN(key_mask)(len);
    const uint32_t top = 2u * (len - 1u);
    KEY_T f = 0, r = 0;
    uint64_t run = 0, cnt = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t c = orc_code[(uint8_t)s[i]];
        if (c > 3) { run = 0; f = 0; r = 0; continue; }
        f = ((f << 2) | c) & mask;
        r = (r >> 2) | ((KEY_T)(3u - c) << top);
        run++;
        if (run >= len) {
            out[cnt++] = f;
            if (both_strands) out[cnt++] = r;
        }
    }
    return cnt;
}

/*
 * Both-strand multiset of len-mers over all reads, as sorted distinct keys + counts.
 * Restates the host dict fill readLmersKmersCuda (eulercuda.py:141-178) with B1/B3 fixed,
 * == referenceAssembler.build(reads, len, limit=0) (referenceAssembler.py:25-42).
 */
static int FN(count_mers)(const char *buf, const uint64_t *off, uint64_t nreads, uint32_t len,
                          FN(kv_t) *out)
{
    uint64_t *woff = (uint64_t *)malloc((nreads + 1) * sizeof(uint64_t));
    if (!woff) return -1;
    #pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < (int64_t)nreads; j++)
        woff[j + 1] = 2 * FN(count_windows)(buf + off[j], off[j + 1] - off[j], len);
    woff[0] = 0;
    for (uint64_t j = 0; j < nreads; j++) woff[j + 1] += woff[j];
    const uint64_t total = woff[nreads];
    KEY_T *all = (KEY_T *)malloc((total ? total : 1) * sizeof(KEY_T));
    if (!all) { free(woff); return -1; }
    #pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < (int64_t)nreads; j++)
        FN(emit_windows)(buf + off[j], off[j + 1] - off[j], len, 1, all + woff[j]);
    free(woff);

    /* parallel sort: split by the top byte of the (2*len)-bit key, sort buckets independently */
    const uint32_t bits = 2u * len;
    const uint32_t sh = bits > 8 ? bits - 8 : 0;
    uint64_t bstart[257];
    memset(bstart, 0, sizeof(bstart));
    for (uint64_t i = 0; i < total; i++) bstart[1 + (uint32_t)((all[i] >> sh) & 0xff)]++;
    for (int i = 0; i < 256; i++) bstart[i + 1] += bstart[i];
    KEY_T *srt = (KEY_T *)malloc((total ? total : 1) * sizeof(KEY_T));
    if (!srt) { free(all); return -1; }
    {
        uint64_t cur[256];
        memcpy(cur, bstart, sizeof(cur));
        for (uint64_t i = 0; i < total; i++) srt[cur[(uint32_t)((all[i] >> sh) & 0xff)]++] = all[i];
    }
    free(all);
    #pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < 256; b++)
        FN(sort_keys)(srt + bstart[b], bstart[b + 1] - bstart[b], len);

    /* run-length encode */
    uint64_t nd = 0;
    for (uint64_t i = 0; i < total; i++) if (i == 0 || srt[i] != srt[i - 1]) nd++;
    out->keys = (KEY_T *)malloc((nd ? nd : 1) * sizeof(KEY_T));
    out->vals = (uint32_t *)malloc((nd ? nd : 1) * sizeof(uint32_t));
    out->n = nd;
    uint64_t w = 0;
    for (uint64_t i = 0; i < total; i++) {
        if (i == 0 || srt[i] != srt[i - 1]) { out->keys[w] = srt[i]; out->vals[w] = 1; w++; }
        else out->vals[w - 1]++;
    }
    free(srt);
    return 0;
}

/* binary search in sorted distinct keys; returns index or UINT32_MAX.
 * Stands in for getHashValue (pydebruijn.py:57-88): key -> vertex id, 0xffffffff on miss. */
static inline uint32_t FN(find_key)(const KEY_T *keys, uint64_t n, KEY_T key)
{
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = lo + (hi - lo) / 2;
        if (keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    return (lo < n && keys[lo] == key) ? (uint32_t)lo : 0xffffffffu;
}

/*
 * Vertex set: distinct prefix/suffix (l-1)-mers of the distinct l-mers, ascending; id = rank
 * (eulercuda.py:143-146,166-170 kmerMap, with B14: sorted-key ids as GPU-Euler's std::map).
 * prefix/suffix: pyencode.py:109-110.
 */
static int FN(vertex_set)(const KEY_T *lkeys, uint64_t nl, uint32_t l, KEY_T **vkeys_out, uint64_t *nv_out)
{
    const KEY_T kmask = FN(key_mask)(l - 1);
    KEY_T *tmp = (KEY_T *)malloc((2 * nl + 1) * sizeof(KEY_T));
    if (!tmp) return -1;
    for (uint64_t i = 0; i < nl; i++) {
        tmp[2 * i] = (lkeys[i] >> 2) & kmask;
        tmp[2 * i + 1] = lkeys[i] & kmask;
    }
    FN(sort_keys)(tmp, 2 * nl, l - 1);
    uint64_t nv = 0;
    for (uint64_t i = 0; i < 2 * nl; i++) if (i == 0 || tmp[i] != tmp[i - 1]) tmp[nv++] = tmp[i];
    *vkeys_out = tmp;
    *nv_out = nv;
    return 0;
}
