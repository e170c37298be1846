// bucket_lane_check.cpp -- CPU check of the partition pass's per-lane logic (pycuda-euler_b200/csrc/bucket.cuh):
// the same __host__ __device__ functions the CUDA kernel calls (m-mer scores, sliding minimum, window validity,
// piece cutting, record packing) are driven lane by lane over warp tiles of random reads, and the records they
// emit are decoded back and compared with a brute-force statement of the rule:
//   every valid forward l-mer window is delivered, once, to the bucket of its prefix k-mer and to the bucket of
//   its suffix k-mer (once when they coincide), together with "is that end vertex owned by this bucket".
// Build: g++ -O2 -std=c++17 -I<csrc> bucket_lane_check.cpp -o bucket_lane_check ; exit code 0 = all cases agree.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "bucket.cuh"

typedef unsigned long long u64;
typedef unsigned int u32;

static u64 rng_state = 0x9E3779B97F4A7C15ull;
static u64 rnd()
{
    u64 z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static int code_of(char ch)
{
    switch (ch & 0xDF) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    }
    return -1;
}

struct ChunkH { u32 codes, vmask, smask; };
// host restatement of encode.cuh load_chunk (same bit layout)
static ChunkH load_chunk_h(const std::string &buf, const std::vector<char> &is_start, long long chunk)
{
    ChunkH c = {0, 0, 0};
    if (chunk < 0) return c;
    for (int i = 0; i < 16; i++) {
        const u64 pos = (u64)chunk * 16 + i;
        if (pos >= buf.size()) break;
        const int cd = code_of(buf[pos]);
        if (cd >= 0) { c.codes |= (u32)cd << (30 - 2 * i); c.vmask |= 1u << (15 - i); }
        if (is_start[pos]) c.smask |= 1u << (15 - i);
    }
    return c;
}

typedef std::tuple<u32, u64, u32> Item;   // bucket, canonical l-mer, own bits (bit 0 prefix(c), bit 1 suffix(c))

static u32 brute_bucket(u64 kmer, u32 k, BkGeom g)
{
    const u32 m = bk_m_of(k);
    const u64 r = bk_revcomp(kmer, k);
    const u32 mmask = m >= 16 ? 0xffffffffu : ((1u << (2 * m)) - 1u);
    u32 best = 0xffffffffu;
    for (u32 j = 0; j + m <= k; j++) {
        const u32 w = (u32)(kmer >> (2 * (k - m - j))) & mmask, rw = (u32)(r >> (2 * j)) & mmask;
        const u32 sc = bk_mmer_score(w < rw ? w : rw);
        best = sc < best ? sc : best;
    }
    return bk_bucket_of(best, g);
}

template <int W>
static void win_dispatch(const u32 (&sa)[36], u32 Wrt, u32 (&win)[16])
{
    if (W > 0) bk_window_min<(W > 0 ? W : 1)>(sa, win);
    else bk_window_min_any(sa, Wrt, win);
}

static int run_case(u32 l, u32 nranks, u32 nbpr, int nreads, int maxlen, double pn, bool use_template)
{
    const u32 k = l - 1, m = bk_m_of(k), W = k - m + 1;
    const BkGeom g = {nranks, nbpr};
    // reads: random length, random bases, some N / lowercase
    std::string buf;
    std::vector<u64> off;
    for (int r = 0; r < nreads; r++) {
        off.push_back(buf.size());
        const int len = (int)(rnd() % (u64)(maxlen + 1));
        // low-complexity stretches now and then so that minimizers repeat
        for (int i = 0; i < len; i++) {
            char ch = "ACGT"[rnd() & 3];
            if ((rnd() % 1000) < (u64)(pn * 1000)) ch = 'N';
            if ((rnd() & 31) == 0) ch = (char)(ch | 0x20);
            if (i >= 3 && (rnd() & 15) == 0) ch = buf[buf.size() - 3];
            buf.push_back(ch);
        }
    }
    off.push_back(buf.size());
    const u64 B = buf.size();
    std::vector<char> is_start(B + 1, 0);
    for (size_t r = 0; r + 1 < off.size(); r++)
        if (off[r] < B) is_start[off[r]] = 1;

    // ---- expected deliveries, brute force per read
    std::vector<Item> expect;
    u64 exp_nl = 0, exp_nk = 0;
    for (size_t r = 0; r + 1 < off.size(); r++) {
        const u64 a = off[r], e = off[r + 1];
        for (u64 p = a; p < e; p++) {
            // k-mer window ending at p
            bool okk = p + 1 >= a + k, okl = p + 1 >= a + l;
            u64 km = 0, lm = 0;
            for (u32 j = 0; j < l && okl; j++) { const int cd = code_of(buf[p - (l - 1) + j]); if (cd < 0) okl = false; else lm = (lm << 2) | (u64)cd; }
            for (u32 j = 0; j < k && okk; j++) { const int cd = code_of(buf[p - (k - 1) + j]); if (cd < 0) okk = false; else km = (km << 2) | (u64)cd; }
            (void)km;
            if (okk) exp_nk++;
            if (!okl) continue;
            exp_nl++;
            const u64 kmask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
            const u64 pre = lm >> 2, suf = lm & kmask;
            const u64 rp = bk_revcomp(pre, k), rs = bk_revcomp(suf, k);
            const u32 bp = brute_bucket(pre < rp ? pre : rp, k, g), bs = brute_bucket(suf < rs ? suf : rs, k, g);
            const u64 rc = bk_revcomp(lm, l);
            const bool flip = rc < lm;
            const u64 c = flip ? rc : lm;
            for (int side = 0; side < 2; side++) {
                const u32 X = side ? bp : bs;
                if (side && bp == bs) break;
                const u32 opf = bp == X, osf = bs == X;
                expect.push_back(Item(X, c, flip ? (osf | (opf << 1)) : (opf | (osf << 1))));
            }
        }
    }

    // ---- the lane logic over warp tiles, as the kernel drives it
    std::vector<Item> got;
    u64 got_nl = 0, got_nk = 0;
    const int HALO = 2, ADV = 30;
    const u64 nchunks = (B + 15) / 16;
    const u64 ntiles = (nchunks + ADV - 1) / ADV;
    int bad_records = 0;
    for (u64 tile = 0; tile < ntiles; tile++) {
        ChunkH ch[32];
        u32 sc[32][16], win[32][16];
        for (int lane = 0; lane < 32; lane++) ch[lane] = load_chunk_h(buf, is_start, (long long)(tile * ADV) - HALO + lane);
        for (int lane = 0; lane < 32; lane++) bk_chunk_scores(lane >= 1 ? ch[lane - 1].codes : 0u, ch[lane].codes, m, sc[lane]);
        for (int lane = 0; lane < 32; lane++) {
            u32 sa[36];
            for (int t = 0; t < 4; t++) sa[t] = lane >= 2 ? sc[lane - 2][12 + t] : 0xdeadbeefu + t;
            for (int t = 0; t < 16; t++) sa[4 + t] = lane >= 1 ? sc[lane - 1][t] : 0xfeedf00du + t;
            for (int t = 0; t < 16; t++) sa[20 + t] = sc[lane][t];
            if (use_template && W == 20) bk_window_min<20>(sa, win[lane]);
            else if (use_template && W == 10) bk_window_min<10>(sa, win[lane]);
            else if (use_template && W == 1) bk_window_min<1>(sa, win[lane]);
            else if (use_template && W == 7) bk_window_min<7>(sa, win[lane]);
            else if (use_template && W == 16) bk_window_min<16>(sa, win[lane]);
            else bk_window_min_any(sa, W, win[lane]);
        }
        for (int lane = HALO; lane < 32; lane++) {
            const u32 v0 = ch[lane].vmask, v1 = ch[lane - 1].vmask, v2 = ch[lane - 2].vmask;
            const u32 s0 = ch[lane].smask, s1 = ch[lane - 1].smask, s2 = ch[lane - 2].smask;
            const u64 vmw = ((u64)v2 << 48) | ((u64)v1 << 32) | ((u64)v0 << 16);
            const u64 smw = ((u64)s2 << 48) | ((u64)s1 << 32) | ((u64)s0 << 16);
            const u64 VK = bk_valid_kmers(vmw, smw, k);
            const u32 vk16 = bk_own16(VK), vl16 = bk_own16(bk_valid_lmers(VK, smw, k));
            got_nk += bk_popc(vk16);
            got_nl += bk_popc(vl16);
            const u32 win_prev = win[lane - 1][15];
            bk_lane_pieces(ch[lane - 2].codes, ch[lane - 1].codes, ch[lane].codes, win[lane], win_prev, bk_eq16(win[lane], win_prev),
                           vk16, vl16, k, g, [&](u32 bucket, const BkRec &r) {
                               const u32 n = r.hdr & 63u;
                               if (n <= k || n > BK_MAX_BASES) { bad_records++; return; }
                               const u32 nl = n - k;
                               if (nl > 17) { bad_records++; return; }
                               for (u32 j = 0; j < nl; j++) {
                                   const u64 x = bk_record_lmer(r, j, l);
                                   const u64 rc = bk_revcomp(x, l);
                                   const bool flip = rc < x;
                                   const u32 opf = (j == 0 && (r.hdr & BK_HDR_LFF)) ? 0u : 1u;
                                   const u32 osf = (j + 1 == nl && (r.hdr & BK_HDR_RFF)) ? 0u : 1u;
                                   got.push_back(Item(bucket, flip ? rc : x, flip ? (osf | (opf << 1)) : (opf | (osf << 1))));
                               }
                           });
        }
    }
    std::sort(expect.begin(), expect.end());
    std::sort(got.begin(), got.end());
    const bool ok = expect == got && exp_nl == got_nl && exp_nk == got_nk && bad_records == 0;
    if (!ok) {
        fprintf(stderr, "MISMATCH l=%u nranks=%u nbpr=%u reads=%d maxlen=%d: expect %zu items (Nl %llu Nk %llu), got %zu (Nl %llu Nk %llu), bad records %d\n",
                l, nranks, nbpr, nreads, maxlen, expect.size(), exp_nl, exp_nk, got.size(), got_nl, got_nk, bad_records);
        size_t i = 0;
        while (i < expect.size() && i < got.size() && expect[i] == got[i]) i++;
        if (i < expect.size()) fprintf(stderr, "  first expected-only at %zu: bucket %u key %llx own %u\n", i, std::get<0>(expect[i]), std::get<1>(expect[i]), std::get<2>(expect[i]));
        if (i < got.size()) fprintf(stderr, "  first got-only at %zu: bucket %u key %llx own %u\n", i, std::get<0>(got[i]), std::get<1>(got[i]), std::get<2>(got[i]));
    }
    return ok ? 0 : 1;
}

int main()
{
    int fails = 0, cases = 0;
    const u32 ls[] = {2, 3, 5, 10, 13, 14, 18, 22, 23, 27, 28, 31, 32};
    for (u32 l : ls)
        for (int rep = 0; rep < 3; rep++) {
            const u32 nranks = rep == 0 ? 1 : (rep == 1 ? 3 : 8);
            const u32 nbpr = rep == 0 ? 7 : (rep == 1 ? 1 : 5);
            fails += run_case(l, nranks, nbpr, 300, 120, rep == 2 ? 0.02 : 0.0, true);
            fails += run_case(l, nranks, nbpr, 40, 700, 0.003, rep != 1);
            cases += 2;
        }
    // one bucket: everything is owned everywhere
    fails += run_case(32, 1, 1, 200, 150, 0.01, true);
    fails += run_case(22, 1, 1, 200, 150, 0.01, true);
    // many buckets
    fails += run_case(32, 1, 4096, 400, 150, 0.0, true);
    fails += run_case(32, 2, 2048, 400, 150, 0.0, true);
    cases += 4;
    printf("%d cases, %d failed\n", cases, fails);
    return fails ? 1 : 0;
}
