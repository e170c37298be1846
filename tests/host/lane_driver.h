// lane_driver.h -- host helpers shared by the SIMT-emulator checks (test infrastructure only): random reads, the
// read-start bitmap, a brute-force l-mer census and the EXPECTED records of the partition pass, produced by driving the
// lane logic of pycuda-euler_b200/csrc/bucket.cuh over warp tiles on the host exactly as bucket_lane_check.cpp does
// (that file verifies the lane logic itself against the delivery rule).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "bucket.cuh"

typedef unsigned long long u64;
typedef unsigned int u32;

static u64 ld_rng_state = 0x9E3779B97F4A7C15ull;
static inline u64 rnd()
{
    u64 z = (ld_rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline int code_of(char ch)
{
    switch (ch & 0xDF) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    }
    return -1;
}

struct Reads {
    std::string buf;             // all reads back to back
    std::vector<u64> off;        // read offsets, off.back() == buf.size()
    std::vector<char> is_start;  // per base
};

// genome > 0: reads are windows of one random genome, either strand, an N now and then (coverage: repeated l-mers);
// genome == 0: independent random reads with N's, lowercase and low-complexity stretches (palindromes, repeats);
// genome < 0: homopolymers and short tandem repeats only
static inline Reads make_reads(int nreads, int maxlen, int genome_len)
{
    Reads R;
    std::string genome;
    for (int i = 0; i < genome_len; i++) genome.push_back("ACGT"[rnd() & 3]);
    for (int r = 0; r < nreads; r++) {
        R.off.push_back(R.buf.size());
        int len = (int)(rnd() % (u64)(maxlen + 1));
        if (genome_len < 0) {   // low complexity: homopolymers, short tandem repeats (palindromic l-mers, large multiplicities)
            static const char *units[] = {"A", "C", "AT", "ACGT", "AAC", "GGGTTT"};
            const std::string u = units[rnd() % 6];
            len = maxlen / 2 + (int)(rnd() % (u64)(maxlen / 2 + 1));
            for (int i = 0; i < len; i++) R.buf.push_back(u[i % u.size()]);
        } else if (genome_len > 0) {
            if (len > genome_len) len = genome_len;
            const int s = (int)(rnd() % (u64)(genome_len - len + 1));
            std::string rd = genome.substr(s, len);
            if (rnd() & 1) {
                std::reverse(rd.begin(), rd.end());
                for (auto &ch : rd) ch = ch == 'A' ? 'T' : ch == 'C' ? 'G' : ch == 'G' ? 'C' : 'A';
            }
            if ((rnd() % 16) == 0 && len > 0) rd[rnd() % (u64)len] = 'N';
            R.buf += rd;
        } else {
            for (int i = 0; i < len; i++) {
                char ch = "ACGT"[rnd() & 3];
                if ((rnd() % 512) == 0) ch = 'N';
                if ((rnd() & 31) == 0) ch = (char)(ch | 0x20);
                if (i >= 3 && (rnd() & 7) == 0) ch = R.buf[R.buf.size() - 3];
                R.buf.push_back(ch);
            }
        }
    }
    R.off.push_back(R.buf.size());
    R.is_start.assign(R.buf.size() + 1, 0);
    for (size_t r = 0; r + 1 < R.off.size(); r++)
        if (R.off[r] < R.buf.size()) R.is_start[R.off[r]] = 1;
    return R;
}

// bit (pos & 31) of word pos >> 5 = a read starts at base pos (what mark_starts_kernel writes, encode.cu)
static inline std::vector<u32> start_bitmap(const Reads &R)
{
    std::vector<u32> bits(R.buf.size() / 32 + 2, 0u);
    for (size_t r = 0; r + 1 < R.off.size(); r++)
        if (R.off[r] < R.buf.size()) bits[R.off[r] >> 5] |= 1u << (R.off[r] & 31);
    return bits;
}

struct Census {
    std::map<u64, u64> M;   // strand l-mer -> both-strand multiplicity
    std::set<u64> VS;       // vertices: prefix and suffix k-mers of the strand l-mers
    u64 N_l = 0, N_k = 0;   // valid forward l-mer / k-mer windows
};
static inline Census census(const Reads &R, u32 l)
{
    Census c;
    const u32 k = l - 1;
    std::map<u64, u64> occ;
    auto window = [&](u64 p, u64 e, u32 len, u64 &x) {   // the len-mer starting at p, inside the read, all ACGT
        if (p + len > e) return false;
        x = 0;
        for (u32 j = 0; j < len; j++) {
            const int cd = code_of(R.buf[p + j]);
            if (cd < 0) return false;
            x = (x << 2) | (u64)cd;
        }
        return true;
    };
    for (size_t r = 0; r + 1 < R.off.size(); r++)
        for (u64 p = R.off[r]; p < R.off[r + 1]; p++) {
            u64 x;
            if (window(p, R.off[r + 1], k, x)) c.N_k++;
            if (window(p, R.off[r + 1], l, x)) { occ[x]++; c.N_l++; }
        }
    for (auto &kv : occ) {
        c.M[kv.first] += kv.second;
        c.M[bk_revcomp(kv.first, l)] += kv.second;
    }
    const u64 kmask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
    for (auto &kv : c.M) { c.VS.insert(kv.first >> 2); c.VS.insert(kv.first & kmask); }
    return c;
}

struct ChunkH { u32 codes, vmask, smask; };
static inline ChunkH load_chunk_h(const Reads &R, long long chunk)
{
    ChunkH c = {0, 0, 0};
    if (chunk < 0) return c;
    for (int i = 0; i < 16; i++) {
        const u64 pos = (u64)chunk * 16 + i;
        if (pos >= R.buf.size()) break;
        const int cd = code_of(R.buf[pos]);
        if (cd >= 0) { c.codes |= (u32)cd << (30 - 2 * i); c.vmask |= 1u << (15 - i); }
        if (R.is_start[pos]) c.smask |= 1u << (15 - i);
    }
    return c;
}

// the records the partition pass must deliver, per GLOBAL bucket (rank * nb_per_rank + local bucket)
static inline std::vector<std::vector<BkRec>> host_records(const Reads &R, u32 l, BkGeom g)
{
    const u32 k = l - 1, m = bk_m_of(k), W = k - m + 1;
    std::vector<std::vector<BkRec>> recs((size_t)g.nranks * g.nb_per_rank);
    const int HALO = 2, ADV = 30;
    const u64 B = R.buf.size(), nchunks = (B + 15) / 16, ntiles = (nchunks + ADV - 1) / ADV;
    for (u64 tile = 0; tile < ntiles; tile++) {
        ChunkH ch[32];
        u32 sc[32][16], win[32][16];
        for (int lane = 0; lane < 32; lane++) ch[lane] = load_chunk_h(R, (long long)(tile * ADV) - HALO + lane);
        for (int lane = 0; lane < 32; lane++) bk_chunk_scores(lane >= 1 ? ch[lane - 1].codes : 0u, ch[lane].codes, m, sc[lane]);
        for (int lane = 0; lane < 32; lane++) {
            u32 sa[36];
            for (int t = 0; t < 4; t++) sa[t] = lane >= 2 ? sc[lane - 2][12 + t] : 0xdeadbeefu + t;
            for (int t = 0; t < 16; t++) sa[4 + t] = lane >= 1 ? sc[lane - 1][t] : 0xfeedf00du + t;
            for (int t = 0; t < 16; t++) sa[20 + t] = sc[lane][t];
            bk_window_min_any(sa, W, win[lane]);
        }
        for (int lane = HALO; lane < 32; lane++) {
            const u64 vmw = ((u64)ch[lane - 2].vmask << 48) | ((u64)ch[lane - 1].vmask << 32) | ((u64)ch[lane].vmask << 16);
            const u64 smw = ((u64)ch[lane - 2].smask << 48) | ((u64)ch[lane - 1].smask << 32) | ((u64)ch[lane].smask << 16);
            const u64 VK = bk_valid_kmers(vmw, smw, k);
            const u32 vk16 = bk_own16(VK), vl16 = bk_own16(bk_valid_lmers(VK, smw, k));
            const u32 win_prev = win[lane - 1][15];
            bk_lane_pieces(ch[lane - 2].codes, ch[lane - 1].codes, ch[lane].codes, win[lane], win_prev, bk_eq16(win[lane], win_prev), vk16,
                           vl16, k, g, [&](u32 bucket, const BkRec &r) { recs[bucket].push_back(r); });
        }
    }
    return recs;
}

template <typename T>
static inline T *aligned_array(size_t n, int fill)
{
    void *p = nullptr;
    if (posix_memalign(&p, 64, (n + 8) * sizeof(T))) abort();
    memset(p, fill, (n + 8) * sizeof(T));
    return (T *)p;
}
