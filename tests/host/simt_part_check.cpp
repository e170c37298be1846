// simt_part_check.cpp -- the partition pass (pycuda-euler_b200/csrc/bucket_part.cu: bkt_partition_kernel in its direct
// and its multi-GPU stream form, bkt_push_counts_kernel, bkt_regroup_kernel) and, behind it, the per-bucket build
// (bucket_build.cu) run on the CPU under the SIMT emulator of simt_emu.h: `nranks` ranks emulated one after the other,
// every rank's scatter writing into every owner's stream area as the peer stores do on the device.  TEST
// INFRASTRUCTURE ONLY -- the same kernel sources, compiled with -DEULER_SIMT_EMU.
//
// Checked: the records that arrive in every bucket region are, as a multiset, the records the lane logic prescribes
// (lane_driver.h, itself verified by bucket_lane_check.cpp); window counts; the run reservation of the stream form
// (reserved slots that no record took are empty records, the owner skips them; counts are capped at the stream
// capacity and an overflow raises BKT_FLAG_REGION); and that the union of the per-owner graphs is the graph of all
// reads: every strand l-mer on exactly one owner with its multiplicity, every vertex on exactly one owner with its
// eight degree slots, the suffix vertex of an edge either local and right or 0xffffffff when another rank owns it.
//
// Build: g++ -O1 -std=c++17 -pthread -I<csrc> -I<tests/host> simt_part_check.cpp -o simt_part_check
#define EULER_SIMT_EMU
#include "../../pycuda-euler_b200/csrc/bucket_part.cu"
#include "../../pycuda-euler_b200/csrc/bucket_build.cu"

#include "lane_driver.h"

typedef std::vector<std::vector<uint4>> Regions;

static bool rec_less(const uint4 &a, const uint4 &b)
{
    if (a.x != b.x) return a.x < b.x;
    if (a.y != b.y) return a.y < b.y;
    if (a.z != b.z) return a.z < b.z;
    return a.w < b.w;
}
static bool same_multiset(std::vector<uint4> a, std::vector<uint4> b)
{
    if (a.size() != b.size()) return false;
    std::sort(a.begin(), a.end(), rec_less);
    std::sort(b.begin(), b.end(), rec_less);
    for (size_t i = 0; i < a.size(); i++)
        if (a[i].x != b[i].x || a[i].y != b[i].y || a[i].z != b[i].z || a[i].w != b[i].w) return false;
    return true;
}

struct Shard {
    Reads R;
    uint4 *buf16;
    std::vector<u32> bits;
    u64 ntiles;
};

template <bool STREAM>
static void run_partition(const Shard &s, u32 l, BkGeom g, u32 my_rank, u32 rcap, uint4 *const *dst, u32 *cursors, u64 *stats, unsigned grid)
{
    const u32 k = l - 1, W = k - bk_m_of(k) + 1;
    const u64 B = s.R.buf.size();
    if (!B) return;
    if (W == 20) simt::launch(grid, BP_BLOCK, 0, [&] { bkt_partition_kernel<20, STREAM>(s.buf16, B, s.bits.data(), l, g, my_rank, rcap, dst, cursors, s.ntiles, stats); });
    else if (W == 10) simt::launch(grid, BP_BLOCK, 0, [&] { bkt_partition_kernel<10, STREAM>(s.buf16, B, s.bits.data(), l, g, my_rank, rcap, dst, cursors, s.ntiles, stats); });
    else simt::launch(grid, BP_BLOCK, 0, [&] { bkt_partition_kernel<0, STREAM>(s.buf16, B, s.bits.data(), l, g, my_rank, rcap, dst, cursors, s.ntiles, stats); });
}

// the per-bucket build of one owner (what bkt_build() launches), artefacts returned in host vectors
struct Graph {
    std::vector<u64> lkeys, vkeys;
    std::vector<u32> lvals, ev1, ev2, lcount, ecount;
    u64 U = 0, V = 0, E = 0, flags = 0;
};
static Graph run_build(const Regions &regions, u32 l, u32 cap, u64 ucap, u64 vcap)
{
    const u32 nb = (u32)regions.size();
    u32 rcap = 1;
    for (auto &v : regions) rcap = std::max<u32>(rcap, (u32)v.size());
    uint4 *records = aligned_array<uint4>((size_t)nb * rcap, 0);
    u32 *counts = aligned_array<u32>(nb, 0);
    for (u32 b = 0; b < nb; b++) {
        counts[b] = (u32)regions[b].size();
        for (size_t i = 0; i < regions[b].size(); i++) records[(size_t)b * rcap + i] = regions[b][i];
    }
    u64 *lkeys = aligned_array<u64>(ucap + 8, 0);
    u32 *lvals = aligned_array<u32>(ucap + 8, 0), *loffs = aligned_array<u32>(ucap + 8, 0), *ev1 = aligned_array<u32>(ucap + 8, 0),
        *ev2 = aligned_array<u32>(ucap + 8, 0);
    u64 *vkeys = aligned_array<u64>(vcap + 8, 0);
    u32 *lcount = aligned_array<u32>(4 * vcap + 32, 0), *ecount = aligned_array<u32>(4 * vcap + 32, 0), *lstart = aligned_array<u32>(4 * vcap + 32, 0),
        *estart = aligned_array<u32>(4 * vcap + 32, 0);
    euler_vertex *ev = aligned_array<euler_vertex>(vcap + 8, 0);
    u32 *state32 = aligned_array<u32>((size_t)nb + 3 + BKT_REDO_CAP + 8, 0);
    u64 *state64 = aligned_array<u64>(4ull * nb, 0x5C);
    u64 bcap = 64;
    while (bcap < ucap) bcap <<= 1;
    u64 *bkeys = aligned_array<u64>(bcap, 0xFF);
    u32 *bvals = aligned_array<u32>(2 * bcap, 0xFF);
    u64 *stats = aligned_array<u64>(64, 0);
    BkBuildArgs a;
    a.records = records; a.counts = counts; a.nb = nb; a.nranks = 1; a.rcap = rcap; a.l = l; a.cap = cap;
    a.lkeys = lkeys; a.lvals = lvals; a.loffs = loffs; a.ev1 = ev1; a.ev2 = ev2; a.ucap = ucap;
    a.vkeys = vkeys; a.lcount = lcount; a.ecount = ecount; a.lstart = lstart; a.estart = estart; a.ev = ev; a.vcap = vcap;
    a.flag = state32; a.ticket = state32 + nb; a.redo = a.ticket + 2; a.second = 0;
    a.agg_uv = state64; a.agg_e = state64 + nb; a.inc_uv = state64 + 2ull * nb; a.inc_e = state64 + 3ull * nb;
    a.bkeys = bkeys; a.bvals = bvals; a.bcap = bcap; a.stats = stats;
    auto run = [&](unsigned grid, size_t smem) {
        if (l == 32) simt::launch(grid, BB_THREADS, smem, [&] { bkt_build_kernel<32>(a); });
        else if (l == 22) simt::launch(grid, BB_THREADS, smem, [&] { bkt_build_kernel<22>(a); });
        else simt::launch(grid, BB_THREADS, smem, [&] { bkt_build_kernel<0>(a); });
    };
    run(nb, (size_t)29 * cap);
    const u32 nredo = a.redo[0];
    a.second = 1; a.cap = BKT_MAX_CAP;
    run((nredo < BKT_REDO_CAP ? nredo : BKT_REDO_CAP) + 1, (size_t)29 * BKT_MAX_CAP);
    simt::launch(2, 256, 0, [&] { bkt_boundary_publish_kernel(lkeys, ev1, ev2, stats + 3, ucap, l, bkeys, bvals, bcap, stats); });
    simt::launch(2, 256, 0, [&] { bkt_fixup_kernel(lkeys, ev2, stats + 3, ucap, l, bkeys, bvals, bcap); });
    Graph G;
    G.flags = stats[2]; G.U = stats[3]; G.V = stats[4]; G.E = stats[5];
    if (!G.flags && G.U <= ucap && G.V <= vcap) {
        G.lkeys.assign(lkeys, lkeys + G.U); G.lvals.assign(lvals, lvals + G.U);
        G.ev1.assign(ev1, ev1 + G.U); G.ev2.assign(ev2, ev2 + G.U);
        G.vkeys.assign(vkeys, vkeys + G.V);
        G.lcount.assign(lcount, lcount + 4 * G.V); G.ecount.assign(ecount, ecount + 4 * G.V);
    }
    free(records); free(counts); free(lkeys); free(lvals); free(loffs); free(ev1); free(ev2); free(vkeys); free(lcount); free(ecount);
    free(lstart); free(estart); free(ev); free(state32); free(state64); free(bkeys); free(bvals); free(stats);
    return G;
}

struct Case {
    u32 l, nranks, nbpr, cap;
    int nreads, maxlen, genome;
    double scap_factor;   // stream capacity = fullest stream * factor (< 1: BKT_FLAG_REGION expected, nothing else checked)
};

static int run_case(const Case &cs)
{
    const u32 l = cs.l, k = l - 1, nranks = cs.nranks, nbpr = cs.nbpr;
    const BkGeom g = {nranks, nbpr};
    const u64 kmask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
    int bad = 0;
    auto fail = [&](const char *what, u64 x = 0, u64 y = 0) {
        if (bad++ < 8) fprintf(stderr, "  case l=%u nranks=%u nbpr=%u: %s (%llu vs %llu)\n", l, nranks, nbpr, what, x, y);
    };
    // ---- reads of all ranks, and each rank's shard (uneven: rank r takes reads r, r + nranks, ... of the first 3/4, rank 0 the rest)
    const Reads all = make_reads(cs.nreads, cs.maxlen, cs.genome);
    std::vector<Shard> shards(nranks);
    for (size_t r = 0; r + 1 < all.off.size(); r++) {
        const u32 owner = r < (all.off.size() - 1) * 3 / 4 ? (u32)(r % nranks) : 0u;
        Reads &S = shards[owner].R;
        S.off.push_back(S.buf.size());
        S.buf += all.buf.substr(all.off[r], all.off[r + 1] - all.off[r]);
    }
    u64 N_l = 0, N_k = 0;
    std::vector<std::vector<BkRec>> expect((size_t)nranks * nbpr);
    for (auto &s : shards) {
        s.R.off.push_back(s.R.buf.size());
        s.R.is_start.assign(s.R.buf.size() + 1, 0);
        for (size_t r = 0; r + 1 < s.R.off.size(); r++)
            if (s.R.off[r] < s.R.buf.size()) s.R.is_start[s.R.off[r]] = 1;
        s.bits = start_bitmap(s.R);
        s.buf16 = aligned_array<uint4>(s.R.buf.size() / 16 + 2, 0);
        memcpy(s.buf16, s.R.buf.data(), s.R.buf.size());
        s.ntiles = ((s.R.buf.size() + 15) / 16 + ENC_ADV - 1) / ENC_ADV;
        const Census c = census(s.R, l);
        N_l += c.N_l; N_k += c.N_k;
        auto rec = host_records(s.R, l, g);
        for (size_t b = 0; b < rec.size(); b++) expect[b].insert(expect[b].end(), rec[b].begin(), rec[b].end());
    }
    auto expected_region = [&](u32 bucket) {
        std::vector<uint4> v;
        for (auto &r : expect[bucket]) v.push_back(make_uint4(r.hdr, r.d[0], r.d[1], r.d[2]));
        return v;
    };
    const Census C = census(all, l);

    // ---- direct form (one GPU): records straight into the bucket regions
    if (nranks == 1) {
        u32 rcap = 1;
        for (auto &v : expect) rcap = std::max<u32>(rcap, (u32)v.size());
        uint4 *regions = aligned_array<uint4>((size_t)nbpr * rcap, 0x5C);
        u32 *cursors = aligned_array<u32>(nbpr, 0);
        u64 *stats = aligned_array<u64>(64, 0);
        uint4 *dst[16] = {regions};
        run_partition<false>(shards[0], l, g, 0, rcap, dst, cursors, stats, 3);
        if (stats[0] != N_l || stats[1] != N_k) fail("window counts of the direct form", stats[0], N_l);
        if (stats[2]) fail("flags of the direct form", stats[2]);
        for (u32 b = 0; b < nbpr && !bad; b++) {
            std::vector<uint4> got(regions + (size_t)b * rcap, regions + (size_t)b * rcap + std::min(cursors[b], rcap));
            if (!same_multiset(got, expected_region(b))) fail("records of a bucket region (direct form)", b, got.size());
        }
        // a region capacity that is too small: the flag, the true demand in the cursors, nothing written behind a region's end
        if (rcap > 2) {
            const u32 half = rcap / 2;
            uint4 *small = aligned_array<uint4>((size_t)nbpr * half + 64, 0x5C);
            uint4 *dst2[16] = {small};
            memset(cursors, 0, nbpr * sizeof(u32));
            memset(stats, 0, 64 * sizeof(u64));
            run_partition<false>(shards[0], l, g, 0, half, dst2, cursors, stats, 2);
            if (!(stats[2] & BKT_FLAG_REGION)) fail("REGION flag of the direct form");
            for (u32 b = 0; b < nbpr; b++)
                if (cursors[b] != expect[b].size()) fail("cursor of an overflowing region is not the demand", cursors[b], expect[b].size());
            for (int i = 0; i < 64; i++)
                if (small[(size_t)nbpr * half + i].x != 0x5C5C5C5Cu) { fail("write behind the regions"); break; }
            for (u32 b = 0; b < nbpr && !bad; b++) {   // what fits is a sub-multiset of what was due
                std::vector<uint4> got(small + (size_t)b * half, small + (size_t)b * half + std::min(cursors[b], half)), due = expected_region(b);
                std::sort(got.begin(), got.end(), rec_less);
                std::sort(due.begin(), due.end(), rec_less);
                if (!std::includes(due.begin(), due.end(), got.begin(), got.end(), rec_less)) fail("records of a truncated region", b);
            }
            free(small);
        }
        free(regions); free(cursors); free(stats);
    }

    // ---- stream form: every rank scatters into every owner's area, the owner regroups
    std::vector<u64> to_rank(nranks, 0);   // fullest stream, to size scap
    u64 fullest = 1;
    for (u32 s = 0; s < nranks; s++) {
        auto rec = host_records(shards[s].R, l, g);
        for (u32 d = 0; d < nranks; d++) {
            u64 n = 0;
            for (u32 lb = 0; lb < nbpr; lb++) n += rec[(size_t)d * nbpr + lb].size();
            fullest = std::max(fullest, n);
        }
    }
    // a tile reserves a slot for every piece and for the orphan a chunk-leading piece MAY add: streams hold a few more
    // (empty) records than the buckets receive -- 2x is ample for the exact-capacity cases
    const u32 scap = cs.scap_factor >= 1.0 ? (u32)(fullest * 2 * cs.scap_factor) + 64 : (u32)(fullest * cs.scap_factor) + 1;
    const size_t stream_bytes = (size_t)nranks * scap * 16;
    std::vector<uint4 *> area(nranks);
    for (u32 d = 0; d < nranks; d++) area[d] = (uint4 *)aligned_array<unsigned char>(stream_bytes + nranks * 8 + 64, 0x5C);
    u64 flags_scatter = 0, nl_sum = 0, nk_sum = 0, max_stream = 0;
    for (u32 s = 0; s < nranks; s++) {
        u32 *cursors = aligned_array<u32>(16, 0);
        u64 *stats = aligned_array<u64>(64, 0);
        run_partition<true>(shards[s], l, g, s, scap, area.data(), cursors, stats, 2);
        simt::launch(1, 32, 0, [&] { bkt_push_counts_kernel(cursors, area.data(), stream_bytes, nranks, s, scap, stats + 6); });
        flags_scatter |= stats[2]; nl_sum += stats[0]; nk_sum += stats[1]; max_stream = std::max<u64>(max_stream, stats[6]);
        free(cursors); free(stats);
    }
    if (cs.scap_factor < 1.0) {
        if (!(flags_scatter & BKT_FLAG_REGION)) fail("REGION flag of an overflowing stream");
        if (max_stream <= scap) fail("fullest stream not reported", max_stream, scap);
        for (u32 d = 0; d < nranks; d++) {
            const u64 *counts = (const u64 *)((const char *)area[d] + stream_bytes);
            for (u32 s = 0; s < nranks; s++)
                if (counts[s] > scap) fail("a pushed count above the stream capacity", counts[s], scap);
            const unsigned char *tail = (const unsigned char *)area[d] + stream_bytes + nranks * 8;
            if (tail[0] != 0x5C || tail[63] != 0x5C) fail("write behind a stream area");
        }
    } else {
        if (flags_scatter) fail("flags of the stream form", flags_scatter);
        if (nl_sum != N_l || nk_sum != N_k) fail("window counts of the stream form", nl_sum, N_l);
        std::map<u64, u64> seen_l;          // strand l-mer -> multiplicity, over all owners
        std::map<u64, u32> vertex_owner;
        u64 e_total = 0;
        for (u32 d = 0; d < nranks && !bad; d++) {
            const u64 *counts = (const u64 *)((const char *)area[d] + stream_bytes);
            u32 rcap = 1;
            for (u32 lb = 0; lb < nbpr; lb++) rcap = std::max<u32>(rcap, (u32)expect[(size_t)d * nbpr + lb].size());
            uint4 *regions = aligned_array<uint4>((size_t)nbpr * rcap, 0x5C);
            u32 *cursors = aligned_array<u32>(nbpr, 0);
            u64 *stats = aligned_array<u64>(64, 0);
            simt::launch(2, 256, 0, [&] { bkt_regroup_kernel(area[d], counts, nranks, scap, nbpr, rcap, regions, cursors, stats); });
            if (stats[2]) fail("flags of the regroup", stats[2]);
            Regions reg(nbpr);
            for (u32 lb = 0; lb < nbpr && !bad; lb++) {
                reg[lb].assign(regions + (size_t)lb * rcap, regions + (size_t)lb * rcap + std::min(cursors[lb], rcap));
                if (!same_multiset(reg[lb], expected_region(d * nbpr + lb))) fail("records of a bucket region (stream form)", d * nbpr + lb, reg[lb].size());
            }
            free(regions); free(cursors); free(stats);
            if (bad) break;
            // ---- the owner's graph
            const Graph G = run_build(reg, l, cs.cap, C.M.size() + 8, C.VS.size() + 8);
            if (G.flags) { fail("flags of the build", G.flags); break; }
            e_total += G.E;
            std::map<u64, u32> vid;
            for (u64 i = 0; i < G.V; i++) {
                vid[G.vkeys[i]] = (u32)i;
                if (vertex_owner.count(G.vkeys[i])) fail("a vertex on two owners", G.vkeys[i]);
                vertex_owner[G.vkeys[i]] = d;
            }
            for (u64 i = 0; i < G.V && !bad; i++)
                for (u32 t = 0; t < 4; t++) {
                    const u64 v = G.vkeys[i], out = (v << 2) | t, in = ((u64)t << (2 * k)) | v;
                    const u64 mo = C.M.count(out) ? C.M.at(out) : 0, mi = C.M.count(in) ? C.M.at(in) : 0;
                    if (G.lcount[4 * i + t] != mo) fail("lcount", G.lcount[4 * i + t], mo);
                    if (G.ecount[4 * i + t] != mi) fail("ecount", G.ecount[4 * i + t], mi);
                }
            for (u64 i = 0; i < G.U && !bad; i++) {
                const u64 x = G.lkeys[i];
                if (seen_l.count(x)) fail("a strand l-mer on two owners", x);
                seen_l[x] = G.lvals[i];
                if (G.ev1[i] >= G.V || G.vkeys[G.ev1[i]] != (x >> 2)) fail("prefix vertex of an edge", i);
                const bool local = vid.count(x & kmask) != 0;
                if (local ? (G.ev2[i] >= G.V || G.vkeys[G.ev2[i]] != (x & kmask)) : G.ev2[i] != EULER_NO_ID) fail("suffix vertex of an edge", i, G.ev2[i]);
            }
        }
        if (!bad) {
            if (seen_l != C.M) fail("union of the owners' edge records", seen_l.size(), C.M.size());
            if (vertex_owner.size() != C.VS.size()) fail("union of the owners' vertices", vertex_owner.size(), C.VS.size());
            for (auto v : C.VS)
                if (!vertex_owner.count(v)) { fail("a vertex on no owner", v); break; }
            if (e_total != 2 * C.N_l) fail("edge total over the owners", e_total, 2 * C.N_l);
        }
    }
    for (u32 d = 0; d < nranks; d++) free(area[d]);
    for (auto &s : shards) free(s.buf16);
    printf("l=%2u ranks=%u buckets/rank=%2u scap=%u: N_l=%llu U=%zu V=%zu fullest stream=%llu %s\n", l, nranks, nbpr, scap,
           (unsigned long long)C.N_l, C.M.size(), C.VS.size(), (unsigned long long)fullest, bad ? "FAILED" : "ok");
    fflush(stdout);
    return bad ? 1 : 0;
}

int main()
{
    const Case cases[] = {
        // l, ranks, buckets per rank, table slots, reads, maxlen, genome, stream capacity factor
        {32, 1, 6, 1536, 120, 130, 1200, 1.0},    // one GPU: direct form and a 1-rank stream form
        {22, 1, 5, 1536, 120, 100, 1000, 1.0},
        {13, 1, 4, 1536, 60, 90, 0, 1.0},         // generic kernel (k = m)
        {32, 3, 4, 1536, 160, 140, 1500, 1.0},    // three ranks, uneven shards
        {22, 8, 2, 1536, 200, 100, 1800, 1.0},    // eight ranks
        {6, 2, 3, 1536, 60, 60, 0, 1.0},          // short l-mers: palindromes, k < m
        {32, 16, 1, 1536, 120, 120, 900, 1.0},    // the most ranks the run reservation holds, one bucket each, some shards empty
        {22, 3, 3, 1536, 150, 120, -1, 1.0},      // homopolymers and tandem repeats: long runs of one minimizer, many chunk-leading pieces
        {32, 3, 4, 1536, 160, 140, 1500, 0.5},    // streams too small: flag, capped counts, nothing behind the area
    };
    int fails = 0, n = 0;
    for (const Case &c : cases) { fails += run_case(c); n++; }
    printf("%d cases, %d failed\n", n, fails);
    return fails ? 1 : 0;
}
