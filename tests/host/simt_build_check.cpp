// simt_build_check.cpp -- the per-bucket build kernels of pycuda-euler_b200/csrc/bucket_build.cu (bkt_build_kernel, first
// and second pass, bkt_boundary_publish_kernel, bkt_fixup_kernel) run on the CPU under the SIMT emulator of simt_emu.h
// and checked against a brute-force statement of the graph they must produce.  TEST INFRASTRUCTURE ONLY: the product
// runs these kernels on the GPU; here the SAME SOURCE is compiled with -DEULER_SIMT_EMU, every CUDA thread is an OS
// thread, warp collectives are rendezvous that abort when lanes disagree about which collective they are at, and a
// deadlock shows as a hang (the caller runs this binary under a timeout).
//
// Input records come from the partition pass's lane logic (lane_driver.h: the bucket.cuh functions that
// bucket_lane_check.cpp verifies).
// Checked, for l = 32 / 22 (templated kernels) and generic lengths: edge records (both strands, multiplicities,
// offsets), vertices, the eight degree slots of every vertex, the scans, EulerVertex, prefix / suffix vertex ids of
// every edge (cross-bucket ones through the post-pass), the totals -- and the overflow protocol: buckets that do not
// fit the first pass's tables are rebuilt by the second pass (same graph), a bucket that fits neither raises
// BKT_FLAG_TABLE, artefact arrays that are too small raise BKT_FLAG_OUTPUT; nothing may hang or write out of bounds.
//
// Build: g++ -O1 -std=c++17 -pthread -I<csrc> -I<tests/host> simt_build_check.cpp -o simt_build_check
#define EULER_SIMT_EMU
#include "../../pycuda-euler_b200/csrc/bucket_build.cu"

#include "lane_driver.h"

struct Case {
    u32 l, nb, cap;
    int nreads, maxlen, genome;   // genome > 0: reads are windows of one random genome (coverage: repeated l-mers)
    double slack;                 // artefact capacity = exact size * slack (< 1: the OUTPUT flag is expected)
    u64 expect_flags;             // flags that must be set (0: a clean run whose graph is checked)
    bool expect_redo;             // the second pass must have rebuilt at least one bucket
};

static int run_case(const Case &cs, int rep)
{
    const u32 l = cs.l, k = l - 1;
    const BkGeom g = {1, cs.nb};
    const Reads R = make_reads(cs.nreads, cs.maxlen, cs.genome);
    const Census C = census(R, l);
    const std::map<u64, u64> &M = C.M;     // strand l-mer -> both-strand multiplicity
    const std::set<u64> &VS = C.VS;
    const u64 N_l = C.N_l;
    const u64 kmask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
    // ---- records per bucket from the partition pass's lane logic
    std::vector<std::vector<BkRec>> recs = host_records(R, l, g);
    u32 rcap = 1;
    for (auto &v : recs) rcap = std::max<u32>(rcap, (u32)v.size());
    uint4 *records = aligned_array<uint4>((size_t)cs.nb * rcap, 0);
    u32 *counts = aligned_array<u32>(cs.nb, 0);
    for (u32 b = 0; b < cs.nb; b++) {
        counts[b] = (u32)recs[b].size();
        // the arrival order of the records is arbitrary on the device: shuffle
        for (size_t i = recs[b].size(); i > 1; i--) std::swap(recs[b][i - 1], recs[b][rnd() % i]);
        for (size_t i = 0; i < recs[b].size(); i++) records[(size_t)b * rcap + i] = make_uint4(recs[b][i].hdr, recs[b][i].d[0], recs[b][i].d[1], recs[b][i].d[2]);
    }

    // ---- artefacts, state, launch (what bkt_build() does on the device)
    const u64 U = M.size(), V = VS.size();
    const u64 ucap = (u64)(U * cs.slack) + (cs.slack >= 1.0 ? 8 : 0), vcap = (u64)(V * cs.slack) + (cs.slack >= 1.0 ? 8 : 0);
    const int GUARD = 0x5C;   // bytes behind the capacities must stay untouched
    u64 *lkeys = aligned_array<u64>(ucap + 64, GUARD);
    u32 *lvals = aligned_array<u32>(ucap + 64, GUARD), *loffs = aligned_array<u32>(ucap + 64, GUARD);
    u32 *ev1 = aligned_array<u32>(ucap + 64, GUARD), *ev2 = aligned_array<u32>(ucap + 64, GUARD);
    u64 *vkeys = aligned_array<u64>(vcap + 64, GUARD);
    u32 *lcount = aligned_array<u32>(4 * vcap + 256, GUARD), *ecount = aligned_array<u32>(4 * vcap + 256, GUARD);
    u32 *lstart = aligned_array<u32>(4 * vcap + 256, GUARD), *estart = aligned_array<u32>(4 * vcap + 256, GUARD);
    euler_vertex *ev = aligned_array<euler_vertex>(vcap + 64, GUARD);
    const size_t zero_words = (size_t)cs.nb + 2 + 1 + BKT_REDO_CAP;
    u32 *state32 = aligned_array<u32>(zero_words + 8, 0);
    u64 *state64 = aligned_array<u64>(4ull * cs.nb, GUARD);   // published before they are read: garbage on entry
    u64 bcap = 64;
    while (bcap < U) bcap <<= 1;
    u64 *bkeys = aligned_array<u64>(bcap, 0xFF);
    u32 *bvals = aligned_array<u32>(2 * bcap, 0xFF);
    u64 *stats = aligned_array<u64>(64, 0);

    BkBuildArgs a;
    a.records = records; a.counts = counts; a.nb = cs.nb; a.nranks = 1; a.rcap = rcap; a.l = l; a.cap = cs.cap;
    a.lkeys = lkeys; a.lvals = lvals; a.loffs = loffs; a.ev1 = ev1; a.ev2 = ev2; a.ucap = ucap;
    a.vkeys = vkeys; a.lcount = lcount; a.ecount = ecount; a.lstart = lstart; a.estart = estart; a.ev = ev; a.vcap = vcap;
    a.flag = state32; a.ticket = state32 + cs.nb; a.redo = a.ticket + 2; a.second = 0;
    a.agg_uv = state64; a.agg_e = state64 + cs.nb; a.inc_uv = state64 + 2ull * cs.nb; a.inc_e = state64 + 3ull * cs.nb;
    a.bkeys = bkeys; a.bvals = bvals; a.bcap = bcap; a.stats = stats;
    auto run = [&](unsigned grid, size_t smem) {
        if (l == 32) simt::launch(grid, BB_THREADS, smem, [&] { bkt_build_kernel<32>(a); });
        else if (l == 22) simt::launch(grid, BB_THREADS, smem, [&] { bkt_build_kernel<22>(a); });
        else simt::launch(grid, BB_THREADS, smem, [&] { bkt_build_kernel<0>(a); });
    };
    run(cs.nb, (size_t)29 * cs.cap);
    const u32 nredo = a.redo[0];
    a.second = 1; a.cap = BKT_MAX_CAP;
    run((nredo < BKT_REDO_CAP ? nredo : BKT_REDO_CAP) + 2, (size_t)29 * BKT_MAX_CAP);   // two blocks find nothing to do
    simt::launch(3, 256, 0, [&] { bkt_boundary_publish_kernel(lkeys, ev1, ev2, stats + 3, ucap, l, bkeys, bvals, bcap, stats); });
    simt::launch(3, 256, 0, [&] { bkt_fixup_kernel(lkeys, ev2, stats + 3, ucap, l, bkeys, bvals, bcap); });

    // ---- checks
    int bad = 0;
    auto fail = [&](const char *what, u64 x = 0, u64 y = 0) {
        if (bad++ < 8) fprintf(stderr, "  case l=%u nb=%u cap=%u rep=%d: %s (%llu vs %llu)\n", l, cs.nb, cs.cap, rep, what, x, y);
    };
    const u64 flags = stats[2];
    if (stats[7] != nredo) fail("stats[7] is not the number of listed buckets", stats[7], nredo);
    if (cs.expect_redo && nredo == 0) fail("the second pass had nothing to do");
    // guards: nothing behind the capacities may have been written
    for (u64 i = ucap; i < ucap + 64; i++)
        if (lvals[i] != 0x5C5C5C5Cu || ev2[i] != 0x5C5C5C5Cu || lkeys[i] != 0x5C5C5C5C5C5C5C5Cull) { fail("write behind the edge capacity", i); break; }
    for (u64 i = vcap; i < vcap + 64; i++)
        if (vkeys[i] != 0x5C5C5C5C5C5C5C5Cull || lcount[4 * i] != 0x5C5C5C5Cu) { fail("write behind the vertex capacity", i); break; }
    if (cs.expect_flags) {
        if ((flags & cs.expect_flags) != cs.expect_flags) fail("expected flags not raised", flags, cs.expect_flags);
        if (flags & BKT_FLAG_INTERNAL) fail("internal consistency flag", flags);
    } else {
        if (flags) fail("flags raised on a run that fits", flags);
        if (stats[3] != U) fail("U", stats[3], U);
        if (stats[4] != V) fail("V", stats[4], V);
        if (stats[5] != 2 * N_l) fail("E", stats[5], 2 * N_l);
        if (!bad) {
            std::vector<std::pair<u64, u64>> got, exp(M.begin(), M.end());
            for (u64 i = 0; i < U; i++) got.push_back({lkeys[i], lvals[i]});
            std::sort(got.begin(), got.end());
            if (got != exp) fail("edge records (key, multiplicity)");
            std::map<u64, u32> vid;
            for (u64 i = 0; i < V; i++) vid[vkeys[i]] = (u32)i;
            if (vid.size() != V || !std::equal(VS.begin(), VS.end(), vid.begin(), [](u64 x, const std::pair<const u64, u32> &y) { return x == y.first; }))
                fail("vertex set");
            u64 run_l = 0, run_e = 0, run_o = 0;
            for (u64 i = 0; i < V && !bad; i++) {
                const u64 v = vkeys[i], rv = bk_revcomp(v, k);
                u64 ls = 0, es = 0;
                for (u32 t = 0; t < 4; t++) {
                    const u64 out = (v << 2) | t, in = ((u64)t << (2 * k)) | v;
                    const u64 mo = M.count(out) ? M.at(out) : 0, mi = M.count(in) ? M.at(in) : 0;
                    if (lcount[4 * i + t] != mo) fail("lcount", lcount[4 * i + t], mo);
                    if (ecount[4 * i + t] != mi) fail("ecount", ecount[4 * i + t], mi);
                    if (lstart[4 * i + t] != (u32)(run_l + ls)) fail("lstart", lstart[4 * i + t], run_l + ls);
                    if (estart[4 * i + t] != (u32)(run_e + es)) fail("estart", estart[4 * i + t], run_e + es);
                    ls += mo; es += mi;
                }
                if (ev[i].vid != v || ev[i].lp != (u32)run_l || ev[i].ep != (u32)run_e || ev[i].lcount != ls || ev[i].ecount != es) fail("EulerVertex", i);
                run_l += ls; run_e += es;
                if (v != rv) {   // the two strands of a vertex have adjacent ids
                    const u32 j = vid[rv];
                    if (j + 1 != i && i + 1 != j) fail("strands of a vertex are not adjacent", i, j);
                }
            }
            if (run_l != 2 * N_l || run_e != 2 * N_l) fail("degree totals", run_l, 2 * N_l);
            for (u64 i = 0; i < U && !bad; i++) {
                const u64 x = lkeys[i];
                if (loffs[i] != (u32)run_o) fail("lmer offsets", loffs[i], run_o);
                run_o += lvals[i];
                if (ev1[i] >= V || vkeys[ev1[i]] != (x >> 2)) fail("prefix vertex of an edge", i, ev1[i]);
                if (ev2[i] >= V || vkeys[ev2[i]] != (x & kmask)) fail("suffix vertex of an edge", i, ev2[i]);
            }
        }
    }
    printf("l=%2u nb=%3u cap=%4u rep=%d: U=%llu V=%llu N_l=%llu records=%zu fullest=%u redo=%u flags=%llx %s\n", l, cs.nb, cs.cap, rep,
           (unsigned long long)U, (unsigned long long)V, (unsigned long long)N_l, [&] { size_t s = 0; for (auto &v : recs) s += v.size(); return s; }(),
           rcap, nredo, (unsigned long long)flags, bad ? "FAILED" : "ok");
    fflush(stdout);
    free(records); free(counts); free(lkeys); free(lvals); free(loffs); free(ev1); free(ev2); free(vkeys); free(lcount); free(ecount);
    free(lstart); free(estart); free(ev); free(state32); free(state64); free(bkeys); free(bvals); free(stats);
    return bad ? 1 : 0;
}

int main(int argc, char **argv)
{
    const int reps = argc > 1 ? atoi(argv[1]) : 1;
    const Case cases[] = {
        // l, nb, cap, reads, maxlen, genome, slack, expected flags, second pass expected
        {32, 6, 1536, 150, 130, 1500, 1.0, 0, false},            // the benchmark length: coverage, both strands, N's
        {22, 5, 1536, 150, 100, 1200, 1.0, 0, false},            // the other templated length
        {13, 4, 768, 120, 90, 0, 1.0, 0, false},                 // generic kernel; k = m: one k-mer per minimizer window
        {6, 3, 768, 60, 60, 0, 1.0, 0, false},                   // short l-mers: palindromes, dense graph, many repeats
        {2, 2, 256, 30, 40, 0, 1.0, 0, false},                   // k = 1
        {31, 5, 1536, 120, 130, 1200, 1.0, 0, false},            // odd lengths take the generic kernel
        {16, 4, 1536, 100, 100, 800, 1.0, 0, false},
        {32, 3, 1536, 300, 120, -1, 1.0, 0, false},              // homopolymers and tandem repeats: multiplicities in the thousands
        {8, 2, 256, 200, 100, -1, 1.0, 0, false},                // the same with palindromic l-mers (even l)
        {32, 1, 256, 40, 100, 0, 1.0, 0, true},                  // one bucket that cannot fit 256 slots: rebuilt by the second pass
        {22, 7, 256, 260, 100, 2500, 1.0, 0, true},              // several buckets overflow, the others do not
        {32, 1, 256, 420, 150, 0, 1.0, BKT_FLAG_TABLE, true},    // too large even for the second pass: the host must repartition
        {32, 4, 1536, 100, 120, 1000, 0.5, BKT_FLAG_OUTPUT, false},   // artefact arrays too small: flag, no write behind them
    };
    int fails = 0, n = 0;
    for (int rep = 0; rep < reps; rep++)
        for (const Case &c : cases) { fails += run_case(c, rep); n++; }
    printf("%d cases, %d failed\n", n, fails);
    return fails ? 1 : 0;
}
