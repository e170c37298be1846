"""CPU tests (no GPU): the oracle against the golden fixtures produced by the unmodified reference
CPU assembler, against the hand-derived kernel known answers of SURVEY §8c, and its own
internal consistency (Euler decomposition properties)."""
import hashlib
import json
import os
from collections import Counter

import numpy as np
import pytest

import oracle
from util import check_graph_against_pins, random_reads

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# (k, limit) -> (distinct, sum, kmer sha, contigs, contig sha) obtained by running
# /root/reference/src/referenceassembler/referenceAssembler.py unmodified on tests/g200reads.fa
SURVEY_G200 = {
    (9, 1): (374, 2234, "5e1d268e6b27eb8a", 2, "19948bc781401203"),
    (9, 0): (376, 2236, "b46d4c8728bc4e56", 1, "1f8767e1fe471434"),
    (11, 1): (364, 1846, "f795ab2f7c9ab781", 3, "7eb01d08f2d35cf9"),
    (11, 0): (372, 1854, "e021a5877cae713c", 1, "1f8767e1fe471434"),
    (17, 1): (210, 616, "a8548b611843707f", 17, "f5eba92361f867cb"),
    (17, 0): (320, 726, "fe18b3702316a6c6", 13, "88587aca392df312"),
    (18, 1): (152, 410, "f56a11ff3b91124b", 21, "78d3359ad4f40513"),
    (18, 0): (284, 542, "292f60e77fba92b1", 18, "b1027c044080ab8b"),
    (21, 1): (0, 0, "e3b0c44298fc1c14", 0, "e3b0c44298fc1c14"),
}


def _sha(items):
    return hashlib.sha256("\n".join(sorted(items)).encode()).hexdigest()[:16]


def _load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def test_fixture_reads_are_the_reference_fixture():
    fx = _load("g200.json")
    assert len(fx["reads"]) == 100 and sum(len(r) for r in fx["reads"]) == 1899
    assert fx["reads"][0] == "TGGGATAATATGGTACGATC" and fx["reads"][11] == "GCC"


@pytest.mark.parametrize("fixture", ["g200.json", "synth_small.json"])
def test_python_restatement_equals_reference_outputs(fixture):
    fx = _load(fixture)
    for case in fx["cases"]:
        d = oracle.py_build(fx["reads"], case["k"], case["limit"])
        assert sorted([km, c] for km, c in d.items()) == case["kmers"]
        assert _sha(["%s\t%d" % kv for kv in d.items()]) == case["kmer_sha"]
        contigs = oracle.py_all_contigs(d, case["k"])
        assert contigs == case["contigs"]          # same dict order => identical, cycles included
        assert _sha(oracle.canonical_contigs(contigs)) == case["contig_sha"]


def test_golden_digests_match_the_survey_table():
    fx = _load("g200.json")
    for case in fx["cases"]:
        key = (case["k"], case["limit"])
        if key not in SURVEY_G200:
            continue
        n, total, ksha, nc, csha = SURVEY_G200[key]
        assert len(case["kmers"]) == n and sum(c for _, c in case["kmers"]) == total
        assert case["kmer_sha"] == ksha and len(case["contigs"]) == nc and case["contig_sha"] == csha


@pytest.mark.parametrize("fixture", ["g200.json", "synth_small.json"])
def test_c_counting_equals_reference_build(fixture):
    """C restatement of the both-strand multiset == referenceAssembler.build(reads, k, limit=0)."""
    fx = _load(fixture)
    buf, off = oracle.pack_reads(fx["reads"])
    for case in fx["cases"]:
        if case["limit"] != 0:
            continue
        lo, hi, vals = oracle.count_mers(buf, off, case["k"])
        got = sorted([oracle.decode_key(a, b, case["k"]), int(c)] for a, b, c in zip(lo, hi, vals))
        assert got == case["kmers"]


def test_kernel_known_answers():
    """SURVEY §8c 'kernel-spec known answers' (derived from the reference kernel sources)."""
    assert oracle.encode_str("ACGT") == 27
    assert oracle.encode_str("T" * 32) == 2 ** 64 - 1 and oracle.encode_str("A" * 32) == 0
    f, r, v = oracle.encode_positions(*oracle.pack_reads(["TGGGATAATATGGTACGATC"]), 10)
    assert f[0] == 959244 == 0xEA30C and r[0] == 848724 and v[:11].all() and not v[11:].any()
    assert oracle.decode_key(848724, 0, 10) == "TATTATCCCA" == oracle.twin("TGGGATAATA")
    pk, sk = oracle.compute_kmers(np.array([959244], np.uint64), (1 << 18) - 1)
    assert pk[0] == 239811 and sk[0] == 172812
    assert 959244 & 3 == 0 and (959244 >> 18) & 3 == 3         # transitionTo / transitionFrom
    assert [oracle.hash_h(959244, b) for b in (1, 7, 409, 1000003)] == [0, 6, 22, 209841]
    assert oracle.hash_h(0, 409) == 96 and oracle.hash_h(27, 409) == 76


def test_windows_never_cross_reads_or_non_acgt():
    reads = ["ACGTN", "ACGT", "NNNN", "acgtacgt", "AC-GT"]
    buf, off = oracle.pack_reads(reads)
    f, r, v = oracle.encode_positions(buf, off, 4)
    starts = np.flatnonzero(v).tolist()
    assert starts == [0, 5, 13, 14, 15, 16, 17]
    assert f[0] == f[5] == f[13] == 27            # case-insensitive, "ACGT"
    lo, hi, vals = oracle.count_mers(buf, off, 4)
    d = {oracle.decode_key(a, b, 4): int(c) for a, b, c in zip(lo, hi, vals)}
    assert d["ACGT"] == 8                          # palindrome: counted on both strands
    assert sum(d.values()) == 2 * 7


@pytest.mark.parametrize("l", [4, 10, 17, 32])
def test_graph_invariants_and_euler_decomposition(l):
    reads = random_reads(11, 300, genome_len=2500)
    buf, off = oracle.pack_reads(reads)
    g = oracle.graph_build(buf, off, l, expand=True)
    assert g.ne == int(g.lvals.sum()) == int(g.ev["lcount"].sum()) == int(g.ev["ecount"].sum())
    assert np.array_equal(np.sort(g.lev), np.arange(g.ne)) and np.array_equal(np.sort(g.ent), np.arange(g.ne))
    assert np.array_equal(np.bincount(g.ee["v1"], minlength=g.nv), g.ev["lcount"])
    assert np.array_equal(np.bincount(g.ee["v2"], minlength=g.nv), g.ev["ecount"])
    # ids = rank in ascending key order (B14)
    assert (np.diff(g.vk_lo.astype(np.float64)) > 0).all() and (np.diff(g.lk_lo.astype(np.float64)) > 0).all()
    contigs, _ = oracle.euler_contigs(buf, off, l)
    cnt = Counter()
    for c in contigs:
        for i in range(len(c) - l + 1):
            cnt[c[i:i + l]] += 1
    d = {oracle.decode_key(a, b, l): int(v) for a, b, v in zip(g.lk_lo, g.lk_hi, g.lvals)}
    assert cnt == d                                # every edge is spelled exactly `multiplicity` times
    assert sum(len(c) - (l - 1) for c in contigs) == g.ne


def test_wide_keys_k63():
    """l = 64 (k = 63): 128-bit keys in the oracle (CPU only)."""
    reads = random_reads(5, 60, genome_len=1500, lens=(100, 150), n_frac=0.0)
    buf, off = oracle.pack_reads(reads)
    lo, hi, vals = oracle.count_mers(buf, off, 64)
    d = oracle.py_build(reads, 64, 0)
    got = {oracle.decode_key(a, b, 64): int(c) for a, b, c in zip(lo, hi, vals)}
    assert got == d
    g = oracle.graph_build(buf, off, 64, expand=True)
    assert g.ne == sum(d.values())


def test_synthetic_generator_is_deterministic_and_sliceable():
    a = oracle.synth_reads(50000, 100, err_ppm=10000, first=0, count=300)
    b = oracle.synth_reads(50000, 100, err_ppm=10000, first=100, count=100)
    assert np.array_equal(a[100 * 100:200 * 100], b)
    assert set(np.unique(a).tolist()) <= set(b"ACGT")
    clean = oracle.synth_reads(50000, 100, err_ppm=0, first=0, count=300)
    frac = (a != clean).mean()
    assert 0.005 < frac < 0.015                    # ~1 % substitutions


@pytest.mark.parametrize("name", ["g200", "synth_small"])
def test_oracle_graph_equals_the_reference_derived_pins(name):
    """Degree arrays, vertex set and edge lists of the oracle's D1-D6 restatement against pins derived from the
    unmodified reference's build(reads, l, 0) + fw / bw (tests/golden/graph_pins.json, make_golden.py)."""
    reads = _load("g200.json" if name == "g200" else "synth_small.json")["reads"]
    buf, off = oracle.pack_reads(reads)
    for pin in _load("graph_pins.json")[name]:
        g = oracle.graph_build(buf, off, pin["l"], expand=False)
        check_graph_against_pins(pin, g.vk_lo, g.lcount, g.ecount, g.lk_lo, g.lvals, g.ev1, g.ev2, oracle.decode_key)
