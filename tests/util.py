"""Shared helpers for the parity tests (inputs only; no product code here)."""
import random

import numpy as np


def random_reads(seed, n, genome_len=4000, lens=(20, 35, 64, 100, 100, 150), n_frac=0.1, lower_frac=0.05,
                 alphabet="ACGT"):
    rng = random.Random(seed)
    genome = "".join(rng.choice(alphabet) for _ in range(genome_len))
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    reads = []
    for _ in range(n):
        L = rng.choice(lens)
        s = rng.randrange(0, genome_len - L)
        r = genome[s:s + L]
        if rng.random() < 0.5:
            r = "".join(comp[c] for c in reversed(r))
        if rng.random() < n_frac:
            p = rng.randrange(L)
            r = r[:p] + "N" + r[p + 1:]
        if rng.random() < 0.02:
            r = r[:rng.randrange(1, 5)]
        reads.append(r)
    return reads


def lower_some(reads, seed, frac=0.3):
    rng = random.Random(seed)
    return [r.lower() if rng.random() < frac else r for r in reads]


def check_graph_against_pins(pin, vkeys, lcount, ecount, lkeys, lvals, ev1, ev2, decode_key):
    """Compare graph artefacts (ids = rank in ascending key order) with a case of tests/golden/graph_pins.json,
    which make_golden.py derived from the UNMODIFIED reference's build(reads, l, 0) and fw / bw alone:
    vertex set, per-vertex lcount[4] / ecount[4], edge multiset and the (prefix, suffix) vertex of every edge."""
    l = pin["l"]
    verts = [decode_key(int(v), 0, l - 1) for v in vkeys]
    assert verts == pin["vertices"]
    assert np.asarray(lcount, np.uint32).reshape(-1, 4).tolist() == pin["lcount"]
    assert np.asarray(ecount, np.uint32).reshape(-1, 4).tolist() == pin["ecount"]
    edges = [[decode_key(int(x), 0, l), int(m)] for x, m in zip(lkeys, lvals)]
    assert edges == pin["edges"]
    for (x, _), a, b in zip(edges, ev1, ev2):
        assert verts[int(a)] == x[:-1] and verts[int(b)] == x[1:]
