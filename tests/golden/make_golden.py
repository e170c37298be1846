#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference CPU
assembler (/root/reference/src/referenceassembler/referenceAssembler.py) in this container.

The reference is Python and cannot travel to the GPU box, so its outputs are committed here as
small JSON fixtures together with this script.  Re-run:  python tests/golden/make_golden.py

Fixtures
  g200.json         reference fixture tests/g200reads.fa (the 100 reads pinned by the
                    reference's own tests/test_fasta_reader.py) x k in {9,11,17,18,21} x
                    limit in {0,1}: build() table and all_contigs() output.
  synth_small.json  600 seeded random reads (with N's and ragged lengths) from a 3 kbp genome,
                    k in {15,16,31}: same outputs.
  graph_pins.json   the de Bruijn GRAPH artefacts as a pure function of the reference's own build(reads, l, 0)
                    (:25-42) and fw / bw (:16-23) for the same two read sets: the vertex set, per-vertex
                    lcount[4] / ecount[4] (multiplicity of the l-mer v+x / x+v) and the edge multiset
                    (l-mer, multiplicity).  Pins the degree arrays and edge lists of the oracle and of the GPU
                    path independently of the oracle's own restatement of pydebruijn.py.
"""
import hashlib
import json
import os
import random
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load_reference():
    # its only third-party import (`from dask import delayed`, :5) is unused
    sys.modules.setdefault("dask", types.SimpleNamespace(delayed=lambda f=None, **kw: f))
    sys.path.insert(0, os.path.join(REF, "src", "referenceassembler"))
    import referenceAssembler
    return referenceAssembler


def sha16(items):
    return hashlib.sha256("\n".join(sorted(items)).encode()).hexdigest()[:16]


def run_case(ra, reads, k, limit):
    d = ra.build(reads, k, limit)
    G, contigs = ra.all_contigs(d, k)
    canon = sorted(min(c, ra.twin(c)) for c in contigs)
    # get_contig from the first k-mer of every contig and from one in its middle (string only)
    probes = []
    for c in contigs:
        for km in (c[:k], c[(len(c) - k) // 2:(len(c) - k) // 2 + k]):
            probes.append([km, ra.get_contig(d, km)[0]])
    return {
        "k": k, "limit": limit,
        "links": [[[list(x) for x in G[i][0]], [list(x) for x in G[i][1]]] for i in range(len(contigs))],
        "get_contig": probes,
        "kmers": sorted([km, c] for km, c in d.items()),
        "kmer_sha": sha16(["%s\t%d" % (km, c) for km, c in d.items()]),
        "contigs": contigs,
        "contig_sha": sha16(canon),
    }


def graph_case(ra, reads, l):
    """vertices / degree slots / edges of the l-mer graph from the reference's build(), fw() and bw() alone"""
    d = ra.build(reads, l, 0)
    verts = sorted({x[:-1] for x in d} | {x[1:] for x in d})
    # fw('A' + v) yields v + x, bw(v + 'A') yields x + v  (x over 'ACGT')
    lcount = [[d.get(km, 0) for km in ra.fw("A" + v)] for v in verts]
    ecount = [[d.get(km, 0) for km in ra.bw(v + "A")] for v in verts]
    return {"l": l, "vertices": verts, "lcount": lcount, "ecount": ecount, "edges": sorted([x, c] for x, c in d.items())}


def read_fasta_lines(path):
    out = []
    with open(path) as f:
        for line in f:
            if line[0] != ">":
                out.append(line.strip())
    return out


def main():
    ra = load_reference()
    reads = read_fasta_lines(os.path.join(REF, "tests", "g200reads.fa"))
    g200 = {"source": "reference tests/g200reads.fa", "reads": reads,
            "cases": [run_case(ra, reads, k, lim) for k in (9, 11, 17, 18, 21) for lim in (0, 1)]}
    with open(os.path.join(HERE, "g200.json"), "w") as f:
        json.dump(g200, f, indent=0)

    rng = random.Random(20261018)
    genome = "".join(rng.choice("ACGT") for _ in range(3000))
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    sreads = []
    for _ in range(600):
        L = rng.choice((20, 33, 47, 60, 60, 60))
        s = rng.randrange(0, len(genome) - L)
        r = genome[s:s + L]
        if rng.random() < 0.5:
            r = "".join(comp[c] for c in reversed(r))
        if rng.random() < 0.1:
            p = rng.randrange(L)
            r = r[:p] + "N" + r[p + 1:]
        if rng.random() < 0.03:
            p = rng.randrange(L)
            r = r[:p] + rng.choice("ACGT") + r[p + 1:]
        sreads.append(r)
    synth = {"source": "random.Random(20261018), see make_golden.py", "reads": sreads,
             "cases": [run_case(ra, sreads, k, lim) for k in (15, 16, 31) for lim in (0, 1)]}
    with open(os.path.join(HERE, "synth_small.json"), "w") as f:
        json.dump(synth, f, indent=0)
    pins = {"source": "referenceAssembler.build(reads, l, 0) + fw/bw, see make_golden.py",
            "g200": [graph_case(ra, reads, l) for l in (10, 18)],
            "synth_small": [graph_case(ra, sreads, l) for l in (16, 32)]}
    with open(os.path.join(HERE, "graph_pins.json"), "w") as f:
        json.dump(pins, f, separators=(",", ":"))
    for name in ("g200", "synth_small"):
        for c in pins[name]:
            print("graph", name, c["l"], len(c["vertices"]), len(c["edges"]), sum(m for _, m in c["edges"]))
    for name, fx in (("g200", g200), ("synth_small", synth)):
        for c in fx["cases"]:
            print(name, c["k"], c["limit"], len(c["kmers"]), c["kmer_sha"], len(c["contigs"]), c["contig_sha"])


if __name__ == "__main__":
    main()
