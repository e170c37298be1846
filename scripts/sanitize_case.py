"""Small parity configs for compute-sanitizer (memcheck / racecheck): 64-bit keys through the bucketed path and the
global-table path, 128-bit keys, and a 2-rank emulation of the bucketed multi-GPU path.  Every result is also checked
against the oracle, so a sanitizer-clean run is a correct run.
  compute-sanitizer --tool memcheck  python scripts/sanitize_case.py
  compute-sanitizer --tool racecheck python scripts/sanitize_case.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pycuda-euler_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import oracle
import _native as N
from util import random_reads

ctx = N.Context(0)
reads = random_reads(4, 250, genome_len=3000) + ["A" * 70, "ACGT" * 12, "ACGTN", ""]
buf, off = oracle.pack_reads(reads)


def check(l, tag):
    g = oracle.graph_build(buf, off, l, expand=True)
    st = ctx.run_host(buf, off, l, N.RUN_EXPAND_EDGES | N.RUN_CANONICAL_IDS)
    assert (st.distinct_lmers, st.distinct_kmers, st.edge_count) == (g.nl, g.nv, g.ne), (tag, l)
    assert np.array_equal(ctx.download(N.ART_LMER_KEYS), g.lk_lo) and np.array_equal(ctx.download(N.ART_EV), g.ev)
    assert np.array_equal(ctx.download(N.ART_EE), g.ee)
    contigs = ctx.pipeline_contigs()
    ref, _ = oracle.euler_contigs(buf, off, l)
    assert contigs == ref, (tag, l)
    print("ok", tag, l, "path", st.path, "U", st.distinct_lmers, flush=True)


for l in (32, 22, 12):
    check(l, "bucketed")
os.environ["EULER_B200_BUCKETED"] = "0"
check(32, "global-table")
del os.environ["EULER_B200_BUCKETED"]
check(48, "128-bit")

from eulercuda.dist import emulate_partitioned_bucketed
l, world = 32, 2
shards = [oracle.pack_reads(reads[r::world]) for r in range(world)]
parts, windows = emulate_partitioned_bucketed(ctx, shards, l, world)
g = oracle.graph_build(buf, off, l, expand=False)
all_l = np.concatenate([p["LMER_KEYS"] for p in parts])
all_m = np.concatenate([p["LMER_VALUES"] for p in parts])
o = np.argsort(all_l, kind="stable")
assert np.array_equal(all_l[o], g.lk_lo) and np.array_equal(all_m[o], g.lvals)
assert np.array_equal(np.sort(np.concatenate([p["KMER_KEYS"] for p in parts])), g.vk_lo)
print("ok 2-rank bucketed emulation", flush=True)
ctx.close()
print("ALL OK")
