"""Single-GPU emulation of one rank of the partitioned path at world = 8 (profiling aid).
Builds rank 0's received key set from 8 read shards (config-2 weak-scaling shape), then repeats the
scatter of shard 0 and the build so that ncu / CUDA events see steady-state launches."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pycuda-euler_b200"))
import _native  # noqa: E402

WORLD = int(os.environ.get("WORLD", "8"))
L, LMER = 100, 32
G = int(os.environ.get("G_PER_RANK", "4600000")) * WORLD
NREADS = 1_380_000
REPS = int(os.environ.get("REPS", "3"))

ctx = _native.default_context()
off = torch.arange(0, (NREADS + 1) * L, L, dtype=torch.int64, device="cuda")
reads = torch.empty(NREADS * L + 16, dtype=torch.uint8, device="cuda")
windows = NREADS * (L - LMER + 1)
seg_cap = int(windows * 1.25 * 1.15 / WORLD) + 4096
send = torch.empty(seg_cap * WORLD, dtype=torch.int64, device="cuda")
recv = torch.empty(seg_cap * WORLD, dtype=torch.int64, device="cuda")
recv_counts = []
for r in range(WORLD):
    ctx.synth_reads_dev(reads.data_ptr(), G, L, 0, r * NREADS, NREADS)
    counts = ctx.dist_scatter_segments(reads.data_ptr(), off.data_ptr(), NREADS, NREADS * L, LMER, WORLD, send.data_ptr(), seg_cap)
    c0 = int(counts[0])
    recv[r * seg_cap:r * seg_cap + c0] = send[:c0]
    recv_counts.append(c0)
torch.cuda.synchronize()
print("recv keys", sum(recv_counts), "windows/shard", windows)
ctx.synth_reads_dev(reads.data_ptr(), G, L, 0, 0, NREADS)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(REPS):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.dist_scatter_segments(reads.data_ptr(), off.data_ptr(), NREADS, NREADS * L, LMER, WORLD, send.data_ptr(), seg_cap)
    t1 = time.perf_counter()
    st = ctx.dist_build_regions(recv.data_ptr(), seg_cap, recv_counts, LMER, 0, WORLD, 0)
    t2 = time.perf_counter()
    print("scatter %.3f ms  build %.3f ms (count %.3f graph %.3f, count kernel %.3f)  U_l %d V %d" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), st.ms_count, st.ms_graph, st.ms_count_kernel, st.distinct_lmers, st.distinct_kmers))
