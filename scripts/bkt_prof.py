"""Runs the bench workload a few times through the bucketed path (profiling target for ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pycuda-euler_b200")):
    sys.path.insert(0, p)
import torch
import _native as N

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = N.Context(0)
G, L, cov, l = 4_600_000, 100, 30, 32
R = G * cov // L
d_reads = torch.empty(R * L, dtype=torch.uint8, device="cuda")
ctx.synth_reads_dev(d_reads.data_ptr(), G, L, 0, 0, R)
d_off = torch.arange(R + 1, dtype=torch.int64, device="cuda") * L
ctx.sync()
for it in range(steps):
    st = ctx.run_dev(d_reads.data_ptr(), d_off.data_ptr(), R, R * L, l, 0, G)
print("path=%d nb=%d ms total %.3f part %.3f build %.3f" % (st.path, st.n_buckets, st.ms_total, st.ms_count_kernel, st.ms_build_kernel))
