"""Runs bench.py's main() WITHOUT a GPU, with the device replaced by stand-ins, so that the control flow of the
benchmark driver (the JSON line, its keys, the end-to-end and extra phases, the line watchdog, the collectives of the
N > 1 path over gloo) is exercised by the CPU test suite.  TEST INFRASTRUCTURE ONLY: nothing is measured here and no
kernel runs; the numbers in the printed line are made up by the stand-ins.

usage: python tests/bench_flow_mock.py [bench.py arguments]      (RANK / WORLD_SIZE / MASTER_* in the environment for N > 1)"""
import contextlib
import importlib.util
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pycuda-euler_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402


def _strip(fn):
    def wrapped(*a, **k):
        k.pop("device", None)
        k.pop("pin_memory", None)
        return fn(*a, **k)
    return wrapped


for name in ("tensor", "empty", "zeros", "arange", "full"):
    setattr(torch, name, _strip(getattr(torch, name)))


class _Event:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return 1e3 * (other.t - self.t) + 1e-3


class _Stream:
    cuda_stream = 0


torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a: None
torch.cuda.empty_cache = lambda: None
torch.cuda.Event = _Event
torch.cuda.Stream = _Stream
torch.cuda.stream = lambda s: contextlib.nullcontext()


class Stats:
    """what a run reports (made-up, self-consistent numbers)"""

    def __init__(self, reads, read_len, l):
        self.n_bases = reads * read_len
        self.n_kmer_windows = reads * (read_len - l + 2)
        self.n_lmer_windows = reads * (read_len - l + 1)
        self.distinct_lmers, self.distinct_kmers = 2000, 2002
        self.edge_count = 2 * self.n_lmer_windows
        self.retries = self.redo_buckets = 0
        self.ms_total, self.ms_count, self.ms_count_kernel, self.ms_build_kernel, self.ms_graph = 2.0, 0.5, 0.4, 1.5, 1.6
        self.kernel_launches, self.path, self.n_buckets, self.bucket_records = 6, 1, 100, 3000
        self.lmer_table_capacity = self.kmer_table_capacity = 4096

    def as_dict(self):
        return dict(self.__dict__)


class Context:
    def __init__(self, device=0):
        self.device = device

    def set_stream(self, s):
        pass

    def sync(self):
        pass

    def close(self):
        pass

    def synth_reads_dev(self, ptr, genome, read_len, err_ppm, first, count):
        self.read_len = read_len

    def run_dev(self, d_buf, d_off, nreads, n_bases, l, flags, hint):
        time.sleep(0.002)
        return Stats(nreads, n_bases // max(nreads, 1), l)

    def run_host_ptr(self, h_buf, h_off, nreads, l, flags, hint):
        time.sleep(0.002)
        return Stats(nreads, 100, l)

    def download_into(self, art, ptr, nbytes):
        return 1000


import _native as native  # noqa: E402  (the real binding module: constants and dtypes; only the device context is replaced)

native.Context = Context

spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)

if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    import torch.distributed as dist
    import eulercuda.dist as edist

    _init = dist.init_process_group
    dist.init_process_group = lambda backend, **k: _init("gloo", timeout=k.get("timeout"))
    bench.parity_check_partitioned = lambda runner: {"ok": True, "path": "stand-in"}

    def build_partitioned(ctx, d_reads, d_off, nreads, n_bases, l, rank, world, distinct_hint=0, group=None, **kw):
        t = torch.tensor([1.0])
        dist.all_reduce(t, group=group)      # the step is collective, like the real one
        st = Stats(nreads, n_bases // max(nreads, 1), l)
        info = {"n_lmer_windows": st.n_lmer_windows, "n_kmer_windows": st.n_kmer_windows, "exchange_bytes": 1, "exact_fallback": False,
                "transport": "stand-in", "phase_ms": {"build": 1.0}, "geometry": {"nb_per_rank": 1}}
        return st, info

    edist.build_partitioned = build_partitioned

if os.environ.get("BENCH_MOCK_STALL"):   # the end-to-end phase never comes back: the watchdog must print the line
    real_guard = bench.LineGuard
    bench.LineGuard = lambda seconds, line: real_guard(1.0 if seconds > 60 else seconds, line)   # (the 30 s one guards the shutdown)
    Context.run_host_ptr = lambda self, *a: time.sleep(3600)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        edist_build = edist.build_partitioned
        calls = {"n": 0}

        def stalling(*a, **k):
            calls["n"] += 1
            if calls["n"] > 8:      # warm-up and timed steps pass, the end-to-end steps stall
                time.sleep(3600)
            return edist_build(*a, **k)

        edist.build_partitioned = stalling

sys.argv = ["bench.py"] + sys.argv[1:]
bench.main()
