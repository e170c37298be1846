#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (encode -> canonical l-mer table -> de Bruijn
graph build) in forward k-mer windows per second (BASELINE.json `metric`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One process per GPU (torchrun for N > 1).  A step is one pass of the hot path over one batch of
synthetic reads.  Rank 0 prints ONE JSON line.

* value      k-mer windows / s, reads already resident in HBM (CUDA events, max over ranks)
* e2e        the same through the host-buffer C-ABI call: pinned host reads -> H2D -> pipeline ->
             D2H of the compressed graph, all inside the timed region
* roofline   the dominant kernel (fused encode+count) against the measured HBM peak
* cpu_baseline / --impl reference   the oracle's C restatement of the reference CPU path
             (kind "port": the reference is pure Python and cannot be compiled or shipped), timed on
             this box's host cores on a bounded sample of the same workload
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "pycuda-euler_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on for 1 GPU
    "ecoli_4.6Mbp_100bp_30x_k31": dict(G=4_600_000, L=100, cov=30, err_ppm=0, k=31),
    # configs[2] (per-GPU share when sharded)
    "100Mbp_150bp_40x_1pct_k31": dict(G=100_000_000, L=150, cov=40, err_ppm=10_000, k=31),
    # configs[3]: 1 Gbp, 8 GPUs (strong scaling: the data set is split over the ranks)
    "1Gbp_150bp_30x_k31": dict(G=1_000_000_000, L=150, cov=30, err_ppm=0, k=31),
    "small_smoke": dict(G=200_000, L=100, cov=20, err_ppm=0, k=31),
}
DEFAULT_WORKLOAD = "ecoli_4.6Mbp_100bp_30x_k31"
METRIC = "k-mers/sec (encode+hash+graph build, device-timed)"
UNIT = "k-mers/s"


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {
                getattr(pynvml, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(pynvml, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(pynvml, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(pynvml, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            while not self.stop_flag:
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.002)
        except Exception as e:  # NVML not available: report that instead of inventing numbers
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def algorithmic_bytes(st, w=8):
    """A = B + N_l (w+8) + U_l (w+16) + U_k (2w+84); the count kernel's share is the
    first two terms (every base read once, one table-slot touch per forward l-mer window)."""
    B, Nl, Ul, Uk = st.n_bases, st.n_lmer_windows, st.distinct_lmers, st.distinct_kmers
    kernel = B + Nl * (w + 8)
    return kernel, kernel + Ul * (w + 16) + Uk * (2 * w + 84)


# ---------------------------------------------------------------------------------------------------------------
# CPU baselines.  `reference` = the UNMODIFIED referenceAssembler.build (referenceAssembler.py:25-42) fanned out
# the way the reference does (dask map_partitions(build) / Spark reduceByKey): multiprocessing.Pool(P)
# over contiguous read shards, dict-sum merge in the parent.  The module is loaded by oracle/ref_loader.py: from
# /root/reference in the authoring container, from the bytecode in oracle/_ref/ on the GPU box.  `port` = the oracle's
# C restatement (OpenMP), used when the reference is not available and reported beside it.
# Sample = a full-coverage data set of a SMALLER genome from the same generator (same read length, coverage and error
# rate), so the distinct-key count scales with the sample like it does in the workload; the rate is the slope between
# two sample sizes (fixed costs such as the pool start-up cancel).
_REF_READS = None   # inherited by the forked pool workers


def _ref_build_shard(args):
    lo, hi, k = args
    from oracle import ref_loader
    ra = ref_loader.load()
    return ra.build(_REF_READS[lo:hi], k, 0)


def _synth_read_strings(wl, G):
    import oracle
    L = wl["L"]
    R = -(-G * wl["cov"] // L)
    buf = oracle.synth_reads(G, L, err_ppm=wl["err_ppm"], first=0, count=R).tobytes()
    return [buf[i * L:(i + 1) * L].decode("ascii") for i in range(R)], R * (L - wl["k"] + 1)


def reference_build_seconds(wl, G, procs):
    """one pass of the reference's build() over a cov-x data set of a G-bp genome: (seconds, k-mer windows, distinct)"""
    global _REF_READS
    import multiprocessing as mp
    _REF_READS, nk = _synth_read_strings(wl, G)
    R = len(_REF_READS)
    bounds = [(R * i // procs, R * (i + 1) // procs, wl["k"]) for i in range(procs)]
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        pool.map(_ref_build_shard, [(0, min(R, 64), wl["k"])] * procs)   # workers started and the module loaded
        t0 = time.perf_counter()
        parts = pool.map(_ref_build_shard, bounds, chunksize=1)
        total = parts[0]
        for d in parts[1:]:   # dict-sum merge (the reduceByKey of src/ref_spark.py:84)
            for km, c in d.items():
                total[km] = total.get(km, 0) + c
        dt = time.perf_counter() - t0
    _REF_READS = None
    return dt, nk, len(total)


def port_build_seconds(wl, G):
    """the oracle's C restatement (encode + both-strand tables + graph build) on the same kind of sample"""
    import oracle
    L, l = wl["L"], wl["k"] + 1
    R = -(-G * wl["cov"] // L)
    buf = oracle.synth_reads(G, L, err_ppm=wl["err_ppm"], first=0, count=R)
    off = oracle.fixed_offsets(R, L)
    t0 = time.perf_counter()
    oracle.graph_build(buf, off, l, expand=False)
    return time.perf_counter() - t0, R * (L - wl["k"] + 1), 0


def cpu_rate(wl, G1, procs, want="auto"):
    """k-mer windows / s of the CPU path from the slope between samples of G1 and 2*G1 bp.  Returns a dict."""
    from oracle import ref_loader
    use_ref = want != "port" and ref_loader.load() is not None and wl["k"] <= 1000
    fn = (lambda G: reference_build_seconds(wl, G, procs)) if use_ref else (lambda G: port_build_seconds(wl, G))
    t1, n1, _ = fn(G1)
    t2, n2, _ = fn(2 * G1)
    slope = (n2 - n1) / max(t2 - t1, 1e-9)
    return {"value": slope, "unit": UNIT, "cores": procs if use_ref else int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1)),
            "kind": "reference" if use_ref else "port",
            "sample": "%d x data sets of %d bp and %d bp genomes from the workload's generator (%d / %d k-mer windows, %.2f s / %.2f s); "
                      "rate = slope between the two" % (wl["cov"], G1, 2 * G1, n1, n2, t1, t2),
            "rates_raw": [n1 / t1, n2 / t2], "seconds": t1 + t2,
            "code": ("referenceAssembler.build, unmodified (%s), multiprocessing.Pool(%d) over read shards + dict-sum merge" % (ref_loader.kind(), procs))
            if use_ref else "oracle/euler_oracle.c graph_build (C restatement, OpenMP)"}


def run_reference(args, wl, rank, world):
    """--impl reference: the reference's own CPU path on this box's host cores, bounded sample per step."""
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(procs)   # torchrun pins it to 1: the port's thread count must be explicit
    G1 = args.ref_genome or max(20_000, min(wl["G"] // 2, 12_500 * procs))
    for _ in range(min(args.warmup, 1)):
        cpu_rate(wl, max(G1 // 8, 10_000), procs)
    vals, last = [], None
    t_budget = time.perf_counter()
    for i in range(args.steps):
        last = cpu_rate(wl, G1, procs)
        vals.append(last["value"])
        if time.perf_counter() - t_budget > 150 and i + 1 >= 2:   # keep the arm within a few minutes
            break
    value = sorted(vals)[len(vals) // 2]
    nk_step = wl["cov"] * 3 * G1 // wl["L"] * (wl["L"] - wl["k"] + 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * last["seconds"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": args.workload, "k": wl["k"], "sample": last["sample"], "code": last["code"],
                   "steps_requested": args.steps,
                   "note": "median over steps of the slope rate; each step runs the two sample sizes once"},
        "cpu_baseline": {k: last[k] for k in ("unit", "cores", "kind", "sample")} | {"value": value},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = genome and reads grow with N (default); strong = the named data set is split over the ranks")
    ap.add_argument("--k", type=int, default=0,
                    help="override the workload's k (BASELINE configs[4] k sweep: 21 / 31 / 63; k = 63 uses 128-bit keys)")
    ap.add_argument("--ref-genome", type=int, default=0, help="--impl reference: genome size of the smaller of the two samples")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = dict(WORKLOADS[args.workload])
    if args.k:
        wl["k"] = args.k
    wl["R"] = -(-wl["G"] * wl["cov"] // wl["L"])

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import numpy as np
    import torch
    import _native as N

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = N.Context(local_rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)

    L, k, l = wl["L"], wl["k"], wl["k"] + 1
    # Weak scaling: the genome (= the k-mer space) grows with the GPU count, every rank encodes its own
    # R reads of the shared data set and owns 1/world of the k-mer space (hash partition, one all-to-all).
    strong = args.scaling == "strong" and world > 1
    G = wl["G"] if strong else wl["G"] * world
    R = -(-wl["R"] // world) if strong else wl["R"]
    first = rank * R
    d_reads = torch.empty(R * L, dtype=torch.uint8, device="cuda")
    ctx.synth_reads_dev(d_reads.data_ptr(), G, L, wl["err_ppm"], first, R)
    d_off = torch.arange(R + 1, dtype=torch.int64, device="cuda") * L
    ctx.sync()
    torch.cuda.synchronize()
    # a user knows the genome size; per rank ~2/world of the canonical l-mers are incident to owned vertices
    if wl["err_ppm"] == 0:
        hint = wl["G"] if world == 1 else int((G / world) * 1.15)   # ~1.05 copies with minimizer ownership
    else:
        hint = 0

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if world > 1:
        from eulercuda.dist import build_partitioned

    def step():
        if world == 1:
            return ctx.run_dev(d_reads.data_ptr(), d_off.data_ptr(), R, R * L, l, 0, hint), None
        return build_partitioned(ctx, d_reads, d_off, R, R * L, l, rank, world, hint)

    st = None
    for _ in range(args.warmup):
        st, info = step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms, graph_ms, launches = 0.0, 0.0, 0
    per_step = []   # the library's own CUDA-event time of each step (N = 1) / host wall time of each step (N > 1)
    t_wall0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        t_s = time.perf_counter()
        st, info = step()
        per_step.append(st.ms_total if world == 1 else 1e3 * (time.perf_counter() - t_s))
        kern_ms += st.ms_count_kernel
        graph_ms += st.ms_graph
        launches += st.kernel_launches + (3 if world > 1 else 0)   # + mark_starts, count pass, scatter pass
    e1.record(stream)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0) / args.steps
    # the partitioned step synchronises the host around the collective: wall clock is the honest time there
    ms = e0.elapsed_time(e1) / args.steps if world == 1 else wall_ms
    clocks = sampler.result()
    kern_ms /= args.steps
    graph_ms /= args.steps
    if world > 1:
        st.n_kmer_windows, st.n_lmer_windows, st.n_bases = info["n_kmer_windows"], info["n_lmer_windows"], R * L

    # ---- e2e: host buffers through the reference-facing C-ABI call, copies inside the timed region
    e2e = None
    if not args.no_e2e and world == 1:
        h_reads = torch.empty(R * L, dtype=torch.uint8, pin_memory=True)
        h_reads.copy_(d_reads)
        h_off = torch.empty(R + 1, dtype=torch.int64, pin_memory=True)
        h_off.copy_(d_off)
        torch.cuda.synchronize()
        arts = [N.ART_LMER_KEYS, N.ART_LMER_VALUES, N.ART_LMER_OFFSETS, N.ART_EDGE_V1, N.ART_EDGE_V2, N.ART_EV]
        width = {N.ART_LMER_KEYS: 8, N.ART_LMER_VALUES: 4, N.ART_LMER_OFFSETS: 4, N.ART_EDGE_V1: 4, N.ART_EDGE_V2: 4,
                 N.ART_EV: 24}
        cap_items = int(max(st.distinct_lmers, st.distinct_kmers) * 1.05) + 1024
        h_out = {a: torch.empty(cap_items * width[a], dtype=torch.uint8, pin_memory=True) for a in arts}

        def e2e_step(c=ctx, out=h_out):
            s = c.run_host_ptr(h_reads.data_ptr(), h_off.data_ptr(), R, l, 0, hint)
            nb = 0
            for a in arts:
                nb += c.download_into(a, out[a].data_ptr(), out[a].numel())
            return s, nb

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        d2h = 0
        for _ in range(args.steps):
            s2, d2h = e2e_step()
        e1.record(stream)
        barrier()
        e2e_wall = 1e3 * (time.perf_counter() - t0) / args.steps
        serial_ms = max(e0.elapsed_time(e1) / args.steps, e2e_wall)  # the call returns synchronously: wall is the truth
        # Two contexts (two streams, two host threads) alternate steps, so the PCIe copies of one step
        # overlap the kernels and the opposite-direction copy of the other; every step still moves its own
        # inputs in and its own result out inside the timed region.
        ctx_b = N.Context(local_rank)
        h_out_b = {a: torch.empty(cap_items * width[a], dtype=torch.uint8, pin_memory=True) for a in arts}
        for _ in range(2):
            e2e_step(ctx_b, h_out_b)
        nsteps2 = max(2, args.steps + (args.steps & 1))
        errs = []

        def worker(c, out):
            try:
                for _ in range(nsteps2 // 2):
                    e2e_step(c, out)
            except Exception as exc:   # surfaced below: a failed step must not look like a fast one
                errs.append(exc)

        barrier()
        th = [threading.Thread(target=worker, args=(ctx, h_out)), threading.Thread(target=worker, args=(ctx_b, h_out_b))]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        piped_ms = 1e3 * (time.perf_counter() - t0) / nsteps2
        if errs:
            raise errs[0]
        ctx_b.close()
        e2e = {"ms": min(serial_ms, piped_ms), "serial_ms": serial_ms, "piped_ms": piped_ms, "h2d": int(R * L + (R + 1) * 8),
               "d2h": int(d2h)}

    # ---- reduce over ranks
    nk_local = float(st.n_kmer_windows)
    vals = torch.tensor([ms, e2e["ms"] if e2e else 0.0, kern_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([nk_local, float(launches)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_max, e2e_ms_max, kern_ms_max = [float(x) for x in vals.tolist()]
    nk_total, launches_total = [float(x) for x in tot.tolist()]

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        a_kernel, a_path = algorithmic_bytes(st, 16 if l > 32 else 8)
        achieved = a_kernel / (kern_ms_max * 1e-3) / 1e9
        traffic = None
        tfile = os.path.join(ROOT, "profiles", "count_kernel_traffic.json")
        if os.path.exists(tfile) and args.workload == DEFAULT_WORKLOAD:
            try:
                with open(tfile) as f:
                    traffic = json.load(f).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": nk_total / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
            "vs_baseline": None, "dtype": "u64" if l <= 32 else "u128", "data": "synthetic",
            "config": {
                "workload": args.workload, "genome_bp": wl["G"], "read_len": L, "coverage": wl["cov"], "err_ppm": wl["err_ppm"],
                "k": k, "reads_per_gpu": R, "bases_per_gpu": R * L,
                "parallelism": "1 GPU" if world == 1 else
                "%d GPUs: reads sharded, k-mer space partitioned by the minimizer of each vertex, one exchange of canonical l-mer keys (see dist.transport)" % world,
                "genome_bp_total": G,
                "l2": "inputs (%d MB ASCII) + table (%d MB) exceed the 126 MB L2; the table is re-initialised every step"
                      % (R * L // 10 ** 6, st.lmer_table_capacity * 12 // 10 ** 6),
                "distinct_hint": hint, "ids": "slot order (canonical-id sort not in the timed region)",
                "timing": "CUDA events on the library's stream around the K steps" if world == 1 else
                          "host clock around the K steps, barrier + device synchronize on both sides, max over ranks "
                          "(every step synchronises with the host at the count exchange, so device events see the same interval)",
            },
            "counts": {"n_kmer_windows": int(st.n_kmer_windows), "n_lmer_windows": int(st.n_lmer_windows),
                       "distinct_lmers": int(st.distinct_lmers), "distinct_kmers": int(st.distinct_kmers),
                       "edges": int(st.edge_count), "lmer_table_capacity": int(st.lmer_table_capacity),
                       "retries": int(st.retries)},
            "stage_ms": {"count_kernel": kern_ms_max, "graph": graph_ms, "step_wall": wall_ms,
                         "step_median": sorted(per_step)[len(per_step) // 2], "step_best": min(per_step)},
            "roofline": {"bound": "hbm", "kernel": ("count_compact_kernel" if l <= 32 else "wide_count_kernel") if world == 1 else
                         ("dist_count_keys_kernel" if l <= 32 else "wide_count_keys_kernel") + " (one launch per source rank)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": a_kernel,
                         "path_algorithmic_bytes": a_path,
                         "path_frac": a_path / (ms_max * 1e-3) / 1e9 / peak},
            "clocks": clocks,
            "gpu_launches": int(launches_total),
        }
        if world > 1 and info:
            line["dist"] = {k: info[k] for k in ("sent_keys", "recv_keys", "exchange_bytes", "exact_fallback", "transport", "phase_ms")}
        if e2e:
            line["e2e"] = {"value": nk_total / (e2e_ms_max * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms_max,
                           "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                           "api": "euler_pipeline_run_host + euler_pipeline_download (compressed graph)",
                           "serial_ms_per_step": e2e["serial_ms"], "pipelined_ms_per_step": e2e["piped_ms"],
                           "pipelining": "2 contexts on 2 host threads alternate steps (copies of one overlap the kernels of the other); "
                                         "serial_ms_per_step is one context, one step at a time"}
        if world == 1 and not args.no_cpu:
            nsample = min(R, 200_000)
            rate, dt, nk_s, _ = cpu_port_rate(wl, nsample)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": "first %d reads of the workload (%d k-mer windows), %.2f s" % (nsample, nk_s, dt)}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
