"""Turn an `ncu --set full` report into the small per-kernel CSV kept under profiles/ (one column per kernel of interest,
the last launch of each).  usage: python scripts/ncu_summary_csv.py report.ncu-rep out.csv [kernel-name-substring ...]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    subs = sys.argv[3:] or ["bkt_partition_kernel", "bkt_build_kernel"]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = []
    for s in subs:
        hits = [r for r in data if s in r[idx["Kernel Name"]]]
        if hits:
            cols.append(hits[-1])
    names = [r[idx["Kernel Name"]].split("(")[0].replace("void ", "") for r in cols]
    metrics = [m for m in WANT if m in idx] + sorted(h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio"))
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + names)
        for m in metrics:
            w.writerow([m, units[idx[m]]] + [r[idx[m]] for r in cols])


if __name__ == "__main__":
    main()
