"""Helper (not a test): summarise an ncu launch-list CSV or raw-page CSV into readable text."""
import collections
import csv
import subprocess
import sys


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(u, 1.0)
        a = agg[row["Kernel Name"][:70]]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"{'kernel':70s} {'n':>5s} {'total_us':>11s} {'avg_us':>9s} {'share':>6s}"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k:70s} {v[0]:5d} {v[1] / 1e3:11.1f} {v[1] / v[0] / 1e3:9.1f} {v[1] / tot * 100:5.1f}%")
    return "\n".join(out)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_shared_loads", "sass__inst_executed_global_loads",
        "launch__shared_mem_per_block_dynamic", "sm__maximum_warps_per_active_cycle_pct"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    out = []
    for i, h in enumerate(hdr):
        if h in WANT or h in ("Kernel Name",):
            out.append(f"{h:90s} {units[i]:14s} {[v[i] for v in vals]}")
    return "\n".join(out)


if __name__ == "__main__":
    print(launches(sys.argv[1]) if sys.argv[1].endswith(".csv") else raw(sys.argv[1]))
