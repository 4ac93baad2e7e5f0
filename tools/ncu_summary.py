"""Summaries of ncu outputs for profiles/:
  python tools/ncu_summary.py launches <launches.csv>          per-kernel launch times
  python tools/ncu_summary.py full <report.ncu-rep> <out.csv>   selected raw metrics per kernel"""
import csv
import subprocess
import sys
from collections import defaultdict

WANT = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_uniform.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = defaultdict(list)
    for r in rows[1:]:
        try:
            agg[r[ki].split('(')[0][:60]].append(float(r[vi].replace(',', '')))
        except ValueError:
            pass
    print(f"{'kernel':62s} {'n':>4s} {'mean ms':>9s} {'min ms':>9s}")
    for k, v in agg.items():
        print(f"{k:62s} {len(v):4d} {sum(v) / len(v) / 1e6:9.4f} {min(v) / 1e6:9.4f}")


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "wt", newline="") as f:
        w = csv.writer(f)
        cols = [hdr.index(m) for m in WANT if m in hdr]
        w.writerow([hdr[c] for c in cols])
        w.writerow([units[c] for c in cols])
        for r in rows[2:]:
            w.writerow([r[c][:90] for c in cols])
    print(open(out).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3])
