"""Transpose the raw page of an ncu --set full report into the small per-kernel summary kept under profiles/.
usage: python scripts/ncu_summary.py <report.ncu-rep> <out.csv> [note]"""
import csv
import subprocess
import sys

METRICS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed sm__throughput.avg.pct_of_peak_sustained_elapsed
smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active sm__warps_active.avg.pct_of_peak_sustained_active
smsp__thread_inst_executed_per_inst_executed.ratio l1tex__t_sector_hit_rate.pct lts__t_sector_hit_rate.pct
launch__registers_per_thread launch__grid_size launch__block_size
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio""".split()

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else "ncu --set full"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
head, units, data = rows[0], rows[1], rows[2:]
kn = head.index("Kernel Name")
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [f"{r[kn]} ({note})" for r in data])
    for m in METRICS:
        if m in head:
            c = head.index(m)
            w.writerow([m, units[c]] + [r[c] for r in data])
print(open(out).read())
