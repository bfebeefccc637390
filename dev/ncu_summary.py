"""Dev helper: condense `ncu -i X.ncu-rep --page raw --csv` into the columns profiles/ keeps.
Usage: python dev/ncu_summary.py raw.csv > summary.csv      (raw.csv may be .gz)"""
import csv
import gzip
import sys

KEEP = ["Kernel Name", "Block Size", "Grid Size",
        "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]

path = sys.argv[1]
f = gzip.open(path, "rt", errors="ignore") if path.endswith(".gz") else open(path, errors="ignore")
rows = list(csv.reader(f))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
H, U = rows[hdr], rows[hdr + 1]
cols = [H.index(k) for k in KEEP if k in H]
w = csv.writer(sys.stdout)
w.writerow([H[c] for c in cols])
w.writerow([U[c] for c in cols])
for r in rows[hdr + 2:]:
    if len(r) >= len(H):
        w.writerow([r[c] for c in cols])
