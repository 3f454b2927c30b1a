#!/usr/bin/env python
"""ncu -i X.ncu-rep --page raw --csv  ->  the per-launch extract committed under profiles/ (one row per
launch, header + units rows), restricted to the metrics DESIGN.md and bench.py quote."""
import csv, subprocess, sys
KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
idx = [hdr.index(k) for k in KEEP if k in hdr]
w = csv.writer(open(sys.argv[2], "w", newline=""))
w.writerow([hdr[i] for i in idx])
w.writerow([rows[1][i] if hdr[i] != "Kernel Name" else "" for i in idx])
for r in rows[2:]:
    w.writerow([r[i] for i in idx])
print(len(rows) - 2, "launches ->", sys.argv[2])
