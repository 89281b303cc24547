#!/usr/bin/env python
"""Columns of `ncu -i rep --page raw --csv` kept under profiles/: tools/ncu_extract.py raw.csv > out.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct"]
keep += sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"))
idx = [hdr.index(k) for k in keep if k in hdr]
w = csv.writer(sys.stdout)
for r in rows:
    if len(r) == len(hdr):
        w.writerow([r[i] for i in idx])
