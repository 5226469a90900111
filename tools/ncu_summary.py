#!/usr/bin/env python3
"""Summarise an `ncu --page raw --csv` export: one block of key metrics per profiled launch."""
import csv
import sys

WANT = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.avg.per_second', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {w: hdr.index(w) for w in WANT if w in hdr}
for r in rows[2:]:
    print('---')
    for w, i in idx.items():
        v = r[i][:100]
        print(f"  {w:70s} {v} {units[i]}")
if len(sys.argv) > 2:
    pat = sys.argv[2]
    for i, h in enumerate(hdr):
        if pat in h:
            print(h, [r[i] for r in rows[2:]])
