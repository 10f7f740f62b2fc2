#!/usr/bin/env python
"""Summarise an ncu report (raw page CSV) for the pipeline kernel: usage: read_ncu.py file.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'lts__t_sector_hit_rate.pct', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max',
        'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_lsu.sum']
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:70])
    for w in want:
        if w in hdr:
            print(f'  {w} = {r[hdr.index(w)]} {units[hdr.index(w)]}')
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
            try:
                stalls.append((float(r[i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    for v, n in sorted(stalls, reverse=True)[:8]:
        print(f'    stall {n}: {v:.2f} warps/issue')
