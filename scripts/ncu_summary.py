#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: the metrics that decide a memory-bound kernel."""
import csv
import sys

KEEP = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__waves_per_multiprocessor',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum', 'l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum']
STALL = 'smsp__average_warp'


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    stalls = []
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP:
            print(f'{h} [{u}] = {v}')
        if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct'):
            try:
                stalls.append((float(v), h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', '')))
            except ValueError:
                pass
    for v, h in sorted(stalls, reverse=True)[:8]:
        print(f'  stall {h}: {v:.1f}%')


if __name__ == '__main__':
    main(sys.argv[1])
