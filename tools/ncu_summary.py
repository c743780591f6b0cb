#!/usr/bin/env python3
"""Selected raw metrics of every launch in an ncu report, as text (for profiles/).

  python tools/ncu_summary.py report.ncu-rep > profiles/<name>_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__inst_issued.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum',
    'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum',
    'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum',
    'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for k, r in enumerate(rows[2:]):
        print('== launch %d: %s  grid %s block %s' % (k, r[hdr.index('Kernel Name')], r[hdr.index('Grid Size')], r[hdr.index('Block Size')]))
        for i, h in enumerate(hdr):
            if h in WANT or ('issue_stalled' in h and h.endswith('per_issue_active.ratio')):
                print('   %-90s %-14s %s' % (h, units[i], r[i]))


if __name__ == '__main__':
    main()
