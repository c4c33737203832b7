#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py file.ncu-rep [--stalls]"""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_issued.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'launch__waves_per_multiprocessor', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio']
def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for r in data:
        print('==', r[hdr.index('Kernel Name')][:110])
        for k in KEYS:
            if k in hdr:
                print(f'  {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}')
        if '--stalls' in sys.argv:
            for i, k in enumerate(hdr):
                if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio'):
                    print(f'  {k[34:-23]:40s} {r[i]}')
main()
