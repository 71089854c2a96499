#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: stall-reason shares and pipe instruction counts per kernel."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_fma.sum', 'smsp__inst_executed_pipe_alu.sum',
        'smsp__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_xu.sum', 'smsp__inst_executed_op_shared_ld.sum',
        'smsp__inst_executed_op_shared_st.sum', 'smsp__inst_executed_op_global_ld.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread']
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')][:70])
    st = []
    for h, v in zip(hdr, r):
        if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
            try:
                st.append((float(v.replace(',', '')), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError:
                pass
    tot = sum(s for s, _ in st) or 1
    print('   stalls: ' + ', '.join(f"{h} {s / tot * 100:.1f}%" for s, h in sorted(st, reverse=True)[:8]))
    for k in keys:
        if k in hdr:
            print('   ', k, r[hdr.index(k)])
