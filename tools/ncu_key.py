#!/usr/bin/env python
"""Key metrics of one or more .ncu-rep captures side by side (last kernel of each).  Usage: ncu_key.py a.ncu-rep b.ncu-rep"""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__t_sector_op_read_hit_rate.pct', 'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active'] + \
       [f'smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio' for k in
        ('long_scoreboard', 'short_scoreboard', 'wait', 'math_pipe_throttle', 'mio_throttle', 'lg_throttle', 'barrier', 'membar', 'no_instruction',
         'dispatch_stall', 'not_selected', 'branch_resolving', 'sleeping', 'imc_miss', 'drain', 'tex_throttle')]
cols = []
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    cols.append({h: (v, u) for h, u, v in zip(hdr, units, vals)})
print(f"{'metric':88s}" + ''.join(f'{r[-28:]:>30s}' for r in sys.argv[1:]))
for w in WANT:
    print(f'{w:88s}' + ''.join(f"{(c.get(w, ('-', ''))[0] + ' ' + c.get(w, ('', ''))[1]):>30s}" for c in cols))
