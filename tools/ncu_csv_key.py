#!/usr/bin/env python
"""Key metrics from `ncu --page raw --csv` exports, side by side.  Usage: ncu_csv_key.py a_raw.csv b_raw.csv ..."""
import csv, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_op_read_hit_rate.pct', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active'] + \
       [f'smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio' for k in
        ('long_scoreboard', 'short_scoreboard', 'wait', 'math_pipe_throttle', 'mio_throttle', 'barrier', 'no_instruction', 'dispatch_stall', 'not_selected', 'branch_resolving')]
cols = []
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    cols.append({h: (v, u) for h, u, v in zip(hdr, units, vals)})
print(f"{'metric':86s}" + ''.join(f'{f.split("/")[-1][:22]:>24s}' for f in sys.argv[1:]))
for w in WANT:
    print(f'{w:86s}' + ''.join(f"{(c.get(w, ('-', ''))[0][:14] + ' ' + c.get(w, ('', ''))[1][:8]):>24s}" for c in cols))
