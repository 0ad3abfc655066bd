#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + top stalled SASS lines).  Usage: ncu_summary.py REPORT [n_top]"""
import csv, subprocess, sys, io, re
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for w in want:
    for h, u, v in zip(hdr, units, vals):
        if h == w:
            print(f"{h:75s} {v} {u}")
stalls = [(h, float(v or 0)) for h, v in zip(hdr, vals) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("not_issued")]
tot = sum(v for _, v in stalls) or 1
print("stalls:", ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')}={100 * v / tot:.0f}%" for h, v in sorted(stalls, key=lambda t: -t[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hi]; iS = h.index("Source"); iN = h.index("# Samples")
data = [(int(r[iN] or 0), r[iS]) for r in rows[hi + 1:] if len(r) > iN]
tot = sum(d[0] for d in data) or 1
for n, s in sorted(data, reverse=True)[:ntop]:
    print(f"  {100 * n / tot:5.1f}%  {s[:100]}")
