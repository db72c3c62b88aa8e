#!/usr/bin/env python
"""Summarise an ncu report: per kernel the headline counters, stall reasons per issue, and per-phase
(barrier-delimited) instruction / stall-sample shares.  usage: ncu_summary.py file.ncu-rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "inst_executed", "sm__inst_executed.avg.per_cycle_active", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__grid_size"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:90], d.get("launch__grid_size"))
    for k in want:
        if k in d: print(f"   {k:75s} {d[k]}")
    st = sorted(((float(v), k.split('issue_stalled_')[1].split('_per_')[0]) for k, v in d.items()
                 if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio')), reverse=True)
    print("   stalls/issue: " + " ".join(f"{n}:{v:.2f}" for v, n in st[:9]))
