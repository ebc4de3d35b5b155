#!/usr/bin/env python3
"""Print the metrics we track from an .ncu-rep (raw page) -- usage: tools/ncu_summary.py file.ncu-rep [more.ncu-rep ...]"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fmalite.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum"]
cols = {}
for f in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    cols[f] = d
names = list(cols)
allk = [k for k in KEYS if any(k in cols[n] for n in names)]
allk += sorted(k for k in cols[names[0]] if "issue_stalled" in k and k.endswith("per_issue_active.ratio") or ("issue_stalled" in k and "_per_warp_active" in k and False))
for k in allk:
    print("%-90s" % k[:90], "  ".join("%16s" % (cols[n].get(k, ("-", ""))[0]) for n in names))
