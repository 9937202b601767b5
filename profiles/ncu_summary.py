#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics + dynamic instruction mix and
stall sampling of the step kernel.   usage: python profiles/ncu_summary.py <file.ncu-rep> <warp_steps>"""
import collections
import csv
import io
import subprocess
import sys

rep, warp_steps = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print("%-70s %-12s %s" % (h, u, v))
for h, u, v in zip(hdr, units, vals):
    if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(v) > 0.05:
        print("%-70s %s" % (h.replace("smsp__average_warps_issue_stalled_", "stall/issue: "), v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
byop, samples, total, nsamp = collections.Counter(), collections.Counter(), 0, 0
for r in rows[2:]:
    try:
        ex = int(r[ix["Instructions Executed"]])
    except (ValueError, IndexError):
        continue
    parts = r[ix["Source"]].split()
    op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
    byop[op] += ex
    total += ex
    ns = int(r[ix["# Samples"]] or 0)
    samples[op] += ns
    nsamp += ns
print("warp-instructions per warp-step: %.1f" % (total / warp_steps))
for op, c in byop.most_common(22):
    print("  %-8s %7.1f   stall samples %5.1f%%" % (op, c / warp_steps, 100.0 * samples[op] / max(nsamp, 1)))
fp64 = sum(byop[o] for o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print("fp64-pipe instructions per warp-step: %.1f" % (fp64 / warp_steps))
