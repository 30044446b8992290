#!/usr/bin/env python3
"""Summarise an .ncu-rep (run here, no GPU needed): key metrics, stall reasons, opcode mix, hottest functions."""
import csv, subprocess, sys, io
from collections import Counter, defaultdict
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
for r in rows[2:3]:
    d = dict(zip(hdr, r))
    units = dict(zip(hdr, rows[1]))
    print("kernel:", d.get("Kernel Name", "")[:60])
    for k in keys:
        if k in d:
            print(f"  {k:75s} {d[k]:>16s} {units.get(k,'')}")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = Counter(); samples = inst = 0; op = Counter(); ops = Counter(); n_static = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        ns = int(r[idx["# Samples"]] or 0); ie = int(r[idx["Instructions Executed"]] or 0)
    except ValueError:
        continue
    n_static += 1; samples += ns; inst += ie
    for s in stalls:
        try: tot[s] += int(r[idx[s]] or 0)
        except ValueError: pass
    src = r[idx["Source"]].split()
    o = (src[1] if src and src[0].startswith("@") else (src[0] if src else "?")).split(".")[0]
    op[o] += ie; ops[o] += ns
print(f"static SASS instructions: {n_static}; warp-instructions executed: {inst}; samples: {samples}")
print("stall reasons (share of samples):")
for s, v in tot.most_common(8):
    print(f"  {s:26s} {100*v/max(samples,1):5.1f}%")
print("opcode mix (executed share | sample share):")
for o, c in op.most_common(16):
    print(f"  {o:10s} {100*c/max(inst,1):5.1f}% | {100*ops[o]/max(samples,1):5.1f}%")
