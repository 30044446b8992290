#!/usr/bin/env python3
"""Write profiles/ncu_latest.json from an `ncu --set full` capture of the CURRENT kernels.

The bench line quotes ncu evidence (DRAM bytes per launch, FP64 pipe %, issue %, lanes per instruction) for the
kernel it times.  That evidence is only valid for the binary that was profiled, so the JSON records the sha256
of the kernel sources the capture was taken from (written on the GPU box next to the .ncu-rep by the capture
script) and bench.py refuses the file when the tree's sources hash differently.

usage: ncu_latest.py REP.ncu-rep SOURCE_SHA.txt WORKLOAD [OUT.json]
"""
import csv
import io
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M = {"duration_ns": "gpu__time_duration.sum", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
     "fp64_pipe_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
     "fp64_inst_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
     "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "threads_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
     "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
     "registers": "launch__registers_per_thread", "grid": "launch__grid_size", "block": "launch__block_size",
     "local_ld_sectors": "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "local_st_sectors": "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
     "inst_executed": "smsp__inst_executed.sum"}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "usecond": 1e3, "msecond": 1e6,
              "nsecond": 1.0, "second": 1e9}


def main():
    rep, sha_file, workload = sys.argv[1], sys.argv[2], sys.argv[3]
    out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "ncu_latest.json")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    kernels = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "")
        short = name.split("<")[0].split("(")[0].split("::")[-1].replace("void ", "").strip()
        k = {}
        for key, metric in M.items():
            if metric in d and d[metric] not in ("", "n/a"):
                v = float(d[metric].replace(",", ""))
                v *= UNIT_SCALE.get(units[hdr.index(metric)], 1.0)
                k[key] = v
        if "dram_read" in k and "dram_write" in k:
            k["dram_bytes"] = k["dram_read"] + k["dram_write"]
        if "duration_ns" in k:
            k["duration_ms_under_ncu"] = k.pop("duration_ns") / 1e6
        k["kernel_name"] = name[:120]
        kernels.setdefault(short, k)  # first captured launch of each kernel
    sha = open(sha_file).read().split()[0]
    git = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    doc = {"what": "ncu --set full --clock-control none, one launch per kernel at the bench's workload (single pass: OUTFIT_B200_STREAMS=1)",
           "workload": workload, "kernel_source_sha256": sha, "git": git, "captured": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
           "report": os.path.basename(rep), "kernels": kernels,
           "note": "durations under ncu are cold-cache and serialised: use them for shares only; bench.py measures the live durations"}
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", out, list(kernels))


if __name__ == "__main__":
    main()
