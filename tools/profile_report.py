#!/usr/bin/env python3
"""Write the text summary of an .ncu-rep that gets committed under profiles/ (the .ncu-rep itself stays
in gpurun_out/).  usage: profile_report.py REP OUT.txt KERNEL [KERNEL ...]"""
import csv, io, os, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
kernels = sys.argv[3:]
here = os.path.dirname(os.path.abspath(__file__))
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_average_branch_targets_threads_uniform.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none summary of {os.path.basename(rep)} (tools/profile_report.py)\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        f.write(f"\n== {d.get('Kernel Name','')[:100]}\n")
        for k in KEYS:
            if k in d:
                f.write(f"  {k:72s} {d[k]:>20s} {units[hdr.index(k)]}\n")
    for k in kernels:
        f.write("\n" + "=" * 100 + "\n")
        f.write(subprocess.run([sys.executable, os.path.join(here, "ncu_sass.py"), rep, k], capture_output=True, text=True).stdout)
        lib = os.path.join(os.path.dirname(here), "outfit_b200", "liboutfit_b200.so")
        f.write("\nhottest source lines (samples joined with nvdisasm -g of the in-tree build; valid when the build matches):\n")
        f.write(subprocess.run([sys.executable, os.path.join(here, "ncu_lines.py"), rep, k, lib, "25"], capture_output=True, text=True).stdout)
print("wrote", out)
