#!/usr/bin/env python3
"""Per-kernel registers / stack / spills from `nvcc -Xptxas -v` output (outfit_b200/csrc/ptxas.log)."""
import os, re, subprocess, sys
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "outfit_b200", "csrc", "ptxas.log")
txt = open(path).read()
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(.*)", txt):
    name = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
    smem = re.search(r"(\d+) bytes smem", m.group(6))
    print(f"{name:48s} regs {int(m.group(5)):3d}  stack {int(m.group(2)):5d} B  spill st/ld {int(m.group(3)):4d}/{int(m.group(4)):4d} B  smem {smem.group(1) if smem else 0}")
