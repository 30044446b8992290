#!/usr/bin/env python3
"""Static look at one kernel of the built library (no GPU): instruction count, opcode histogram and
the source lines that own local-memory (spill / stack) instructions.  usage: sass_locals.py KERNEL [LIB]"""
import os, re, subprocess, sys, tempfile
from collections import Counter
kname = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "outfit_b200", "liboutfit_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
take = False; cur = ("?", 0); ops = Counter(); loc = Counter(); n = 0; secs = []
for line in dis:
    m = re.match(r"\s*//-+ \.text\.(\S+)", line)
    if m:
        take = kname in m.group(1)
        if take: secs.append(m.group(1))
        continue
    if not take: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        t = m.group(2).split()
        o = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[o] += 1; n += 1
        if o in ("LDL", "STL"): loc[cur] += 1
print("sections:", secs)
print("instructions:", n, " FP64:", sum(ops[o] for o in ("DFMA", "DMUL", "DADD", "DSETP")), " local:", ops["LDL"] + ops["STL"], " calls:", ops["CALL"])
print("top opcodes:", ", ".join(f"{o} {c}" for o, c in ops.most_common(14)))
print("local-memory instructions by source line:")
for (f, l), c in loc.most_common(25): print(f"  {c:4d}  {f}:{l}")
