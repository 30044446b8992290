#!/usr/bin/env python3
"""Attribute ncu warp-stall samples of one kernel to CUDA source lines, by joining the .ncu-rep
SASS page (per-instruction samples, in address order) with `nvdisasm -g` of the same build's cubin
(per-instruction file:line).  usage: ncu_lines.py REP KERNEL_SUBSTR LIB.so [top_n]"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import Counter, defaultdict
rep, kname, lib = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 50
sect = sys.argv[5] if len(sys.argv) > 5 else kname  # substring of the mangled section name (template instances)
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", kname],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
prof = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break  # the page repeats the kernel: keep the first copy only
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    try:
        prof.append((r[idx["Source"]].strip(), int(r[idx["# Samples"]] or 0), int(r[idx["Instructions Executed"]] or 0),
                     int(r[idx["Thread Instructions Executed"]] or 0)))
    except ValueError:
        pass
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# all instructions of the sections whose name contains kname, in order; device functions called by the
# kernel live in the same section after the entry (nvdisasm prints them with their own labels)
ins = []; cur = ("?", 0); take = False
for line in dis:
    m = re.match(r"\s*//-+ \.text\.(\S+)", line)
    if m:
        take = sect in m.group(1); continue
    if not take:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((m.group(2).strip(), cur))
print(f"ncu instructions {len(prof)}, nvdisasm instructions {len(ins)}")
n = min(len(prof), len(ins))
bad = sum(1 for i in range(n) if prof[i][0].split()[0:1] != ins[i][0].split()[0:1])
print(f"opcode mismatches in the first {n}: {bad}")
if bad or len(prof) != len(ins):
    print("the in-tree build is not the profiled build: no line attribution")
    sys.exit(0)
by = Counter(); ex = Counter(); th = Counter(); tot = sum(p[1] for p in prof); tex = sum(p[2] for p in prof)
for i in range(n):
    by[ins[i][1]] += prof[i][1]; ex[ins[i][1]] += prof[i][2]; th[ins[i][1]] += prof[i][3]
src_cache = {}
def src(f, l):
    for d in ("outfit_b200/csrc", "include"):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[l - 1].strip()[:100] if 0 < l <= len(L) else ""
    return ""
byfile = Counter()
for (f, l), v in by.items(): byfile[f] += v
print("samples by file:", ", ".join(f"{f} {100*v/tot:.1f}%" for f, v in byfile.most_common(8)))
for (f, l), v in by.most_common(topn):
    print(f"{100*v/tot:5.2f}% smp {100*ex[(f,l)]/tex:5.2f}% exe {th[(f,l)]/max(1,ex[(f,l)]):5.1f} lanes  {f}:{l}: {src(f, l)}")
# lanes per executed instruction by source file (where the divergence is)
fx = Counter(); ft = Counter()
for (f, l), v in ex.items(): fx[f] += v; ft[f] += th[(f, l)]
print("lanes per instruction by file:", ", ".join(f"{f} {ft[f]/max(1,fx[f]):.1f} ({100*fx[f]/tex:.0f}% exe)" for f, _ in fx.most_common(8)))
