#!/usr/bin/env python3
"""Summarise the SASS page of one kernel of an .ncu-rep: stall reasons, opcode mix, local-memory
share and the hottest address ranges.  usage: ncu_sass.py REP KERNEL_NAME [top_n]"""
import csv, io, subprocess, sys
from collections import Counter
rep, kname = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 12
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", kname],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = Counter(); samples = inst = 0; op = Counter(); ops = Counter(); data = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break  # the page repeats the kernel: keep the first copy only
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    try:
        ns = int(r[idx["# Samples"]] or 0); ie = int(r[idx["Instructions Executed"]] or 0)
    except ValueError:
        continue
    samples += ns; inst += ie
    for s in stalls:
        try: tot[s] += int(r[idx[s]] or 0)
        except ValueError: pass
    src = r[idx["Source"]].split()
    o = (src[1] if src and src[0].startswith("@") and len(src) > 1 else (src[0] if src else "?")).split(".")[0]
    op[o] += ie; ops[o] += ns
    data.append((r[idx["Source"]].strip(), ns, ie, o))
print(f"kernel {kname}: static SASS {len(data)}, warp-instructions {inst}, samples {samples}")
print("stall reasons (share of samples):")
for s, v in tot.most_common(8):
    print(f"  {s:26s} {100*v/max(samples,1):5.1f}%")
print("opcode mix (executed share | sample share):")
for o, c in op.most_common(18):
    print(f"  {o:10s} {100*c/max(inst,1):5.1f}% | {100*ops[o]/max(samples,1):5.1f}%")
loc = sum(ie for s, ns, ie, o in data if o in ("LDL", "STL"))
locs = sum(ns for s, ns, ie, o in data if o in ("LDL", "STL"))
print(f"local-memory instructions: {100*loc/max(inst,1):.1f}% of executed, {100*locs/max(samples,1):.1f}% of samples")
# hottest 128-instruction windows
W = 128
win = [(sum(d[1] for d in data[i:i + W]), i) for i in range(0, len(data), W)]
print(f"hottest {W}-instruction windows (share of samples, first instruction index):")
for v, i in sorted(win, reverse=True)[:topn]:
    ex = sum(d[2] for d in data[i:i + W])
    print(f"  [{i:6d}] {100*v/max(samples,1):5.1f}% samples, {100*ex/max(inst,1):5.1f}% executed")
