#!/usr/bin/env python3
"""Dev script: device-resident rate of the two-body ephemeris kernels (BASELINE configs[4])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from outfit_b200 import OutfitB200, synth
n = int(os.environ.get("PERF_N", "1000000")); E = int(os.environ.get("PERF_E", "100"))
ctx = OutfitB200(0); ctx.load_ephemeris(synth.make_ephemeris_table())
kind, epoch, elem = synth.make_ephemeris_orbits(n)
tt, ut1, bf = synth.make_ephemeris_epochs(E)
dev = torch.device("cuda")
d = [torch.from_numpy(x).to(dev) for x in (kind, epoch, elem, tt, ut1)]
d_o = torch.empty(9 * E * n, dtype=torch.float64, device=dev); d_s = torch.empty(E * n, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(2): ctx.ephemeris_twobody_device(n, d[0], d[1], d[2], E, d[3], d[4], bf, d_o, d_s, stream=s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): ctx.ephemeris_twobody_device(n, d[0], d[1], d[2], E, d[3], d[4], bf, d_o, d_s, stream=s)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"ephemeris {n} x {E}: {ms:.3f} ms  {n*E/ms*1e3/1e9:.2f} G entries/s  {76.0*n*E/ms*1e3/1e9:.0f} GB/s  ok={float((d_s==0).float().mean()):.4f}")
import hashlib
hsh = hashlib.sha1()
for q in range(9):
    hsh.update(d_o[q * E * n:(q + 1) * E * n].cpu().numpy().tobytes())
hsh.update(d_s.cpu().numpy().tobytes())
print("sha1", hsh.hexdigest())
