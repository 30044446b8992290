#!/usr/bin/env python3
"""Dev script: device-resident rate of the bulk propagate_universal kernel (BASELINE configs[1])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from outfit_b200 import OutfitB200, SolverType, synth
n = int(os.environ.get("PERF_N", "10000000"))
ctx = OutfitB200(0)
rv, t0, t1 = synth.make_propagation_states(n)
dev = torch.device("cuda")
d_rv, d_t0, d_t1 = (torch.from_numpy(x).to(dev) for x in (rv, t0, t1))
d_o = torch.empty(11 * n, dtype=torch.float64, device=dev); d_s = torch.empty(n, dtype=torch.int32, device=dev)
st = SolverType(kind=2)
s = torch.cuda.current_stream().cuda_stream
for _ in range(2): ctx.propagate_universal_device(n, d_rv, d_t0, d_t1, d_o, d_s, st, stream=s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ctx.propagate_universal_device(n, d_rv, d_t0, d_t1, d_o, d_s, st, stream=s)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"propagate_universal n={n}: {ms:.3f} ms  {n/ms*1e3/1e9:.2f} G/s  {156.0*n/ms*1e3/1e9:.0f} GB/s  ok={float((d_s==0).float().mean()):.4f}")
import hashlib
print("sha1", hashlib.sha1(d_o.cpu().numpy().tobytes() + d_s.cpu().numpy().tobytes()).hexdigest())
