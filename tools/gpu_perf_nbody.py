#!/usr/bin/env python3
"""Dev script: device-resident rate of the bulk N-body propagator (DOP853, frozen perturbers, state + STM)."""
import os, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from outfit_b200 import NBodyConfig, OutfitB200, planet_gm, synth
n = int(os.environ.get("PERF_N", "500000"))
ctx = OutfitB200(0)
kind, epoch, elem = synth.make_ephemeris_orbits(n, seed=1)
rng = np.random.default_rng(2)
t1 = epoch + rng.uniform(20.0, 120.0, n) * float(os.environ.get("PERF_SPAN_SCALE", "1"))
bodies = (0, 5, 6, 3, 4)
gm = np.array([planet_gm(b) for b in bodies])
radius = {0: 0.0, 3: 1.0, 4: 1.52, 5: 5.2, 6: 9.5}
pos = np.zeros((len(bodies), 3, n))
for j, b in enumerate(bodies):
    lon = rng.uniform(0, 2 * np.pi, n)
    pos[j, 0], pos[j, 1] = radius[b] * np.cos(lon), radius[b] * np.sin(lon)
dev = torch.device("cuda")
d = [torch.from_numpy(x).to(dev) for x in (kind, epoch, elem, t1, gm, np.ascontiguousarray(pos))]
d_out = torch.empty(6 * n, dtype=torch.float64, device=dev); d_stm = torch.empty(36 * n, dtype=torch.float64, device=dev)
d_st = torch.empty(n, dtype=torch.int32, device=dev); d_steps = torch.empty(n, dtype=torch.int32, device=dev)
cfg = NBodyConfig(n_perturbers=len(bodies))
s = torch.cuda.current_stream().cuda_stream
def run():
    rc = ctx._L.outfit_b200_propagate_nbody_device(ctx._h, n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), C.byref(cfg),
                                                   d[4].data_ptr(), d[5].data_ptr(), d_out.data_ptr(), d_stm.data_ptr(), d_st.data_ptr(), d_steps.data_ptr(), s)
    assert rc == 0, rc
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
steps = d_steps.cpu().numpy()
print(f"propagate_nbody n={n}, {len(bodies)} perturbers, rtol=atol=1e-12: {ms:.2f} ms  {n/ms*1e3/1e6:.2f} M orbits/s  mean steps {steps.mean():.1f}  "
      f"{steps.sum()*13/ms*1e3/1e9:.2f} G rhs evaluations/s  ok={float((d_st==0).float().mean()):.4f}")
print("sha1", hashlib.sha1(d_out.cpu().numpy().tobytes()).hexdigest())
