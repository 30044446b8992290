#!/usr/bin/env python3
"""Debug (needs a -DOUTFIT_DEBUG_STRAGGLERS build, OUTFIT_B200_LIB=...): longest-running thread of each
f-g stage kernel and the total thread time, to size the straggler tail."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from outfit_b200 import IODParams, OutfitB200, RESULT_DTYPE, synth
from outfit_b200.api import load_library
T = int(os.environ.get("PERF_T", "30000")); K, nn = 30, 10
table = synth.make_ephemeris_table()
batch = synth.make_trajectories(T, 12, seed=20261018, table=table, max_triplets=K, n_noise=nn)
ctx = OutfitB200(0); ctx.load_ephemeris(table); ctx.set_work_counters(False); ctx.set_pass_streams(1)
params = IODParams.builder(n_noise_realizations=nn, noise_scale=1.1, max_triplets=K)
dev = torch.device("cuda")
keys = ["traj_offset", "mjd_tt", "ra", "dec", "sigma_ra", "sigma_dec", "helio_equ", "geo_ecl", "noise_z"]
devb = {k: torch.from_numpy(batch[k].view(np.int64) if batch[k].dtype == np.uint64 else batch[k]).to(dev) for k in keys}
devb["max_obs_per_traj"] = 12
d_out = torch.zeros(T * RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
ctx.fit_full_iod_device(devb, params, d_out); torch.cuda.synchronize()
L = load_library(); L.outfit_b200_debug_counters.argtypes = [C.c_void_p, C.c_void_p]
raw = (C.c_ulonglong * 32)(); L.outfit_b200_debug_counters(ctx._h, raw)
clk = 1.965e9
for s in range(1):
    v = raw[1 + 20 + s]; tot = raw[1 + 24 + s]
    print(f"stage {s}: slowest thread {(v >> 26) * 256 / clk * 1e3:.3f} ms (cid {v & 0x3ffffff}), sum of thread time {tot * 256 / clk:.2f} thread-s")
print(ctx.last_iod_phase_ms())
