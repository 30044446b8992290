#!/usr/bin/env python3
"""Debug (needs a -DOUTFIT_DEBUG_FGHIST build, OUTFIT_B200_LIB=...): distribution of the Kepler Newton steps a
candidate executes in correct_kernel, and of the maximum over each warp -- the lane efficiency of the phase and what
a straggler cut-off could recover."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from outfit_b200 import IODParams, OutfitB200, RESULT_DTYPE, synth
from outfit_b200.api import load_library
T = int(os.environ.get("PERF_T", "30000")); K, nn = 30, 10
table = synth.make_ephemeris_table()
batch = synth.make_trajectories(T, 12, seed=20261018, table=table, max_triplets=K, n_noise=nn)
ctx = OutfitB200(0); ctx.load_ephemeris(table); ctx.set_pass_streams(1)
params = IODParams.builder(n_noise_realizations=nn, noise_scale=1.1, max_triplets=K)
dev = torch.device("cuda")
keys = ["traj_offset", "mjd_tt", "ra", "dec", "sigma_ra", "sigma_dec", "helio_equ", "geo_ecl", "noise_z"]
devb = {k: torch.from_numpy(batch[k].view(np.int64) if batch[k].dtype == np.uint64 else batch[k]).to(dev) for k in keys}
devb["max_obs_per_traj"] = 12
d_out = torch.zeros(T * RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
L = load_library(); L.outfit_b200_debug_fghist.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
raw = (C.c_ulonglong * 384)()
ctx.fit_full_iod_device(devb, params, d_out); torch.cuda.synchronize()
L.outfit_b200_debug_fghist(ctx._h, raw, 0)
h = np.array(raw[:], dtype=np.float64).reshape(3, 128)
lane, wmax = h[0], h[1]
s_it, s_max, n_w = h[2][0], h[2][1], h[2][2]
print(f"candidates {lane.sum():.0f} warps {n_w:.0f}  mean steps/lane {s_it / (32 * n_w):.1f}  mean warp max {s_max / n_w:.1f}  lane efficiency {s_it / (32 * s_max):.3f}")
c = np.cumsum(lane) / lane.sum(); cw = np.cumsum(wmax) / wmax.sum()
print("steps<=  lanes_cdf  warpmax_cdf")
for b in range(0, 128, 2):
    if lane[b:b + 2].sum() or wmax[b:b + 2].sum():
        print(f"{8 * (b + 2):6d}  {c[b + 1]:.4f}  {cw[b + 1]:.4f}")
