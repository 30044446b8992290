"""Device-resident and host-buffer timing of FitLSQ (differential orbit correction) on one GPU."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from outfit_b200 import (DifferentialCorrectionConfig, IODParams, LSQ_RESULT_DTYPE, OBS_FIT_DTYPE, OutfitB200,  # noqa: E402
                         synth)

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
table = synth.make_ephemeris_table()
ctx = OutfitB200(0)
ctx.load_ephemeris(table)
batch = synth.make_trajectories(T, 12, seed=20261018, table=table, max_triplets=30, n_noise=10)
p = IODParams.builder(n_noise_realizations=10, max_triplets=30, noise_scale=1.1)
cfg = DifferentialCorrectionConfig.default()
iod = ctx.fit_full_iod(batch, p)
dev = {k: torch.from_numpy(batch[k]).cuda() for k in ("mjd_tt", "ra", "dec", "sigma_ra", "sigma_dec", "helio_equ", "geo_ecl")}
dev["traj_offset"] = torch.from_numpy(batch["traj_offset"].astype(np.int64)).cuda()
d_iod = torch.from_numpy(iod.view(np.uint8).reshape(len(iod), -1)).cuda()
n = len(batch["mjd_tt"])
d_out = torch.zeros((T, LSQ_RESULT_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
d_fit = torch.zeros((n, OBS_FIT_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ctx.fit_lsq_device(dev, cfg, d_iod, d_out, d_fit, stream=st)
torch.cuda.synchronize()
ms = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ctx.fit_lsq_device(dev, cfg, d_iod, d_out, d_fit, stream=st)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
res = d_out.cpu().numpy().view(LSQ_RESULT_DTYPE).reshape(-1)
t0 = time.perf_counter()
hres, hfit = ctx.fit_lsq(batch, p, cfg, initial_orbits=iod)
host_s = time.perf_counter() - t0
assert hres.tobytes() == res.tobytes()
ok = res["kind"] == 1
import hashlib
print(json.dumps({"T": T, "sha1": hashlib.sha1(res.tobytes() + d_fit.cpu().numpy().tobytes()).hexdigest(), "device_ms": ms, "device_traj_per_s": T / (min(ms) * 1e-3), "host_entry_s": host_s,
                  "host_traj_per_s": T / host_s, "kinds": np.bincount(res["kind"], minlength=3).tolist(),
                  "newton_iterations_total": int(res["total_newton_iterations"].sum()),
                  "rms_median_corrected": float(np.median(res["normalised_rms"][ok]))}))
