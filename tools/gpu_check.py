#!/usr/bin/env python3
"""Dev script (run under gpurun): GPU vs oracle parity statistics + rough timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import binding as O
from outfit_b200 import IODParams, OutfitB200, SolverType, synth

def elem_err(a, b):
    """a, e relative; i, node, argp, anomaly as wrapped absolute angle differences (rad)."""
    d = np.abs(a - b)
    out = np.empty_like(d)
    out[:, 0:2] = d[:, 0:2] / np.maximum(np.abs(b[:, 0:2]), 1e-300)
    ang = np.abs((a[:, 2:6] - b[:, 2:6] + np.pi) % (2 * np.pi) - np.pi)
    out[:, 2:6] = ang
    return out.max(axis=1)

def pct(x):
    return " ".join(f"p{q}={np.percentile(x, q):.2e}" for q in (50, 90, 99, 100)) if len(x) else "-"

def compare(got, want, label):
    st_eq = got["status"] == want["status"]
    ok = (want["status"] == 0) & (got["status"] == 0)
    sel = (got["triplet_idx"][ok] == want["triplet_idx"][ok]).all(axis=1) & (got["realization"][ok] == want["realization"][ok])
    rel = np.abs(got["elem"][ok] - want["elem"][ok]) / np.maximum(np.abs(want["elem"][ok]), 1e-300)
    relr = np.abs(got["rms"][ok] - want["rms"][ok]) / want["rms"][ok]
    print(f"[{label}] n={len(got)} status_eq={st_eq.mean():.4f} ok={ok.sum()} same_selection={sel.mean():.4f} "
          f"elem_rel_max={rel[sel].max() if sel.any() else -1:.3e} rms_rel_max={relr[sel].max() if sel.any() else -1:.3e} "
          f"kind_eq={(got['element_kind'][ok]==want['element_kind'][ok]).mean():.4f} corr_eq={(got['corrected'][ok]==want['corrected'][ok]).mean():.4f}")
    ee = elem_err(got["elem"][ok][sel], want["elem"][ok][sel])
    print("   elem err:", pct(ee), "| rms rel:", pct(relr[sel]), "| epoch abs:", pct(np.abs(got["epoch"][ok][sel]-want["epoch"][ok][sel])))
    bad = np.where(~st_eq)[0][:5]
    for i in bad:
        print("   status mismatch traj", i, "gpu", got["status"][i], got["cause"][i], "oracle", want["status"][i], want["cause"][i])
    idx = np.where(ok)[0][~sel][:5]
    for i in idx:
        print("   selection mismatch traj", i, "gpu", got["triplet_idx"][i], got["realization"][i], got["rms"][i],
              "oracle", want["triplet_idx"][i], want["realization"][i], want["rms"][i])
    fail = want["status"] != 0
    if fail.any():
        print("   failures: cause_eq", (got["cause"][fail] == want["cause"][fail]).mean(), "attempts_eq",
              (got["attempts"][fail] == want["attempts"][fail]).mean())
    return st_eq.all() and sel.all()

table = synth.make_ephemeris_table()
et = O.make_ephem_table(table["cheb"], table["jd_start"], table["block_days"], table["ipt"], table["emrat"])
ctx = OutfitB200(0)
ctx.load_ephemeris(table)
print("fp64 peak TFLOP/s:", ctx.measure_fp64_peak() / 1e12)

for (T, nobs, K, nn, seed) in [(512, 12, 10, 0, 1), (512, 12, 30, 10, 2), (256, (8, 30), 30, 10, 3), (256, (3, 9), 10, 2, 4)]:
    batch = synth.make_trajectories(T, nobs, seed=seed, table=table, max_triplets=K, n_noise=max(nn, 1))
    kw = dict(n_noise_realizations=nn, max_triplets=K, noise_scale=1.1, max_obs_for_triplets=100)
    params = IODParams.builder(**kw)
    t0 = time.time(); got = ctx.fit_full_iod(batch, params); t1 = time.time()
    got = ctx.fit_full_iod(batch, params); t2 = time.time()
    want = O.fit_full_iod(O.from_soa_batch(batch), et, O.default_iod_params(**kw), n_threads=0); t3 = time.time()
    print(f"T={T} nobs={nobs} K={K} nn={nn}: gpu first {t1-t0:.3f}s second {t2-t1:.3f}s ({T/(t2-t1):.0f} traj/s) cpu {t3-t2:.3f}s ({T/(t3-t2):.0f} traj/s)")
    compare(got, want, f"K{K}n{nn}")
    # conditioning floor: the oracle against itself with RA/Dec moved by one ulp
    ob = O.from_soa_batch(batch)
    ob["ra"] = np.nextafter(ob["ra"], np.inf); ob["dec"] = np.nextafter(ob["dec"], -np.inf)
    pert = O.fit_full_iod(ob, et, O.default_iod_params(**kw), n_threads=0)
    print("   -- oracle vs oracle(+1ulp inputs):")
    compare(pert, want, "floor")
    print("   counters", ctx.last_iod_counters())

# body-fixed path (on-device pvobs) vs oracle pvobs
import ctypes as C
batch = synth.make_trajectories(64, 12, seed=9, table=table, max_triplets=10, n_noise=1)
n = batch["mjd_tt"].shape[0]
geo = np.zeros((3, n)); hel = np.zeros((3, n))
for i in range(n):
    dx = O.D3(); dv = O.D3(); h = O.D3()
    O.lib().oo_pvobs(batch["mjd_tt"][i], batch["mjd_ut1"][i], O.d3(batch["body_fixed"][:, i]), O.d3([0, 0, 0]), dx, dv)
    O.lib().oo_helio_position(C.byref(et), batch["mjd_tt"][i], dx, h)
    geo[:, i] = list(dx); hel[:, i] = list(h)
import torch
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
g_geo = torch.zeros(3, n, dtype=torch.float64, device="cuda"); g_hel = torch.zeros_like(g_geo)
ctx.observer_cache_device(n, d(batch["mjd_tt"]), d(batch["mjd_ut1"]), d(batch["body_fixed"]), g_geo, g_hel)
torch.cuda.synchronize()
print("pvobs geo max abs err (AU):", np.abs(g_geo.cpu().numpy() - geo).max(), "rel to |r|:", np.abs(g_geo.cpu().numpy() - geo).max() / 4.26e-5,
      " helio max abs err:", np.abs(g_hel.cpu().numpy() - hel).max())

# bulk propagation
rv, t0a, t1a = synth.make_propagation_states(200000, seed=5)
st = SolverType(kind=2)
t0 = time.time(); out, status = ctx.propagate_universal(rv, t0a, t1a, st); t1 = time.time()
want, wst = O.propagate_universal_batch(rv, t0a, t1a, 2, st.convergency, 0); t2 = time.time()
ok = (status == 0) & (wst == 0)
den = np.maximum(np.abs(want[:, ok]), 1e-12)
rel = np.abs(out[:, ok] - want[:, ok]) / den
print(f"propagate_universal n=200000 gpu(e2e) {t1-t0:.3f}s cpu {t2-t1:.3f}s status_eq={(status==wst).mean():.5f} ok={ok.mean():.4f} "
      f"rel_max r1={rel[0:3].max():.3e} v1={rel[3:6].max():.3e} fg={rel[6:10].max():.3e} psi={rel[10].max():.3e}")
print("status histogram gpu", np.unique(status, return_counts=True), "cpu", np.unique(wst, return_counts=True))
