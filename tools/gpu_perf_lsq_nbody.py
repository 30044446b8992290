#!/usr/bin/env python3
"""Dev script: FitLSQ with PropagatorKind::NBody through the host entry, for several accepted-step budgets."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from outfit_b200 import DifferentialCorrectionConfig, IODParams, NBodyConfig, OutfitB200, planet_gm, synth
T = int(os.environ.get("PERF_T", "20000"))
table = synth.make_ephemeris_table()
ctx = OutfitB200(0); ctx.load_ephemeris(table)
batch = synth.make_trajectories(T, 12, seed=20261018, table=table, max_triplets=30, n_noise=10)
p = IODParams.builder(n_noise_realizations=10, max_triplets=30, noise_scale=1.1)
iod = ctx.fit_full_iod(batch, p)
bodies = (0, 5, 6, 3, 4)
gm = np.array([planet_gm(b) for b in bodies])
rng = np.random.default_rng(11)
pos = np.zeros((len(bodies), 3, T))
for j, rad in enumerate((0.0, 5.2, 9.5, 1.0, 1.52)):
    lon = rng.uniform(0, 2 * np.pi, T)
    pos[j, 0], pos[j, 1] = rad * np.cos(lon), rad * np.sin(lon)
cfg = DifferentialCorrectionConfig.default()
two, _ = ctx.fit_lsq(batch, p, cfg, initial_orbits=iod)
for ms in [int(x) for x in os.environ.get("PERF_BUDGETS", "0,5000,500").split(",")]:
    nb = NBodyConfig(n_perturbers=len(bodies), max_steps=ms)
    ctx.fit_lsq_nbody(batch, iod, gm, pos, cfg, nb)
    t0 = time.perf_counter()
    res, fit = ctx.fit_lsq_nbody(batch, iod, gm, pos, cfg, nb)
    dt = time.perf_counter() - t0
    ok = (res["kind"] == 1) & (two["kind"] == 1)
    print(f"max_steps={ms or 100000}: {dt*1e3:.1f} ms  {T/dt:.0f} traj/s  kinds {np.bincount(res['kind'], minlength=3).tolist()} "
          f"(two-body {np.bincount(two['kind'], minlength=3).tolist()})  max newton it {int(res['total_newton_iterations'].max())}  "
          f"median |d elem| vs two-body {np.median(np.abs(res['elem'][ok] - two['elem'][ok]).max(axis=1)):.2e}")
