#!/usr/bin/env python3
"""Small end-to-end run of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from outfit_b200 import IODParams, OutfitB200, SolverType, synth
table = synth.make_ephemeris_table()
ctx = OutfitB200(0); ctx.load_ephemeris(table)
batch = synth.make_trajectories(300, (3, 14), seed=5, table=table, max_triplets=12, n_noise=3)
p = IODParams.builder(n_noise_realizations=3, max_triplets=12, noise_scale=1.1)
r1 = ctx.fit_full_iod(batch, p)
r2 = ctx.fit_full_iod(batch, p, use_body_fixed=True)
seeded = dict(batch); seeded["noise_z"] = None; seeded["traj_seed"] = np.arange(300, dtype=np.uint64)
r3 = ctx.fit_full_iod(seeded, p)
rv, t0, t1 = synth.make_propagation_states(5000)
o, st = ctx.propagate_universal(rv, t0, t1, SolverType(kind=2))
kind, epoch, elem = synth.make_ephemeris_orbits(700, mixed_kinds=True)
tt, ut1, bf = synth.make_ephemeris_epochs(133)
eo, es = ctx.ephemeris_twobody(kind, epoch, elem, tt, ut1, bf)
from outfit_b200 import DifferentialCorrectionConfig, EphemerisConfig, OutfitGroup
ctx.set_ephemeris_config(EphemerisConfig(aberration=2))
eo2, es2 = ctx.ephemeris_twobody(kind, epoch, elem, tt, ut1, bf)
ctx.set_ephemeris_config(EphemerisConfig())
lres, lfit = ctx.fit_lsq(batch, p, DifferentialCorrectionConfig.default(), initial_orbits=r1)
g = OutfitGroup([0, 0]); g.load_ephemeris(table)
assert g.fit_full_iod(batch, p).tobytes() == r1.tobytes()
one = ctx.fit_iod(batch, p, 7)
print("ok", (r1["status"] == 0).mean(), (r2["status"] == 0).mean(), (r3["status"] == 0).mean(), (st == 0).mean(), (es == 0).mean())
