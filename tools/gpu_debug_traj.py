#!/usr/bin/env python3
"""Debug: one trajectory of a synthetic batch on the GPU against the oracle and the oracle's own
answers under +-1 ulp moves of RA/Dec.  usage: gpu_debug_traj.py T n_obs seed K nn traj_index"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import binding as O
from outfit_b200 import IODParams, OutfitB200, shard, synth
T, n_obs, seed, K, nn, ti = (int(x) for x in sys.argv[1:7])
table = synth.make_ephemeris_table()
et = O.make_ephem_table(table["cheb"], table["jd_start"], table["block_days"], table["ipt"], table["emrat"])
batch = synth.make_trajectories(T, n_obs, seed=seed, table=table, max_triplets=K, n_noise=max(nn, 1))
one = shard.slice_batch(batch, ti, ti + 1)
kw = dict(n_noise_realizations=nn, max_triplets=K, noise_scale=1.1)
ctx = OutfitB200(0); ctx.load_ephemeris(table)
got = ctx.fit_full_iod(one, IODParams.builder(**kw))[0]
op = O.default_iod_params(**kw)
want = O.fit_full_iod(O.from_soa_batch(one), et, op, n_threads=1)[0]
f = ("status", "triplet_idx", "triplet_rank", "realization", "corrected", "rms", "epoch", "elem")
print("gpu   ", {k: got[k] for k in f})
print("oracle", {k: want[k] for k in f})
for sr, sd in ((np.inf, -np.inf), (-np.inf, np.inf), (np.inf, np.inf), (-np.inf, -np.inf)):
    ob = O.from_soa_batch(one); ob["ra"] = np.nextafter(ob["ra"], sr); ob["dec"] = np.nextafter(ob["dec"], sd)
    p = O.fit_full_iod(ob, et, op, n_threads=1)[0]
    print("oracle +-1ulp", {k: p[k] for k in ("triplet_idx", "realization", "rms")})
