import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import binding as O
from outfit_b200 import IODParams, OutfitB200, synth
table = synth.make_ephemeris_table()
et = O.make_ephem_table(table["cheb"], table["jd_start"], table["block_days"], table["ipt"], table["emrat"])
ctx = OutfitB200(0); ctx.load_ephemeris(table)
np.set_printoptions(precision=15, linewidth=200)
for K, nn in [(30, 0), (30, 1), (10, 3), (11, 2), (30, 10)]:
    batch = synth.make_trajectories(64, 12, seed=2, table=table, max_triplets=K, n_noise=max(nn, 1))
    kw = dict(n_noise_realizations=nn, max_triplets=K, noise_scale=1.1)
    got = ctx.fit_full_iod(batch, IODParams.builder(**kw))
    want = O.fit_full_iod(O.from_soa_batch(batch), et, O.default_iod_params(**kw), n_threads=0)
    ok = (got["status"] == 0) & (want["status"] == 0)
    d = np.abs(got["elem"] - want["elem"]).max(axis=1)
    bad = np.where(ok & (d > 1e-6))[0]
    print(f"K={K} nn={nn} ncand={K*(nn+1)} bad={len(bad)}/{ok.sum()}")
    for i in bad[:3]:
        c = got["triplet_rank"][i] * (nn + 1) + got["realization"][i]
        print("  traj", i, "cand", c, "lane", c % 32, "chunk", c // 32, "rank/real", got["triplet_rank"][i], got["realization"][i], want["triplet_rank"][i], want["realization"][i])
        print("   gpu   ", got["epoch"][i], got["elem"][i], got["rms"][i], got["corrected"][i])
        print("   oracle", want["epoch"][i], want["elem"][i], want["rms"][i], want["corrected"][i])
