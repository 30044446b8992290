import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import binding as O
from outfit_b200 import IODParams, OutfitB200, synth
np.set_printoptions(precision=17, linewidth=220)
table = synth.make_ephemeris_table()
et = O.make_ephem_table(table["cheb"], table["jd_start"], table["block_days"], table["ipt"], table["emrat"])
ctx = OutfitB200(0); ctx.load_ephemeris(table)
batch = synth.make_trajectories(400, (3, 9), seed=104, table=table, max_triplets=10, n_noise=2)
kw = dict(n_noise_realizations=2, max_triplets=10, noise_scale=1.0)
got = ctx.fit_full_iod(batch, IODParams.builder(**kw))
want = O.fit_full_iod(O.from_soa_batch(batch), et, O.default_iod_params(**kw), n_threads=0)
ok = (got["status"] == 0) & (want["status"] == 0)
bad = np.where(ok & (np.abs(got["epoch"] - want["epoch"]) > 1e-8))[0]
print("bad", bad)
for i in bad[:4]:
    print(i, "n_obs", batch["traj_offset"][i+1]-batch["traj_offset"][i])
    print(" gpu", got[i]); print(" cpu", want[i])
    o0, o1 = int(batch["traj_offset"][i]), int(batch["traj_offset"][i+1])
    print(" t", batch["mjd_tt"][o0:o1])
