#!/usr/bin/env python3
"""Generate tests/golden/iod_golden.npz: the ORACLE's output (and its 1-ulp sensitivity floor) for a
fixed seeded batch.  The oracle is pinned to the reference by tests/test_oracle_kats.py; the Rust
reference itself cannot run in this image, so these vectors are oracle outputs, not reference outputs."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import binding as O  # noqa: E402
from outfit_b200 import synth  # noqa: E402
from parity_util import oracle_floor  # noqa: E402

meta = dict(T=96, n_obs=[6, 16], seed=424242, K=12, nn=3, noise_scale=1.1)
table = synth.make_ephemeris_table()
et = O.make_ephem_table(table["cheb"], table["jd_start"], table["block_days"], table["ipt"], table["emrat"])
batch = synth.make_trajectories(meta["T"], meta["n_obs"], seed=meta["seed"], table=table, max_triplets=meta["K"], n_noise=meta["nn"])
op = O.default_iod_params(n_noise_realizations=meta["nn"], max_triplets=meta["K"], noise_scale=meta["noise_scale"])
res = O.fit_full_iod(O.from_soa_batch(batch), et, op, n_threads=1)
ef, rf = oracle_floor(O, synth, batch, et, op, res)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "iod_golden.npz"),
                    results=np.frombuffer(res.tobytes(), dtype=np.uint8), elem_floor=ef, rms_floor=rf,
                    meta=json.dumps(meta), input_digest=np.array([batch["ra"].sum(), batch["dec"].sum(), batch["mjd_tt"].sum()]))
print("status histogram", np.unique(res["status"], return_counts=True))
