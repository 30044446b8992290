#!/usr/bin/env python3
"""Generate tests/golden/lsq_golden.npz: the ORACLE's differential-correction output (records, per-observation
fit data, 1-ulp sensitivity floors) for the batch of iod_golden.npz, started from that fixture's IOD records,
with one 40-sigma outlier per third trajectory so the rejection step is on the path.  Oracle outputs, not
reference outputs: the Rust reference cannot run in this image (oracle pins: tests/test_lsq_oracle.py)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import binding as O  # noqa: E402
from outfit_b200 import synth  # noqa: E402
from parity_util import oracle_lsq_floor  # noqa: E402


def golden_inputs(g_iod):
    """(batch, iod records) of the fixture: the IOD golden's batch + a deterministic outlier pattern."""
    meta = json.loads(str(g_iod["meta"]))
    table = synth.make_ephemeris_table()
    batch = synth.make_trajectories(meta["T"], meta["n_obs"], seed=meta["seed"], table=table, max_triplets=meta["K"],
                                    n_noise=meta["nn"])
    iod = np.frombuffer(g_iod["results"].tobytes(), dtype=O.IOD_RESULT_DTYPE).copy()
    off = batch["traj_offset"].astype(np.int64)
    for t in range(0, meta["T"], 3):
        n = off[t + 1] - off[t]
        i = off[t] + (7 * t) % n
        batch["dec"][i] += 40.0 * batch["sigma_dec"][i]
    return table, batch, iod


if __name__ == "__main__":
    g = np.load(os.path.join(HERE, "iod_golden.npz"))
    table, batch, iod = golden_inputs(g)
    et = O.make_ephem_table(table["cheb"], table["jd_start"], table["block_days"], table["ipt"], table["emrat"])
    ob = O.from_soa_batch(batch)
    cfg = O.default_lsq_config()
    res, fit = O.fit_lsq(ob, et, cfg, iod, n_threads=1)
    floors, unstable = oracle_lsq_floor(O, ob, et, cfg, iod, res, fit)
    np.savez_compressed(os.path.join(HERE, "lsq_golden.npz"), results=np.frombuffer(res.tobytes(), dtype=np.uint8),
                        fit=np.frombuffer(fit.tobytes(), dtype=np.uint8), floors=np.stack(floors), unstable=unstable,
                        input_digest=np.array([batch["ra"].sum(), batch["dec"].sum(), batch["mjd_tt"].sum()]))
    print("kinds", np.bincount(res["kind"], minlength=3), "rejected observations", int((fit["selection"] == 1).sum()),
          "unstable", int(unstable.sum()))
