"""Diagnostic: LSQ GPU vs oracle for the one-iteration configuration."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import binding as O
from outfit_b200 import DifferentialCorrectionConfig, OutfitB200, RESULT_DTYPE, synth
from parity_util import _lsq_errs, lsq_int_mismatch, oracle_lsq_floor
O.build()
table = synth.make_ephemeris_table()
ctx = OutfitB200(0); ctx.load_ephemeris(table)
et = O.make_ephem_table(table["cheb"], table["jd_start"], table["block_days"], table["ipt"], table["emrat"])
batch = synth.make_trajectories(600, 12, seed=305, table=table, max_triplets=10, n_noise=1)
iod = O.fit_full_iod(O.from_soa_batch(batch), et, O.default_iod_params(n_noise_realizations=0, max_triplets=10), n_threads=0)
ob = O.from_soa_batch(batch)
kw = dict(max_newton_iterations=1, eccentricity_limit=0.3, rms_divergence_ratio=1.05)
cfg = O.default_lsq_config(**kw)
want, wfit = O.fit_lsq(ob, et, cfg, iod)
got, gfit = ctx.fit_lsq(batch, None, DifferentialCorrectionConfig.default(**kw), initial_orbits=iod.view(RESULT_DTYPE))
fl, un = oracle_lsq_floor(O, ob, et, cfg, iod, want, wfit)
off = ob["traj_offset"]
mism = lsq_int_mismatch(got, want, gfit, wfit, off)
ee, er, ec, eo = _lsq_errs(got, want, gfit, wfit, off, (ob["sigma_ra"], ob["sigma_dec"]))
ok = ~mism & (want["kind"] == 1)
bad = ok & (ee > np.maximum(1e-10, 256 * fl[0]))
print("ok", ok.sum(), "bad", bad.sum())
np.set_printoptions(linewidth=200, precision=6)
for t in np.argwhere(bad).ravel()[:12]:
    print(t, "iod rms", iod["rms"][t], "rms g/w", got["normalised_rms"][t], want["normalised_rms"][t], "floor", fl[0][t], "err", ee[t])
    print("   got ", got["elem"][t]); print("   want", want["elem"][t]); print("   diff", got["elem"][t] - want["elem"][t])
    print("   sigma", want["sigma"][t])
good = ok & ~bad
print("good err median", np.median(ee[good]), "max", ee[good].max())
