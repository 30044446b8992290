#!/usr/bin/env python3
"""The reference's quick start (README / examples/run_full_iod.rs: one MPC 80-column file -> Gauss IOD)
on the B200 path, followed by the differential correction of the orbit found.

    python examples/run_full_iod.py [FILE.obs | FILE.xml | FILE.parquet] [--eop2 latest_eop2.long] [--de DE_FILE] [--no-lsq]

Without a file the committed fixture of the reference's own quick-start input is used
(tests/golden/config1_2015AB.json: the 37 observations of 2015 AB).  Without --de a synthetic DE440-shaped
Chebyshev table stands in for the JPL ephemeris (this image has no DE440), without --eop2 UT1 = UTC.
Needs the compiled library and a CUDA device: there is no CPU fallback.  --dry-run stops after the host
side (reader -> batch) and prints what would be sent to the GPU.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from outfit_b200 import ades, elements, mpc80, synth  # noqa: E402
from outfit_b200.ut1 import Ut1Table  # noqa: E402


def load_trajectories(path):
    if path is None:
        d = json.load(open(os.path.join(ROOT, "tests", "golden", "config1_2015AB.json")))
        return {d["designation"]: d["records"]}
    if path.endswith(".parquet"):
        from outfit_b200 import tabular
        return tabular.parse(path)  # default schema: traj_id, jd (UTC), ra / dec (deg), obscode
    text = open(path).read()
    if path.endswith(".xml"):
        return ades.parse(text)
    return mpc80.parse(text, single_trajectory=True)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("file", nargs="?")
    ap.add_argument("--eop2", help="JPL latest_eop2.long (UT1)")
    ap.add_argument("--de", help="JPL DE binary file (e.g. linux_p1550p2650.440)")
    ap.add_argument("--sigma-arcsec", type=float, default=0.5)
    ap.add_argument("--max-triplets", type=int, default=30)
    ap.add_argument("--noise-realizations", type=int, default=10)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--no-lsq", action="store_true")
    ap.add_argument("--dry-run", action="store_true")
    args = ap.parse_args()

    traj = load_trajectories(args.file)
    ut1 = Ut1Table.from_file(args.eop2) if args.eop2 else None
    ids, batch = mpc80.to_batch(traj, sigma_arcsec=args.sigma_arcsec, ut1_table=ut1)
    t0, t1 = float(batch["mjd_tt"].min()), float(batch["mjd_tt"].max())
    print(f"{len(ids)} trajectory(ies), {len(batch['mjd_tt'])} observations, MJD(TT) {t0:.3f} .. {t1:.3f}")
    if args.de:
        from outfit_b200 import de_reader
        table = de_reader.read_de_binary(args.de)  # the whole file: ~100 MB for DE440, copied to the GPU once
    else:
        table = synth.make_ephemeris_table(mjd_start=32.0 * np.floor((t0 - 64.0) / 32.0), n_blocks=int((t1 - t0) / 32.0) + 6)
    if args.dry_run:
        print("dry run: batch", {k: (v.shape if hasattr(v, "shape") else v) for k, v in batch.items()})
        return 0

    from outfit_b200 import DifferentialCorrectionConfig, IODParams, OutfitB200, STATUS_NAMES
    ctx = OutfitB200(0)
    ctx.load_ephemeris(table)
    params = IODParams.builder(n_noise_realizations=args.noise_realizations, noise_scale=1.1, max_triplets=args.max_triplets)
    batch["traj_seed"] = (np.uint64(args.seed) ^ np.arange(len(ids), dtype=np.uint64))  # base_seed ^ id hash stand-in
    iod = ctx.fit_full_iod(batch, params, use_body_fixed=True)
    for name, r in zip(ids, iod):
        if r["status"] != 0:
            print(f"{name}: IOD failed: {STATUS_NAMES.get(int(r['status']), r['status'])} (cause {STATUS_NAMES.get(int(r['cause']), r['cause'])})")
            continue
        kind = {0: "Keplerian", 2: "Cometary"}[int(r["element_kind"])]
        print(f"{name}: {'Corrected' if r['corrected'] else 'Prelim'}Orbit {kind} epoch {r['epoch']:.6f} rms {r['rms']:.4f}")
        print("   a/q, e, i, Omega, omega, M/nu =", " ".join(f"{x:.10f}" for x in r["elem"]))
    if args.no_lsq:
        return 0
    lsq, fit = ctx.fit_lsq(batch, params, DifferentialCorrectionConfig.default(), initial_orbits=iod, use_body_fixed=True)
    kep = elements.lsq_to_keplerian(lsq)
    off = batch["traj_offset"].astype(np.int64)
    for t, (name, r) in enumerate(zip(ids, lsq)):
        if r["kind"] == 1:
            rej = int((fit["selection"][off[t]:off[t + 1]] == 1).sum())
            print(f"{name}: differential correction: {r['total_newton_iterations']} Newton steps, normalised rms "
                  f"{r['normalised_rms']:.4f}, {rej} observation(s) rejected")
            print("   a, e, i, Omega, omega, M =", " ".join(f"{x:.10f}" for x in kep["elem"][t]))
            print("   1-sigma              =", " ".join(f"{x:.3e}" for x in kep["sigma"][t]))
        elif r["kind"] == 2:
            print(f"{name}: differential correction fell back to the IOD orbit ({STATUS_NAMES.get(int(r['fallback_cause']))})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
