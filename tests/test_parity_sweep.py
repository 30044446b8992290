"""Full-configuration parity sweep (tools/parity_report.py): the CUDA path against the oracle on ALL 100 000
C3 trajectories, 20 000 ragged C4 trajectories, the 10 M C2 batch, 100 000 x 100 C5 entries and the
differential correction of 20 000 orbits; the figures are written to profiles/parity_r02.json (and to
gpurun_out/ so that they travel back from the GPU box) and the bounds asserted here are those figures minus a
hair.  OUTFIT_PARITY_SCALE (default 1.0) shrinks every size, e.g. 0.05 for a quick run.

The reference's end-to-end IOD test compares with 1e-11 / 1e-13 tolerances on ONE trajectory
(tests/test_gauss_iod.rs:150-168); at 10^5 trajectories the rule has to say what happens where Gauss' method
is ill-conditioned.  It is: integer fields equal, except on trajectories whose ORACLE answer itself flips
under a one-ulp move of its inputs (each flip carries that proof in the report); floats inside the north-star
tolerance on the measured fraction, and everything else within 256 x the oracle's own one-ulp sensitivity,
near-parabolic (chaotic initial guess) or oracle-discontinuous -- nothing unexplained."""
import json
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def report(oracle):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import parity_report
    scale = float(os.environ.get("OUTFIT_PARITY_SCALE", "1.0"))
    rep = parity_report.run(scale, None)
    name = "parity_r02.json" if scale == 1.0 else f"parity_r02_scale{scale:g}.json"
    for d in (os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")) if scale == 1.0 else (os.path.join(ROOT, "gpurun_out"),):
        try:
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, name), "w") as f:
                json.dump(rep, f, indent=1)
        except OSError:
            pass
    return rep


# measured (profiles/parity_r02.json, full scale): 0.9628 / 0.9662 / 0.9689 inside the plain north-star tolerance
# (the oracle against itself with RA moved by one ulp: 0.912 / 0.927 / 0.920)
@pytest.mark.parametrize("key,min_plain", [("c3", 0.955), ("c4", 0.955), ("c3_strict_no_noise", 0.96)])
def test_iod_sweep(report, key, min_plain):
    r = report[key]
    n = r["n_trajectories"]
    assert r["error_payloads_equal"]
    # selection flips: only where the oracle's own selection is one-ulp unstable, and rare
    assert r["n_flips"] == r["flips_proven_oracle_unstable"], [f for f in r["flips"] if not f["oracle_unstable_under_1ulp"]]
    assert r["n_flips"] <= max(1, int(5e-4 * n)), r["n_flips"]
    assert r["plain_both_fraction"] >= min_plain, r["plain_both_fraction"]
    # the GPU sits inside the algorithm's own one-ulp sensitivity: at least as many trajectories inside the plain
    # tolerance as the oracle keeps against itself when RA moves by one ulp (minus sampling noise)
    assert r["plain_both_fraction"] >= r["oracle_vs_itself_ra_plus_1ulp"]["plain_both_fraction"] - 0.02
    assert r["outside_plain"]["unexplained"] == 0, r["outside_plain"]
    assert r["epoch_abs_err_days_max"] <= 1e-8 or r["outside_plain"]["oracle_discontinuous_under_1ulp"] > 0


def test_lsq_sweep(report):
    r = report["lsq"]
    assert r["n_outcome_flips"] <= max(3, int(2e-3 * r["n_trajectories"])), r
    assert r["fallback_orbits_bitwise_equal"]
    assert r["plain_fraction"] >= 0.985, r  # measured 0.9911


def test_c2_sweep(report):
    r = report["c2"]
    # a status may differ only where the ORACLE's own status flips under random 1-ulp moves of the state (a Brent /
    # Newton budget on the edge): measured 1 in 10^7 (profiles/parity_r02.json)
    assert r["n_status_mismatch"] == r["status_mismatches_proven_oracle_unstable"] and r["n_status_mismatch"] <= max(1, int(1e-6 * r["n"]))
    assert r["ok_fraction"] > 0.999
    assert r["within_tolerance_fraction"] == 1.0, r
    assert r["r_rel_err"]["p50"] < 1e-13


def test_c5_sweep(report):
    r = report["c5"]
    assert r["status_exact_fraction"] == 1.0 and r["failed_entries_are_nan"]
    assert r["within_tolerance_fraction"] >= 0.999999, r
