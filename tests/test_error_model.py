"""The photom steps of `prepare_iod` (obs_dataset_api.rs:262-265) restated on the host side: FCCT14 station RMS and
the batch RMS correction.  Parity with photom is unpinned (crate not vendored): these tests pin the SEMANTICS of the
published rules and the plumbing into the batch."""
import math

import numpy as np

from outfit_b200 import error_model as em
from outfit_b200 import mpc80

LINE = "     K09R05F* C2009 09 15.22735 22 52 23.37 -14 47 05.4          20.7 Vr~097wG96"


def test_station_rms_catalog_rules_and_defaults():
    assert em.model_rms_arcsec({"obscode": "F51"}) == 0.2 and em.model_rms_arcsec({"obscode": "704"}) == 1.0
    assert em.model_rms_arcsec({"obscode": "568", "catalog": "t"}) == 0.25
    assert em.model_rms_arcsec({"obscode": "568", "catalog": "c"}) == 0.5
    # a station without an entry: by observation type and whether the reduction catalog is known
    assert em.model_rms_arcsec({"obscode": "Z99", "catalog": "U", "note2": "C"}) == 1.0
    assert em.model_rms_arcsec({"obscode": "Z99", "catalog": "", "note2": "C"}) == 1.5
    assert em.model_rms_arcsec({"obscode": "Z99", "note2": "P"}) == 2.5
    # overrides
    assert em.model_rms_arcsec({"obscode": "F51"}, rules={"F51": 0.15}) == 0.15


def test_rules_file_roundtrip(tmp_path):
    p = tmp_path / "rules.txt"
    p.write_text("# code [catalog] rms\nG96 0.45\n568 t 0.2\n568 * 0.6\n")
    r = em.load_rules(str(p))
    assert r["G96"] == 0.45 and r["568"] == {"t": 0.2, None: 0.6}
    assert em.model_rms_arcsec({"obscode": "568", "catalog": "x"}, rules=r) == 0.6


def test_batch_factor_rule():
    # one station, 6 observations within 8 h of each other, then 2 more a day later; another station interleaved
    mjd = [59000.00, 59000.01, 59000.02, 59000.30, 59000.31, 59000.32, 59001.5, 59001.51, 59000.015]
    code = ["G96"] * 8 + ["F51"]
    f = em.batch_factors(mjd, code, 8.0 / 24.0)
    assert np.allclose(f[:6], math.sqrt(6 / 4)) and np.all(f[6:] == 1.0)
    # a gap longer than gap_max splits the night: two batches of 3 stay uncorrected
    f2 = em.batch_factors(mjd, code, 0.1)
    assert np.all(f2 == 1.0)
    # exactly four observations are not inflated, five are
    assert np.all(em.batch_factors([0, .01, .02, .03], ["G96"] * 4, 1 / 3) == 1.0)
    assert np.allclose(em.batch_factors([0, .01, .02, .03, .04], ["G96"] * 5, 1 / 3), math.sqrt(5 / 4))


def test_prepare_feeds_the_batch_sigmas():
    lines = [LINE[:23] + f"{15.22735 + 0.004 * d:8.5f} " + LINE[32:] for d in range(6)]
    lines.append(LINE[:23] + f"{17.22735:8.5f} " + LINE[32:77] + "F51")
    tr = mpc80.parse("\n".join(lines), single_trajectory=True)
    rec = list(tr.values())[0]
    assert rec[0]["note2"] == "C" and rec[0]["catalog"] == "r"
    prepared = em.prepare(tr)
    ids, batch = mpc80.to_batch(prepared)
    s = batch["sigma_ra"] / em.ARCSEC
    assert np.allclose(s[:6], 0.5 * math.sqrt(6 / 4)) and np.isclose(s[6], 0.2)
    assert np.array_equal(batch["sigma_ra"], batch["sigma_dec"])
    # ADES-style records that carry their own uncertainties keep them through the model step
    own = em.apply_model_errors([dict(rec[0], sigma_ra=1e-7, sigma_dec=2e-7)])
    assert own[0]["sigma_ra"] == 1e-7 and own[0]["sigma_dec"] == 2e-7
