"""Result export: records -> named columns -> pandas / Parquet / JSON lines and back."""
import json

import numpy as np


def _records():
    from outfit_b200 import LSQ_RESULT_DTYPE, RESULT_DTYPE
    iod = np.zeros(3, dtype=RESULT_DTYPE)
    iod["status"] = [0, 14, 0]
    iod["cause"] = [0, 2, 0]
    iod["corrected"] = [1, 0, 0]
    iod["element_kind"] = [0, 0, 2]
    iod["epoch"] = [59000.5, 0.0, 59001.25]
    iod["elem"][0] = [2.5, 0.1, 0.2, 1.0, 2.0, 3.0]
    iod["elem"][2] = [0.9, 1.2, 0.3, 0.5, 0.6, 0.1]
    iod["rms"] = [0.8, 0.0, 3.5]
    iod["triplet_idx"] = [[0, 5, 11], [0, 0, 0], [1, 4, 9]]
    lsq = np.zeros(3, dtype=LSQ_RESULT_DTYPE)
    lsq["kind"] = [1, 0, 2]
    lsq["status"] = [0, 14, 0]
    lsq["fallback_cause"] = [0, 0, 20]
    lsq["epoch"] = [59000.5, 0.0, 59001.25]
    lsq["elem"][0] = [2.5, 0.01, 0.02, 0.03, 0.04, 1.5]
    lsq["sigma"][0] = np.arange(1, 7) * 1e-6
    c = np.arange(36, dtype=float).reshape(6, 6)
    lsq["covariance"][0] = (c + c.T).T.reshape(-1)  # symmetric, stored column-major
    lsq["normalised_rms"] = [0.9, 0.0, 3.5]
    return iod, lsq


def test_iod_and_lsq_columns_round_trip_through_parquet_and_jsonl(tmp_path):
    from outfit_b200 import export
    iod, lsq = _records()
    ids = ["K09R05F", "bad", "C/2020 X1"]
    ci, cl = export.iod_columns(ids, iod), export.lsq_columns(ids, lsq)
    assert ci["error"] == ["", "NoViableOrbit", ""] and ci["cause"][1] == "GaussNoRootsFound"
    assert ci["orbit"][0] == "CorrectedOrbit" and ci["element_type"][2] == "Cometary" and np.isnan(ci["element_0"][1])
    assert cl["result"] == ["DifferentialCorrection", "", "IODGauss (fallback)"] and cl["fallback_cause"][2] == "DifferentialCorrectionDiverged"
    assert cl["semi_major_axis"][0] == 2.5 and np.isnan(cl["semi_major_axis"][2]) and cl["cov_12"][0] == (1 * 6 + 2) + (2 * 6 + 1)
    p = tmp_path / "iod.parquet"
    export.write_parquet(str(p), ci)
    import pyarrow.parquet as pq
    t = pq.read_table(str(p)).to_pydict()
    assert t["traj_id"] == ids and t["element_0"][0] == 2.5 and t["triplet_2"] == [11, 0, 9]
    df = export.to_pandas(cl)
    assert list(df["traj_id"]) == ids and df["sigma_mean_longitude"][0] == 6e-6
    j = tmp_path / "lsq.jsonl"
    export.write_jsonl(str(j), cl)
    rows = [json.loads(x) for x in open(j)]
    assert len(rows) == 3 and rows[1]["semi_major_axis"] is None and rows[0]["newton_iterations"] == 0 and rows[0]["ok"] is True


def test_ephemeris_columns_and_display():
    from outfit_b200 import export
    out = np.arange(9 * 2 * 3, dtype=float).reshape(9, 2, 3)
    st = np.zeros((2, 3), dtype=np.int32)
    st[1, 2] = 10
    c = export.ephemeris_columns(["a", "b", "c"], [60000.0, 60001.0], out, st)
    assert list(c["orbit_id"]) == ["a", "b", "c", "a", "b", "c"] and list(c["mjd_tt"]) == [60000.0] * 3 + [60001.0] * 3
    assert c["ra"][4] == out[0, 1, 1] and c["d_dec_dt"][5] == out[8, 1, 2] and c["error"][5] == "InvalidOrbit"
    txt = export.display(0, 59000.5, [2.5, 0.1, 0.2, 1.0, 2.0, 3.0])
    assert "Keplerian" in txt and "semi_major_axis" in txt and "deg" in txt


def test_ephemeris_mode_epochs_follow_the_reference_expansion():
    """EphemerisMode::epochs (ephemeris/request.rs:246-267 and its tests :400-470): inclusive end when reachable,
    nothing for a non-positive step or start > end, integer-nanosecond stepping."""
    from outfit_b200 import ephemeris_mode_epochs as ep
    assert list(ep(("single", 60000.5))) == [60000.5]
    assert list(ep(("at", [60001.0, 60000.0, 60001.0]))) == [60001.0, 60000.0, 60001.0]
    r = ep(("range", 60000.0, 60002.0, 86400.0))
    assert list(r) == [60000.0, 60001.0, 60002.0]
    assert len(ep(("range", 60000.0, 60002.0, 86400.0 * 0.75))) == 3        # 0, 0.75, 1.5 (2.25 > 2)
    assert len(ep(("range", 60000.0, 59999.0, 60.0))) == 0 and len(ep(("range", 60000.0, 60001.0, 0.0))) == 0
    assert len(ep(("range", 60000.0, 60001.0, -5.0))) == 0
    long = ep(("range", 60000.0, 60100.0, 0.1))                             # 86.4 M steps would drift as a float sum
    assert len(long) == 100 * 864000 + 1 and long[-1] == 60100.0
