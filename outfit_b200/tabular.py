"""Host-side reader for columnar observation tables (Parquet / Arrow / pandas) -> the trajectory records
`mpc80.to_batch` turns into the SoA batch of the C-ABI.

The reference ingests such tables through photom's `ObsDataset::from_lazy` (examples/run_full_iod.rs:69-87:
`tests/data/test_data_traj_str.parquet`, filtered on its `traj_id` column).  photom is not vendored and that
file is absent, so the schema beyond `traj_id` is not known here: the column names, the time format / scale
and the angle unit are explicit arguments, with the common survey-alert conventions as defaults (`jd` in UTC,
`ra` / `dec` in degrees, an MPC observatory code column).  Rows with a null trajectory id are dropped and
trajectories shorter than `min_obs` are skipped, as the example's polars filter does.
"""
import math

import numpy as np

from . import mpc80

DEFAULT_COLUMNS = dict(traj_id="traj_id", time="jd", ra="ra", dec="dec", obscode="obscode",
                       sigma_ra=None, sigma_dec=None)


def _to_columns(source, names):
    """dict / pandas DataFrame / pyarrow Table / path to a Parquet file -> {name: numpy array or list}."""
    if isinstance(source, str):
        import pyarrow.parquet as pq
        source = pq.read_table(source, columns=names)
    if hasattr(source, "column_names"):  # pyarrow Table
        return {n: source.column(n).to_pylist() for n in names}
    if hasattr(source, "columns") and hasattr(source, "__getitem__") and not isinstance(source, dict):  # pandas
        return {n: source[n].tolist() for n in names}
    return {n: list(source[n]) for n in names}


def parse(source, columns=None, time_format="jd", time_scale="utc", angles="deg", sigma_unit="arcsec",
          default_obscode="500", min_obs=3):
    """-> {traj_id: [record, ...]} in table order (records as mpc80 / ades produce them)."""
    col = dict(DEFAULT_COLUMNS)
    col.update(columns or {})
    want = [c for k, c in col.items() if c is not None and not (k == "obscode" and c is None)]
    try:
        data = _to_columns(source, want)
    except KeyError as e:
        raise KeyError(f"column {e} not in the table; pass columns={{...}} to map the schema") from None
    if time_scale not in ("utc", "tt"):
        raise ValueError("time_scale must be 'utc' or 'tt'")
    ang = math.pi / 180.0 if angles == "deg" else 1.0
    sig = {"arcsec": mpc80.ARCSEC, "deg": math.pi / 180.0, "rad": 1.0}[sigma_unit]
    out = {}
    n = len(data[col["time"]])
    for i in range(n):
        tid = data[col["traj_id"]][i]
        if tid is None or (isinstance(tid, float) and tid != tid):
            continue
        t = float(data[col["time"]][i])
        mjd = t - 2400000.5 if time_format == "jd" else t
        if time_scale == "tt":  # records carry UTC: undo TT - UTC at that date (leap seconds + 32.184 s)
            guess = mjd - (mpc80.tai_minus_utc(mjd) + 32.184) / 86400.0
            mjd = mjd - (mpc80.tai_minus_utc(guess) + 32.184) / 86400.0
        code = data[col["obscode"]][i] if col.get("obscode") else default_obscode
        rec = {"designation": str(tid), "number": "", "discovery": False, "mjd_utc": mjd,
               "ra": float(data[col["ra"]][i]) * ang, "dec": float(data[col["dec"]][i]) * ang,
               "obscode": str(code) if code is not None else default_obscode, "mag": float("nan"), "band": ""}
        if col.get("sigma_ra") and col.get("sigma_dec"):
            rec["sigma_ra"] = float(data[col["sigma_ra"]][i]) * sig
            rec["sigma_dec"] = float(data[col["sigma_dec"]][i]) * sig
        out.setdefault(str(tid), []).append(rec)
    return {k: v for k, v in out.items() if len(v) >= min_obs}
