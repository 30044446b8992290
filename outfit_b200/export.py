"""Host-side result export: the result records of the C-ABI as named columns, for pandas / Arrow /
Parquet / JSON lines -- the way back of `mpc80` / `ades` / `tabular`, which bring observations in.

The reference returns `FullOrbitResult = HashMap<TrajId, Result<FitOrbitResult, OutfitError>>`
(initial_orbit_determination/obs_dataset_api.rs:145-207) and prints orbits through the `Display` impls of its element
types (orbit_type/keplerian_element.rs:429, equinoctial_element.rs:1170, cometary_element.rs:508); it has no file
format of its own.  The columns below carry every field of those values: an `Err` is a row with `ok = False` and the
error's variant name in `error`.
"""
import json

import numpy as np

from .api import STATUS_NAMES

KEPLERIAN = ("semi_major_axis", "eccentricity", "inclination", "ascending_node_longitude", "periapsis_argument",
             "mean_anomaly")                      # KeplerianElements, keplerian_element.rs:145-160
COMETARY = ("perihelion_distance", "eccentricity", "inclination", "ascending_node_longitude", "periapsis_argument",
            "true_anomaly")                       # CometaryElements, cometary_element.rs:150-165
EQUINOCTIAL = ("semi_major_axis", "eccentricity_sin_lon", "eccentricity_cos_lon", "tan_half_incl_sin_node",
               "tan_half_incl_cos_node", "mean_longitude")   # EquinoctialElements, equinoctial_element.rs:188-196
_KIND = {0: "Keplerian", 1: "Equinoctial", 2: "Cometary"}


def iod_columns(ids, results):
    """FitOrbitResult::IODGauss rows: one per trajectory, in batch order.  results: RESULT_DTYPE array."""
    r = np.asarray(results)
    ok = r["status"] == 0
    cols = {
        "traj_id": list(ids), "ok": ok, "error": [STATUS_NAMES.get(int(s), str(int(s))) if s else "" for s in r["status"]],
        "cause": [STATUS_NAMES.get(int(c), "") if c else "" for c in r["cause"]], "attempts": r["attempts"].astype(np.int64),
        "orbit": ["CorrectedOrbit" if c else "PrelimOrbit" for c in r["corrected"]],   # GaussResult, gauss_result.rs:99-102
        "element_type": [_KIND.get(int(k), "") for k in r["element_kind"]],
        "reference_epoch": np.where(ok, r["epoch"], np.nan), "rms": np.where(ok, r["rms"], np.nan),
        "triplet_0": r["triplet_idx"][:, 0], "triplet_1": r["triplet_idx"][:, 1], "triplet_2": r["triplet_idx"][:, 2],
        "triplet_rank": r["triplet_rank"], "realization": r["realization"], "span": r["span"],
    }
    for j in range(6):
        cols[f"element_{j}"] = np.where(ok, r["elem"][:, j], np.nan)
    return cols


def lsq_columns(ids, results):
    """FitOrbitResult::DifferentialCorrection rows (equinoctial elements, 1-sigma, covariance upper triangle), the IOD
    fallback rows and the errors.  results: LSQ_RESULT_DTYPE array."""
    r = np.asarray(results)
    ok = r["kind"] != 0
    cor = r["kind"] == 1
    cols = {
        "traj_id": list(ids), "ok": ok, "error": [STATUS_NAMES.get(int(s), str(int(s))) if s else "" for s in r["status"]],
        "result": [("DifferentialCorrection", "IODGauss (fallback)")[k - 1] if k else "" for k in r["kind"]],
        "fallback_cause": [STATUS_NAMES.get(int(c), "") if c else "" for c in r["fallback_cause"]],
        "reference_epoch": np.where(ok, r["epoch"], np.nan), "normalised_rms": np.where(ok, r["normalised_rms"], np.nan),
        "newton_iterations": r["total_newton_iterations"].astype(np.int64), "num_measurements": r["num_measurements"].astype(np.int64),
    }
    for j, name in enumerate(EQUINOCTIAL):
        cols[name] = np.where(cor, r["elem"][:, j], np.nan)
        cols["sigma_" + name] = np.where(cor, r["sigma"][:, j], np.nan)
    cov = r["covariance"].reshape(-1, 6, 6)
    for a in range(6):
        for b in range(a, 6):
            cols[f"cov_{a}{b}"] = np.where(cor, cov[:, b, a], np.nan)   # column-major storage: [b][a] = C(a, b)
    return cols


def ephemeris_columns(orbit_ids, mjd_tt, out, status):
    """Combined rows (ephemeris/request.rs:102-205), one per (epoch, orbit): out (9, E, n), status (E, n)."""
    from .api import OutfitB200
    E, n = status.shape
    cols = {"orbit_id": np.tile(np.asarray(list(orbit_ids), dtype=object), E), "mjd_tt": np.repeat(np.asarray(mjd_tt), n),
            "ok": (status == 0).ravel(), "error": [STATUS_NAMES.get(int(s), str(int(s))) if s else "" for s in status.ravel()]}
    for j, name in enumerate(OutfitB200.EPHEMERIS_FIELDS):
        cols[name] = out[j].reshape(-1)
    return cols


def to_pandas(cols):
    import pandas as pd
    return pd.DataFrame(cols)


def to_arrow(cols):
    import pyarrow as pa
    return pa.table({k: (v if isinstance(v, list) else np.asarray(v)) for k, v in cols.items()})


def write_parquet(path, cols):
    import pyarrow.parquet as pq
    pq.write_table(to_arrow(cols), path)


def write_jsonl(path, cols):
    """One JSON object per row; NaN (an absent value) becomes null."""
    keys = list(cols)
    n = len(cols[keys[0]])
    with open(path, "w") as f:
        for i in range(n):
            row = {}
            for k in keys:
                v = cols[k][i]
                if isinstance(v, (np.floating, float)):
                    v = None if v != v else float(v)
                elif isinstance(v, (np.integer,)):
                    v = int(v)
                elif isinstance(v, (np.bool_,)):
                    v = bool(v)
                row[k] = v
            f.write(json.dumps(row) + "\n")


def display(kind, epoch, elem):
    """The text the reference's `Display` impls give for one orbit, field by field (angles in radians and degrees)."""
    names = {0: KEPLERIAN, 1: EQUINOCTIAL, 2: COMETARY}[int(kind)]
    lines = [f"{_KIND[int(kind)]} elements @ epoch (MJD TT): {epoch:.6f}"]
    for nme, v in zip(names, elem):
        angle = nme in ("inclination", "ascending_node_longitude", "periapsis_argument", "mean_anomaly", "true_anomaly", "mean_longitude")
        lines.append(f"  {nme:28s} = {v:.12g}" + (f" rad ({np.degrees(v):.6f} deg)" if angle else ""))
    return "\n".join(lines)
