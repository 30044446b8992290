"""Host-side astrometric error model and batch RMS correction: the two photom steps of `prepare_iod`
(src/initial_orbit_determination/obs_dataset_api.rs:262-265 of the reference:
`dataset.with_error_model(model).apply_model_errors().apply_batch_rms_correction(params.gap_max)`), which turn the
raw records of a reader into the per-observation sigmas the C-ABI batch carries (`sigma_ra`, `sigma_dec`).

**Parity unpinned.**  photom 0.4.0 is not vendored with the reference and its data files are not in this image, so
both steps are restated from their PUBLISHED definitions, not from photom's source:

* `FCCT14` -- Farnocchia, Chesley, Chamberlin & Tholen, "Star catalog position and proper motion corrections in
  asteroid astrometry", Icarus 245 (2015): astrometric weights as an RMS per observing station (their table of
  station-specific values), with defaults by observation type and by whether the reduction catalog is known.  The
  table below holds the station values of that publication as far as they could be restated without the paper at
  hand; every entry can be overridden (`rules=`), and a different table (e.g. photom's own rules file, or OrbFit's
  `fcct14.rules`) can be loaded with `load_rules`.
* batch RMS correction -- the rule OrbFit and Veres et al. (2017, Icarus 296) publish for over-represented nights:
  the observations of one station whose consecutive epochs are at most `gap_max` apart (8 h by default,
  IODParams::gap_max, mod.rs:321) form a batch; when a batch holds N > 4 observations their RMS is inflated by
  sqrt(N / 4), so that the batch as a whole weighs like four observations.

The reference's own numbers that depend on these steps (tests/test_gauss_iod.rs, trajectory.rs:628-694) also need
DE440 and UT1; `tests/test_reference_goldens.py` runs them when the files are mounted and reports the distance.
"""
import math

import numpy as np

ARCSEC = math.pi / 648000.0

# station code -> RMS in arcsec, or {catalog code: RMS, None: RMS for any other catalog}
FCCT14_STATION_RMS = {
    "704": 1.0, "G96": 0.5, "703": 1.0, "691": 0.6, "644": 0.6, "699": 0.8, "E12": 0.75, "608": 0.6, "D29": 0.75,
    "C51": 1.0, "J75": 1.0, "F51": 0.2, "F52": 0.2, "H01": 0.3, "673": 0.3, "645": 0.3, "689": 0.5, "950": 0.5,
    "568": {"t": 0.25, "q": 0.25, None: 0.5}, "309": 0.3, "T05": 0.5, "T08": 0.5, "W84": 0.5, "Y28": 0.3,
}
# defaults by observation type (MPC note 2, column 15) when the station has no entry of its own
FCCT14_DEFAULTS = {
    "ccd_known_catalog": 1.0,     # 'C' / blank with a reduction catalog code in column 72
    "ccd_unknown_catalog": 1.5,   # 'C' / blank without one
    "photographic": 2.5,          # 'P', 'A', 'N'
    "transit_circle": 1.5,        # 'T', 'M'
    "encoder": 0.75,              # 'E'
    "occultation": 0.2,           # 'H' (Hipparcos geocentric), 'O'
    "other": 1.5,
}
_TYPE_OF_NOTE2 = {"C": "ccd", " ": "ccd", "": "ccd", "B": "ccd", "n": "ccd", "P": "photographic", "A": "photographic",
                  "N": "photographic", "T": "transit_circle", "M": "transit_circle", "E": "encoder", "H": "occultation",
                  "O": "occultation"}


def model_rms_arcsec(record, rules=None, defaults=None):
    """RMS (arcsec, the same for RA cos(dec) and Dec) of one reader record: dict with `obscode` and optionally `catalog`
    (MPC column 72) and `note2` (column 15)."""
    rules = FCCT14_STATION_RMS if rules is None else {**FCCT14_STATION_RMS, **rules}
    defaults = FCCT14_DEFAULTS if defaults is None else {**FCCT14_DEFAULTS, **defaults}
    cat = (record.get("catalog") or "").strip() or None
    entry = rules.get(record["obscode"])
    if isinstance(entry, dict):
        return float(entry.get(cat, entry.get(None)))
    if entry is not None:
        return float(entry)
    kind = _TYPE_OF_NOTE2.get((record.get("note2") or "C"), "other")
    if kind == "ccd":
        return float(defaults["ccd_known_catalog" if cat else "ccd_unknown_catalog"])
    return float(defaults[kind])


def load_rules(path):
    """A plain-text rules table: `CODE RMS` or `CODE CATALOG RMS` per line, `#` comments."""
    out = {}
    with open(path) as f:
        for ln in f:
            p = ln.split("#", 1)[0].split()
            if len(p) == 2:
                out[p[0]] = float(p[1])
            elif len(p) == 3:
                out.setdefault(p[0], {})
                if not isinstance(out[p[0]], dict):
                    out[p[0]] = {None: out[p[0]]}
                out[p[0]][None if p[1] in ("*", "-") else p[1]] = float(p[2])
    return out


def apply_model_errors(records, rules=None, defaults=None):
    """`with_error_model(..).apply_model_errors()`: every record gets sigma_ra = sigma_dec = the model RMS (radians).
    Records that already carry their own uncertainties (ADES rmsRA / rmsDec) keep them."""
    out = []
    for r in records:
        r = dict(r)
        if "sigma_ra" not in r or "sigma_dec" not in r:
            s = model_rms_arcsec(r, rules, defaults) * ARCSEC
            r["sigma_ra"], r["sigma_dec"] = s, s
        out.append(r)
    return out


def batch_factors(mjd, obscode, gap_max, min_batch=5, reference=4.0):
    """Inflation factor of every observation of ONE trajectory: sqrt(N / 4) for the members of a batch of N >= 5
    observations of the same station whose consecutive epochs are at most gap_max days apart, else 1."""
    mjd = np.asarray(mjd, dtype=np.float64)
    fac = np.ones(len(mjd))
    codes = np.asarray(obscode)
    for code in np.unique(codes):
        idx = np.flatnonzero(codes == code)
        idx = idx[np.argsort(mjd[idx], kind="stable")]
        start = 0
        for j in range(1, len(idx) + 1):
            if j == len(idx) or mjd[idx[j]] - mjd[idx[j - 1]] > gap_max:
                n = j - start
                if n >= min_batch:
                    fac[idx[start:j]] = math.sqrt(n / reference)
                start = j
    return fac


def apply_batch_rms_correction(records, gap_max=8.0 / 24.0):
    """`apply_batch_rms_correction(gap_max)` on the records of ONE trajectory (after apply_model_errors)."""
    if not records:
        return []
    fac = batch_factors([r["mjd_utc"] for r in records], [r["obscode"] for r in records], gap_max)
    out = []
    for r, f in zip(records, fac):
        r = dict(r)
        r["sigma_ra"], r["sigma_dec"] = r["sigma_ra"] * f, r["sigma_dec"] * f
        out.append(r)
    return out


def prepare(trajectories, gap_max=8.0 / 24.0, rules=None, defaults=None):
    """The photom part of `prepare_iod` for {id: [record, ...]}: error model, then batch RMS correction.  The result
    feeds mpc80.to_batch / ades.to_batch, which honour per-record sigmas."""
    return {k: apply_batch_rms_correction(apply_model_errors(v, rules, defaults), gap_max) for k, v in trajectories.items()}
