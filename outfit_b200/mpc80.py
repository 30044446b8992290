"""Host-side reader for MPC 80-column optical astrometry (the format of the reference's
tests/data/*.obs, read there by the un-vendored `photom` crate) -> the SoA batch of the C-ABI.

This is the "wire format" row of SURVEY 8f: it lets BASELINE configs[0] (one trajectory from an MPC
80-column file) run without Rust.  What photom additionally does -- the FCCT14 astrometric error
model and the batch RMS correction -- lives in outfit_b200/error_model.py (restated from the published definitions,
parity with photom unpinned); `to_batch` honours the per-record sigmas it produces and otherwise applies a constant
sigma.

Columns (1-based, MPC "Format for optical astrometric observations"): 1-5 packed number, 6-12
packed provisional designation, 13 discovery asterisk, 14 note 1, 15 note 2, 16-32 date of
observation `YYYY MM DD.dddddd` (UTC), 33-44 RA `HH MM SS.ddd`, 45-56 Dec `sDD MM SS.dd`, 66-70
magnitude, 71 band, 78-80 observatory code.
"""
import math

import numpy as np

ARCSEC = math.pi / 648000.0
AU_KM = 149597870.7
ERAU = 6378.137 / AU_KM  # Earth equatorial radius in AU (observer_extension.rs:159-171 scale)

# TAI - UTC (s) from the given MJD on (IERS Bulletin C); TT = TAI + 32.184 s
_LEAP = [(41317.0, 10), (41499.0, 11), (41683.0, 12), (42048.0, 13), (42413.0, 14), (42778.0, 15), (43144.0, 16),
         (43509.0, 17), (43874.0, 18), (44239.0, 19), (44786.0, 20), (45151.0, 21), (45516.0, 22), (46247.0, 23),
         (47161.0, 24), (47892.0, 25), (48257.0, 26), (48804.0, 27), (49169.0, 28), (49534.0, 29), (50083.0, 30),
         (50630.0, 31), (51179.0, 32), (53736.0, 33), (54832.0, 34), (56109.0, 35), (57204.0, 36), (57754.0, 37)]

# MPC observatory parallax constants: east longitude (deg), rho cos(phi'), rho sin(phi') -- the codes of
# the reference's tests/data/2015AB.obs, transcribed from the MPC observatory list
OBSERVATORIES = {
    "500": (0.0, 0.0, 0.0),              # geocentre
    "204": (8.7700, 0.69740, 0.71440),   # Schiaparelli Observatory
    "291": (248.4010, 0.84950, 0.52640), # LPL/Spacewatch II
    "705": (254.17942, 0.841939, 0.538633),  # Apache Point
    "F51": (203.74409, 0.936241, 0.351543),  # Pan-STARRS 1, Haleakala
    "G96": (249.21128, 0.845111, 0.533614),  # Mt. Lemmon Survey
}


def calendar_to_mjd(year, month, day):
    """Gregorian calendar date (day may carry a fraction) -> MJD."""
    a = (14 - month) // 12
    y = year + 4800 - a
    m = month + 12 * a - 3
    jdn = int(day) + (153 * m + 2) // 5 + 365 * y + y // 4 - y // 100 + y // 400 - 32045
    return jdn - 2400001 + (day - int(day))  # JDN is the JD at noon: MJD = JDN - 2400000.5 - 0.5


def tai_minus_utc(mjd_utc):
    out = 0
    for start, v in _LEAP:
        if mjd_utc >= start:
            out = v
    return out


def utc_to_tt(mjd_utc):
    return mjd_utc + (tai_minus_utc(mjd_utc) + 32.184) / 86400.0


def parse_line(line):
    """One 80-column record -> dict, or None for blank / non-optical (satellite, radar, roving) lines."""
    line = line.rstrip("\n")
    if len(line) < 80 or line[14] in "RrVvSs":
        return None
    try:
        year, month, day = int(line[15:19]), int(line[20:22]), float(line[23:32])
        ra_h, ra_m, ra_s = int(line[32:34]), int(line[35:37]), float(line[38:44])
        sign = -1.0 if line[44] == "-" else 1.0
        de_d, de_m, de_s = int(line[45:47]), int(line[48:50]), float(line[51:56])
    except ValueError:
        return None
    mag = line[65:70].strip()
    return {
        "number": line[0:5].strip(), "designation": line[5:12].strip(), "discovery": line[12] == "*",
        "mjd_utc": calendar_to_mjd(year, month, day),
        "ra": (ra_h + ra_m / 60.0 + ra_s / 3600.0) * 15.0 * math.pi / 180.0,
        "dec": sign * (de_d + de_m / 60.0 + de_s / 3600.0) * math.pi / 180.0,
        "mag": float(mag) if mag else float("nan"), "band": line[70].strip(), "obscode": line[77:80],
        "note2": line[14], "catalog": line[71].strip(),  # observation type and reduction catalog: the error model's keys
    }


def parse(text, single_trajectory=False):
    """All optical records of an 80-column file as {id: [record, ...]} in file order, grouped by
    object id (number, else designation) -- or, with single_trajectory, as ONE trajectory named after
    the last record (how the reference's quick start treats tests/data/2015AB.obs, whose 37 records
    carry the two designations K09R05F and K15A00B of the same object)."""
    out = {}
    for ln in text.splitlines():
        rec = parse_line(ln)
        if rec is None:
            continue
        out.setdefault(rec["number"] or rec["designation"], []).append(rec)
    if single_trajectory and out:
        allrec = [r for v in out.values() for r in v]
        return {allrec[-1]["number"] or allrec[-1]["designation"]: allrec}
    return out


def parse_obscodes(text):
    """MPC observatory list (ObsCodes.html / obscodes.txt) -> {code: (east longitude deg, rho cos phi', rho sin phi')}.
    Fixed columns: code 1-3, longitude 5-13, cos 14-21, sin 22-30, name from 31.  Space-based / roving entries
    (blank constants) are skipped."""
    out = {}
    for ln in text.splitlines():
        if len(ln) < 30 or ln.startswith(("<", "Code")):
            continue
        code = ln[0:3]
        try:
            lon, rc, rs = float(ln[4:13]), float(ln[13:21]), float(ln[21:30].replace(" ", ""))
        except ValueError:
            continue
        out[code] = (lon, rc, rs)
    return out


def load_obscodes(path):
    with open(path, errors="replace") as f:
        return parse_obscodes(f.read())


def body_fixed_position(obscode, observatories=None):
    """Earth-fixed observer position in AU (observer_extension.rs:159-171: lon, rho cos, rho sin).
    `observatories`: extra / overriding {code: (lon deg, rho cos, rho sin)} entries, e.g. parse_obscodes()."""
    table = OBSERVATORIES if not observatories else {**OBSERVATORIES, **observatories}
    if obscode not in table:
        raise KeyError(f"observatory code {obscode!r} is not in the built-in table ({', '.join(sorted(OBSERVATORIES))}): pass "
                       "observatories={code: (east_lon_deg, rho_cos_phi, rho_sin_phi)} or the parsed MPC ObsCodes list "
                       "(mpc80.load_obscodes)")
    lon, rc, rs = table[obscode]
    lon = math.radians(lon)
    return np.array([ERAU * rc * math.cos(lon), ERAU * rc * math.sin(lon), ERAU * rs])


def to_batch(trajectories, sigma_arcsec=0.5, dut1_s=0.0, ut1_table=None, observatories=None):
    """{id: [record, ...]} -> (ids, batch): the body-fixed flavour of OutfitObsBatch.  Each trajectory is
    sorted by TT epoch (obs_dataset_api.rs:222-223); UT1 = UTC + dut1_s, or -- with a `ut1.Ut1Table`
    read from JPL's latest_eop2.long -- what `epoch.to_ut1(provider).to_mjd_tai_days()` gives
    (observer_extension.rs:191-192).  `observatories`: see body_fixed_position."""
    ids = list(trajectories)
    rows = []
    offs = [0]
    for k in ids:
        recs = sorted(trajectories[k], key=lambda r: r["mjd_utc"])
        rows.extend(recs)
        offs.append(len(rows))
    n = len(rows)
    mjd_utc = np.array([r["mjd_utc"] for r in rows])
    batch = {
        "traj_offset": np.asarray(offs, dtype=np.uint64),
        "mjd_tt": np.array([utc_to_tt(t) for t in mjd_utc]),
        "mjd_ut1": mjd_utc + dut1_s / 86400.0,
        "ra": np.array([r["ra"] for r in rows]), "dec": np.array([r["dec"] for r in rows]),
        "sigma_ra": np.full(n, sigma_arcsec * ARCSEC), "sigma_dec": np.full(n, sigma_arcsec * ARCSEC),
        "body_fixed": np.ascontiguousarray(np.stack([body_fixed_position(r["obscode"], observatories) for r in rows], axis=1)) if n
        else np.zeros((3, 0)),
        "noise_z": None,
    }
    for i, r in enumerate(rows):  # per-record uncertainties (ADES rmsRA / rmsDec) override the constant
        if "sigma_ra" in r and "sigma_dec" in r:
            batch["sigma_ra"][i], batch["sigma_dec"][i] = r["sigma_ra"], r["sigma_dec"]
    if ut1_table is not None and n:
        batch["mjd_ut1"] = ut1_table.mjd_ut1(batch["mjd_tt"])
    return ids, batch
