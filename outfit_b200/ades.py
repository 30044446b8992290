"""Host-side reader for ADES XML optical astrometry (the format of the reference's
tests/data/example_ades*.xml / flat_ades.xml, read there by the un-vendored `photom` crate) -> the
trajectory records `mpc80.to_batch` turns into the SoA batch of the C-ABI.

Both ADES layouts are accepted: `<ades><obsBlock><obsData><optical>` (with an `<obsContext>`) and the flat
`<ades><optical>` list.  Per `<optical>` element: `stn` (MPC observatory code), `obsTime` (ISO 8601 UTC,
`Z`), `ra` / `dec` (decimal degrees), optional `rmsRA` / `rmsDec` (arcsec; rmsRA is RA*cos(Dec) as ADES
defines it), `mag`, `band`; the trajectory id is the first present of `trkSub`, `permID`, `provID`,
`artSat`, `trkID`.  What photom additionally does and this reader does NOT: the FCCT14 error model for
records without rms fields (a constant sigma is applied by `to_batch` instead) -- same caveat as mpc80.py.
"""
import math
import xml.etree.ElementTree as ET

from . import mpc80

ID_FIELDS = ("trkSub", "permID", "provID", "artSat", "trkID")


def iso_utc_to_mjd(s):
    """`YYYY-MM-DDThh:mm:ss[.fff]Z` -> MJD(UTC)."""
    s = s.strip()
    if s.endswith("Z"):
        s = s[:-1]
    date, _, clock = s.partition("T")
    y, m, d = (int(x) for x in date.split("-"))
    hh, mm, ss = (clock.split(":") + ["0", "0"])[:3] if clock else ("0", "0", "0")
    frac = (int(hh) * 3600.0 + int(mm) * 60.0 + float(ss)) / 86400.0
    return mpc80.calendar_to_mjd(y, m, d) + frac


def _text(el, tag):
    c = el.find(tag)
    return c.text.strip() if c is not None and c.text and c.text.strip() else None


def parse(text, id_fields=ID_FIELDS):
    """All `<optical>` records as {id: [record, ...]} in file order.  Records carry `sigma_ra` /
    `sigma_dec` (radians) when the file gives rmsRA / rmsDec."""
    root = ET.fromstring(text)
    if root.tag != "ades":
        raise ValueError(f"not an ADES document: root element <{root.tag}>")
    out = {}
    for el in root.iter("optical"):
        stn, t, ra, dec = _text(el, "stn"), _text(el, "obsTime"), _text(el, "ra"), _text(el, "dec")
        if stn is None or t is None or ra is None or dec is None:
            continue  # offset / occultation records (raStar, deltaRA ...) are not plain astrometry
        ident = next((v for v in (_text(el, f) for f in id_fields) if v is not None), None)
        if ident is None:
            raise ValueError("optical record without trkSub / permID / provID")
        rec = {"designation": ident, "number": "", "discovery": False, "mjd_utc": iso_utc_to_mjd(t),
               "ra": math.radians(float(ra)), "dec": math.radians(float(dec)), "obscode": stn,
               "mag": float(_text(el, "mag")) if _text(el, "mag") else float("nan"), "band": _text(el, "band") or ""}
        rms_ra, rms_dec = _text(el, "rmsRA"), _text(el, "rmsDec")
        if rms_ra is not None and rms_dec is not None:
            rec["sigma_ra"] = float(rms_ra) * mpc80.ARCSEC
            rec["sigma_dec"] = float(rms_dec) * mpc80.ARCSEC
        out.setdefault(ident, []).append(rec)
    return out
