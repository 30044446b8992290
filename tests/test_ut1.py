"""UT1 table reader (outfit_b200/ut1.py): JPL EOP2 namelist parsing and the step-function lookup that
hifitime's `Epoch::to_ut1` performs (observer_extension.rs:191-192).  CPU only; parity with the crate is
unpinned, so these tests pin the documented semantics on a synthetic file."""
import numpy as np
import pytest

from outfit_b200 import mpc80
from outfit_b200.ut1 import NS_PER_DAY, Ut1Table

EOP2 = """ $EOP2
 EOP2LBL='EOP. LAST DATUM 2025-01-01 '
 EOP2UNITS= 'MJD     ','MAS     ','MAS     ','MS      ','MAS     ','MAS     ','MS      ',
 EOP2=
 57000.00000,   30.0000,  280.0000,  35432.1000, 0.01, 0.01, 0.003, 0.1, 0.2, 0.3,
 57001.00000,   31.0000,  281.0000,  35433.2500, 0.01, 0.01, 0.003, 0.1, 0.2, 0.3,
 57002.00000,   32.0000,  282.0000,  35434.5000, 0.01, 0.01, 0.003, 0.1, 0.2, 0.3,
 57204.00000,   33.0000,  283.0000,  36300.0000, 0.01, 0.01, 0.003, 0.1, 0.2, 0.3,
 $END
 trailing text that must be ignored, 1, 2, 3
"""


def test_parse_eop2_namelist():
    t = Ut1Table.from_eop2_text(EOP2)
    assert len(t) == 4
    assert t.epoch_ns[0] == 57000 * NS_PER_DAY and t.offset_ns[1] == 35_433_250_000
    with pytest.raises(ValueError):
        Ut1Table.from_eop2_text("no marker here\n1,2,3,4\n")
    with pytest.raises(ValueError):
        Ut1Table.from_eop2_text(" EOP2=\n 57000.0, 1.0\n $END\n")


def test_step_function_lookup_is_strictly_after_the_entry():
    t = Ut1Table.from_eop2_text(EOP2)
    day = NS_PER_DAY
    assert t.offset_ns_at(56999 * day) == 0                      # before the table: Duration::ZERO
    assert t.offset_ns_at(57000 * day) == 0                      # `self > entry.epoch` is strict
    assert t.offset_ns_at(57000 * day + 1) == 35_432_100_000
    assert t.offset_ns_at(57001 * day + 5) == 35_433_250_000
    assert t.offset_ns_at(57100 * day) == 35_434_500_000         # no interpolation across the gap
    assert t.offset_ns_at(60000 * day) == 36_300_000_000         # after the last datum: last value
    # an unsorted table is scanned from the end like the crate does
    u = Ut1Table([57002.0, 57000.0, 57001.0], [3.0, 1.0, 2.0])
    assert u.offset_ns_at(57001 * day + 1) == 2_000_000 and u.offset_ns_at(57002 * day + 1) == 2_000_000


def test_mjd_ut1_is_tt_minus_offsets():
    t = Ut1Table.from_eop2_text(EOP2)
    tt = np.array([57001.5, 57003.25, 56000.0])
    got = t.mjd_ut1(tt)
    want = [57001.5 - (32.184 + 35.43325) / 86400.0, 57003.25 - (32.184 + 35.4345) / 86400.0, 56000.0 - 32.184 / 86400.0]
    assert np.abs(got - want).max() < 2e-11  # f64 MJD resolution
    # UT1 - UTC for callers that carry UTC: leap seconds (35 s in early 2015) minus TAI - UT1
    d = t.dut1_seconds(np.array([57001.5]), 35.0)
    assert abs(d[0] - (35.0 - 35.43325)) < 1e-9


def test_mpc80_batch_takes_the_table():
    recs = {"X": [dict(mjd_utc=57001.2, ra=1.0, dec=0.1, obscode="500"), dict(mjd_utc=57000.7, ra=1.0, dec=0.1, obscode="F51"),
                  dict(mjd_utc=57001.9, ra=1.0, dec=0.1, obscode="G96")]}
    t = Ut1Table.from_eop2_text(EOP2)
    _, b0 = mpc80.to_batch(recs)
    _, b1 = mpc80.to_batch(recs, ut1_table=t)
    assert np.array_equal(b0["mjd_tt"], b1["mjd_tt"]) and list(b1["mjd_tt"]) == sorted(b1["mjd_tt"])
    # UT1 - UTC = 35 s - (TAI - UT1): about -0.43 s in this synthetic table
    dut1 = (b1["mjd_ut1"] - b0["mjd_ut1"]) * 86400.0
    assert np.all(np.abs(dut1 - np.array([35.0 - 35.4321, 35.0 - 35.43325, 35.0 - 35.43325])) < 1e-5)
