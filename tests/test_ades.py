"""ADES XML reader (outfit_b200/ades.py): both layouts, ids, units, per-record sigmas -> batch.  CPU only."""
import math

import numpy as np
import pytest

from outfit_b200 import ades, mpc80

BLOCK = """<?xml version='1.0' encoding='UTF-8'?>
<ades version="2017">
  <obsBlock>
    <obsContext><observatory><mpcCode>F51</mpcCode></observatory></obsContext>
    <obsData>
      <optical><permID>1234456</permID><trkSub>aa</trkSub><mode>CCD</mode><stn>F51</stn>
        <obsTime>2016-08-29T12:32:34Z</obsTime><ra>0</ra><dec>90</dec><rmsRA>0.15</rmsRA><rmsDec>0.13</rmsDec>
        <mag>21.9</mag><band>w</band></optical>
      <optical><permID>1234457</permID><trkSub>aa</trkSub><mode>CCD</mode><stn>F51</stn>
        <obsTime>2016-08-29T12:02:34.5Z</obsTime><ra>0.1</ra><dec>-30.0</dec><rmsRA>0.15</rmsRA><rmsDec>0.13</rmsDec></optical>
      <optical><provID>2016 QW1</provID><mode>CCD</mode><stn>G96</stn>
        <obsTime>2016-08-30T00:00:00Z</obsTime><ra>180.5</ra><dec>12.25</dec></optical>
      <optical><trkSub>bb</trkSub><stn>G96</stn><obsTime>2016-08-30T00:00:00Z</obsTime><raStar>1.0</raStar></optical>
    </obsData>
  </obsBlock>
</ades>
"""
FLAT = """<ades version="2017">
  <optical><permID>5</permID><mode>PHO</mode><stn>500</stn><obsTime>1948-11-06T20:31:03.4Z</obsTime>
    <ra>18.75</ra><dec>-0.00</dec><mag>10.2</mag><band>V</band></optical>
</ades>
"""


def test_iso_time_to_mjd():
    assert ades.iso_utc_to_mjd("2000-01-01T12:00:00Z") == 51544.5
    assert abs(ades.iso_utc_to_mjd("2016-08-29T12:32:34Z") - (57629.0 + (12 * 3600 + 32 * 60 + 34) / 86400.0)) < 1e-11
    assert ades.iso_utc_to_mjd("1858-11-17T00:00:00.0Z") == 0.0


def test_block_layout_ids_units_and_sigmas():
    tr = ades.parse(BLOCK)
    assert list(tr) == ["aa", "2016 QW1"]           # trkSub wins over permID; offset record skipped
    a = tr["aa"]
    assert len(a) == 2 and a[0]["obscode"] == "F51"
    assert a[0]["dec"] == math.radians(90.0) and a[1]["dec"] == math.radians(-30.0) and a[1]["ra"] == math.radians(0.1)
    assert a[0]["sigma_ra"] == 0.15 * mpc80.ARCSEC and a[0]["sigma_dec"] == 0.13 * mpc80.ARCSEC
    assert "sigma_ra" not in tr["2016 QW1"][0] and math.isnan(tr["2016 QW1"][0]["mag"])
    ids, b = mpc80.to_batch(tr, sigma_arcsec=0.5)
    assert ids == ["aa", "2016 QW1"] and list(b["traj_offset"]) == [0, 2, 3]
    assert b["mjd_tt"][0] < b["mjd_tt"][1]          # time-sorted inside the trajectory
    assert np.allclose(b["sigma_ra"], [0.15 * mpc80.ARCSEC, 0.15 * mpc80.ARCSEC, 0.5 * mpc80.ARCSEC], rtol=0, atol=0)
    assert b["body_fixed"].shape == (3, 3) and np.all(np.linalg.norm(b["body_fixed"], axis=0) < 4.3e-5)
    # UTC -> TT: 36 leap seconds + 32.184 s in August 2016
    assert abs((b["mjd_tt"][2] - 57630.0) * 86400.0 - 68.184) < 1e-5


def test_flat_layout_and_errors():
    tr = ades.parse(FLAT)
    assert list(tr) == ["5"] and tr["5"][0]["band"] == "V" and tr["5"][0]["mag"] == 10.2
    with pytest.raises(ValueError):
        ades.parse("<notades/>")
    with pytest.raises(ValueError):
        ades.parse("<ades><optical><stn>500</stn><obsTime>2000-01-01T00:00:00Z</obsTime><ra>1</ra><dec>2</dec></optical></ades>")
