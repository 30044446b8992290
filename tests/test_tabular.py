"""Columnar (Parquet / Arrow / dict) observation tables -> trajectory records -> batch.  CPU only."""
import numpy as np
import pytest

from outfit_b200 import mpc80, tabular


def _table():
    return {
        "traj_id": ["A", "A", None, "B", "A", "B", "C", "B"],
        "jd": [2457000.5 + 0.1 * i for i in range(8)],
        "ra": [10.0 + i for i in range(8)],
        "dec": [-5.0 + i for i in range(8)],
        "obscode": ["F51", "F51", "F51", "G96", "F51", "G96", "500", "G96"],
        "e_ra": [0.1] * 8, "e_dec": [0.2] * 8,
    }


def test_dict_table_grouping_filter_and_units():
    tr = tabular.parse(_table())
    assert list(tr) == ["A", "B"]                       # null id dropped, "C" has fewer than 3 observations
    assert [r["obscode"] for r in tr["B"]] == ["G96"] * 3
    assert tr["A"][0]["mjd_utc"] == 57000.0 and abs(tr["A"][2]["mjd_utc"] - 57000.4) < 1e-9
    assert tr["A"][1]["ra"] == np.radians(11.0) and tr["A"][1]["dec"] == np.radians(-4.0)
    ids, b = mpc80.to_batch(tr, sigma_arcsec=0.3)
    assert ids == ["A", "B"] and list(b["traj_offset"]) == [0, 3, 6] and np.all(b["sigma_ra"] == 0.3 * mpc80.ARCSEC)
    assert len(tabular.parse(_table(), min_obs=1)) == 3


def test_column_mapping_sigmas_and_time_scales():
    t = _table()
    t["mjd_tt"] = [mpc80.utc_to_tt(x - 2400000.5) for x in t.pop("jd")]
    tr = tabular.parse(t, columns=dict(time="mjd_tt", sigma_ra="e_ra", sigma_dec="e_dec"), time_format="mjd", time_scale="tt")
    assert abs(tr["A"][0]["mjd_utc"] - 57000.0) < 1e-9          # TT -> UTC undone (35 leap seconds + 32.184 s)
    assert tr["A"][0]["sigma_ra"] == 0.1 * mpc80.ARCSEC and tr["A"][0]["sigma_dec"] == 0.2 * mpc80.ARCSEC
    _, b = mpc80.to_batch(tr)
    assert np.all(b["sigma_dec"] == 0.2 * mpc80.ARCSEC)
    assert abs(b["mjd_tt"][0] - t["mjd_tt"][0]) < 1e-9
    with pytest.raises(KeyError):
        tabular.parse(_table(), columns=dict(ra="alpha"))
    rad = tabular.parse(_table(), angles="rad")
    assert rad["A"][0]["ra"] == 10.0


def test_parquet_and_arrow_sources(tmp_path):
    pa = pytest.importorskip("pyarrow")
    import pyarrow.parquet as pq
    tab = pa.table(_table())
    path = str(tmp_path / "obs.parquet")
    pq.write_table(tab, path)
    strip = lambda tr: {k: [{f: v for f, v in r.items() if f != "mag"} for r in rs] for k, rs in tr.items()}  # mag is NaN
    a, b, c = strip(tabular.parse(tab)), strip(tabular.parse(path)), strip(tabular.parse(_table()))
    assert a == b == c
    pd = pytest.importorskip("pandas")
    assert strip(tabular.parse(pd.DataFrame(_table()))) == c
