"""The C++17 host-side I/O header (outfit_b200/host/outfit_b200_io.hpp: MPC 80-column reader, UT1 table,
Keplerian form of the LSQ records) against the Python modules that do the same.  CPU only."""
import os
import subprocess

import numpy as np

from outfit_b200 import elements, mpc80
from outfit_b200.ut1 import Ut1Table

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINE = "     K09R05F* C2009 09 15.22735 22 52 23.37 -14 47 05.4          20.7 Vr~097wG96"
EOP2 = " $EOP2\n EOP2=\n 55000.00000, 30.0, 280.0, 34100.2500, 0.01,\n 55089.00000, 31.0, 281.0, 34160.7500, 0.01,\n 55091.00000, 31.0, 281.0, 34162.1250, 0.01,\n $END\n"


def test_cpp_host_io_matches_python(tmp_path):
    lines = [LINE[:23] + f"{15.22735 + d:8.5f} " + LINE[32:] for d in (2.0, 0.0, 1.0)]
    lines[1] = lines[1][:77] + "F51"
    lines.append(LINE[:14] + "S" + LINE[15:])        # satellite record: skipped
    lines.append("short line")
    obs = tmp_path / "x.obs"
    obs.write_text("\n".join(lines) + "\n")
    eop = tmp_path / "eop2.long"
    eop.write_text(EOP2)
    exe = str(tmp_path / "io_smoke")
    libdir = os.path.join(ROOT, "outfit_b200")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "tests", "cpp", "host_io_smoke.cpp"),
                           "-o", exe, "-L" + libdir, "-loutfit_b200", "-Wl,-rpath," + libdir])
    from outfit_b200 import de_reader, synth
    de = str(tmp_path / "synth.440")
    de_reader.write_de_binary(de, synth.make_ephemeris_table(n_blocks=12))
    out = subprocess.run([exe, str(obs), str(eop), de], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    rows = [ln.split() for ln in out.stdout.splitlines()]
    assert rows[0] == ["ntraj", "1", "id", "K09R05F", "n", "3"]
    assert ["sites", "3"] in rows  # parse_obscodes: 500, G96, F51 (the space-based entry is skipped)
    got = np.array([[float(x) for x in r[1:]] for r in rows if r[0] == "obs"])
    # the Python chain on the same files
    tr = mpc80.parse(obs.read_text(), single_trajectory=True)
    assert list(tr) == ["K09R05F"] and len(tr["K09R05F"]) == 3
    table = Ut1Table.from_eop2_text(EOP2)
    recs = tr["K09R05F"]                              # file order (the C++ rows are printed before sorting)
    for g, r in zip(got, recs):
        tt = mpc80.utc_to_tt(r["mjd_utc"])
        assert g[0] == tt and g[1] == r["ra"] and g[2] == r["dec"] and g[3] == 0.5 * mpc80.ARCSEC
        assert np.allclose(g[4:7], mpc80.body_fixed_position(r["obscode"]), rtol=1e-14, atol=0)  # deg -> rad rounding
        assert g[7] == table.mjd_ut1(tt)[0]
    assert [r for r in rows if r[0] == "batch"][0][1:] == ["1", "3", "sorted", "1"]
    # Keplerian form of an equinoctial record with a covariance
    eq = np.array([[1.8017360713, 0.2693736809404963, 0.08856415260522467, 0.0008089970142830734, 0.10168201110394352, 1.693697008]])
    cov = np.zeros(36)
    for i in range(36):
        cov[i] = 1e-8 * (1 + i // 7) if i % 7 == 0 else 1e-10 * ((i * 7) % 5)
    cm = cov.reshape(6, 6)                            # [col][row]
    for c in range(6):
        for r in range(c):
            cm[c, r] = cm[r, c]
    kep = np.array([float(x) for x in [r for r in rows if r[0] == "kep"][0][1:]])
    assert list(kep) == [1.8017360713, 0.2835591457, 0.20267383289999996, 0.007955979, 1.2451951388, 0.4405458902000001]
    assert np.abs(elements.equinoctial_to_keplerian(eq)[0] - kep).max() <= 4e-16
    want = elements.propagate_covariance(cm.T[None], elements.jacobian_to_keplerian(eq))[0]
    gcov = np.array([float(x) for x in [r for r in rows if r[0] == "cov"][0][1:]]).reshape(6, 6).T
    assert np.allclose(gcov, want, rtol=1e-12, atol=1e-24)
    gsig = np.array([float(x) for x in [r for r in rows if r[0] == "sig"][0][1:]])
    assert np.allclose(gsig, np.sqrt(np.diag(want)), rtol=1e-12)
    # DE binary file: the C++ parser against the Python one on the same (synthetic) file
    want_de = de_reader.read_de_binary(de)
    g = [r for r in rows if r[0] == "de"][0][1:]
    assert int(g[0]) == want_de["cheb"].shape[0] and int(g[1]) == want_de["cheb"].shape[1]
    assert [float(x) for x in g[2:6]] == [want_de["jd_start"], want_de["jd_end"], want_de["block_days"], want_de["emrat"]]
    assert int(g[6]) == want_de["numde"] and [int(x) for x in g[7:16]] == [int(v) for v in want_de["ipt"].reshape(-1)]
    assert abs(float(g[16]) - float(np.sum(want_de["cheb"]))) <= 1e-9 * abs(float(np.sum(want_de["cheb"])))


def test_cpp_quick_start_example_builds_and_parses(tmp_path):
    """examples/run_full_iod.cpp: compiles against the two headers; --dry-run goes from the reference's
    quick-start observations (fixture records re-encoded as 80-column lines) + a DE-layout file to the batch."""
    import json
    from outfit_b200 import de_reader, synth
    d = json.load(open(os.path.join(ROOT, "tests", "golden", "config1_2015AB.json")))
    lines = []
    for r in d["records"]:
        mjd = r["mjd_utc"]
        jd = mjd + 2400000.5 + 0.5
        z = int(jd)
        a = int((z - 1867216.25) / 36524.25)
        aa = z + 1 + a - a // 4
        b = aa + 1524
        c = int((b - 122.1) / 365.25)
        dd = int(365.25 * c)
        e = int((b - dd) / 30.6001)
        day = b - dd - int(30.6001 * e) + (jd - z)
        month = e - 1 if e < 14 else e - 13
        year = c - 4716 if month > 2 else c - 4715
        ra_h = np.degrees(r["ra"]) / 15.0
        hh = int(ra_h); mm = int((ra_h - hh) * 60); ss = (ra_h - hh - mm / 60.0) * 3600
        dec = np.degrees(abs(r["dec"]))
        dg = int(dec); dm = int((dec - dg) * 60); ds = (dec - dg - dm / 60.0) * 3600
        sign = "-" if r["dec"] < 0 else "+"
        ln = f"     {d['designation']:<7s}  C{year:4d} {month:02d} {day:08.5f} {hh:02d} {mm:02d} {ss:06.3f}{sign}{dg:02d} {dm:02d} {ds:05.2f}"
        ln = ln.ljust(77) + r["obscode"]
        assert len(ln) == 80, (len(ln), ln)
        lines.append(ln)
    obs = tmp_path / "2015AB.obs"
    obs.write_text("\n".join(lines) + "\n")
    back = mpc80.parse(obs.read_text(), single_trajectory=True)
    assert sum(len(v) for v in back.values()) == len(d["records"])
    de = str(tmp_path / "synth.440")
    de_reader.write_de_binary(de, synth.make_ephemeris_table(mjd_start=54900.0, n_blocks=80))
    exe = str(tmp_path / "run_full_iod")
    libdir = os.path.join(ROOT, "outfit_b200")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "examples", "run_full_iod.cpp"),
                           "-o", exe, "-L" + libdir, "-loutfit_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe, str(obs), de, "--dry-run"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert f"{len(d['records'])} observations" in out.stdout and "80 blocks of 32 days" in out.stdout
    # without --dry-run and without a GPU the context refuses to start: no CPU fallback
    import torch
    out = subprocess.run([exe, str(obs), de], capture_output=True, text=True)
    if not torch.cuda.is_available():
        assert out.returncode == 1 and "no CPU fallback" in out.stderr
    else:  # on a GPU box: the whole chain, same orbit as the Python example / the reference's golden to ~1e-3
        assert out.returncode == 0, out.stderr
        assert "CorrectedOrbit Keplerian" in out.stdout and "differential correction:" in out.stdout, out.stdout
        a = float(out.stdout.split("CorrectedOrbit")[1].splitlines()[1].split()[0])
        assert abs(a - 1.801740835743616) < 2e-3
