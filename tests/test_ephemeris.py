"""Two-body `Combined` ephemeris (BASELINE configs[4]): the oracle restatement on
CPU (size-independent properties) and the CUDA path against it (GPU)."""
import numpy as np
import pytest

FIELDS = ("ra", "dec", "geocentric_dist", "heliocentric_dist", "phase_angle", "solar_elongation",
          "radial_velocity", "d_ra_dt", "d_dec_dt")


@pytest.fixture(scope="module")
def eph_env(oracle):
    from outfit_b200 import synth
    table = synth.make_ephemeris_table()
    et = oracle.make_ephem_table(table["cheb"], table["jd_start"], table["block_days"], table["ipt"], table["emrat"])
    return dict(O=oracle, synth=synth, table=table, et=et)


def test_oracle_ephemeris_properties(eph_env):
    """Rates agree with finite differences of the positions, the triangle Sun-observer-body closes,
    per-orbit conversion errors mark every epoch (ephemeris/mod.rs:196-240)."""
    O, synth, et = eph_env["O"], eph_env["synth"], eph_env["et"]
    kind, epoch, elem = synth.make_ephemeris_orbits(400, seed=5, mixed_kinds=True)
    h = 1.0 / 1024.0
    tt, ut1, bf = synth.make_ephemeris_epochs(41, step=h)
    bf = bf * 0.0  # geocentre: the reference's observer velocity is the Earth's only (apparent_position.rs:283-293)
    out, st = O.ephemeris_twobody_batch(et, kind, epoch, elem, tt, ut1, bf, n_threads=2)
    out2, st2 = O.ephemeris_twobody_batch(et, kind, epoch, elem, tt, ut1, bf, n_threads=2, dedup_observer=True)
    assert np.array_equal(st, st2) and np.array_equal(out[:, st == 0], out2[:, st2 == 0])
    bad_orbit = (kind == 2)                      # hyperbolic / parabolic cometary input
    assert (st[:, bad_orbit] == 9).all() and (st[:, ~bad_orbit] == 0).all()
    assert np.isnan(out[:, :, bad_orbit]).all()
    good = ~bad_orbit
    ra, dec, geo, helio, phase, elong, rv, dra, ddec = (out[q][:, good] for q in range(9))
    # central differences at the interior epochs
    fd = lambda x: (x[2:] - x[:-2]) / (2 * h)
    dra_fd = (np.unwrap(ra, axis=0)[2:] - np.unwrap(ra, axis=0)[:-2]) / (2 * h)
    # (light-time derivative terms ~ v/c = 1e-4 relative are not part of the reference's rates)
    assert (np.abs(dra_fd - dra[1:-1]) < 1e-3 * np.abs(dra[1:-1]) + 2e-6).all()
    assert (np.abs(fd(dec) - ddec[1:-1]) < 1e-3 * np.abs(ddec[1:-1]) + 2e-6).all()
    # law of cosines in the Sun-observer-body triangle: r_h^2 = r_o^2 + rho^2 - 2 r_o rho cos(elong)
    # (aberration-corrected rho is not output; use the phase/elongation/third-angle sum instead)
    third = np.pi - phase - elong
    assert (third > -1e-6).all()
    assert (np.abs(dec) <= np.pi / 2).all() and (ra >= 0).all() and (ra < 2 * np.pi).all()
    assert (geo > 0).all() and (helio > 0).all()


def test_oracle_ephemeris_epoch_out_of_table(eph_env):
    O, synth, et = eph_env["O"], eph_env["synth"], eph_env["et"]
    kind, epoch, elem = synth.make_ephemeris_orbits(8, seed=6)
    tt, ut1, bf = synth.make_ephemeris_epochs(3, mjd0=40000.0)   # before the synthetic table starts
    out, st = O.ephemeris_twobody_batch(et, kind, epoch, elem, tt, ut1, bf, n_threads=1)
    assert (st == 17).all() and np.isnan(out).all()


@pytest.mark.gpu
@pytest.mark.parametrize("n_orbits,n_epochs", [(3000, 37), (1000, 300), (5, 1), (129, 128)])
def test_gpu_ephemeris_matches_oracle(eph_env, n_orbits, n_epochs):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from outfit_b200 import OutfitB200
    O, synth, et = eph_env["O"], eph_env["synth"], eph_env["et"]
    ctx = OutfitB200(0)
    ctx.load_ephemeris(eph_env["table"])
    kind, epoch, elem = synth.make_ephemeris_orbits(n_orbits, seed=7 + n_orbits, mixed_kinds=True)
    tt, ut1, bf = synth.make_ephemeris_epochs(n_epochs, site_idx=2)
    want, wst = O.ephemeris_twobody_batch(et, kind, epoch, elem, tt, ut1, bf)
    got, gst = ctx.ephemeris_twobody(kind, epoch, elem, tt, ut1, bf)
    assert np.array_equal(gst, wst)
    ok = wst == 0
    assert np.isnan(got[:, ~ok]).all()
    # angles: absolute 1e-11 rad (libm differences, 2e-6 arcsec); distances / rates: 1e-10 relative
    dang = np.abs((got[0][ok] - want[0][ok] + np.pi) % (2 * np.pi) - np.pi)
    assert dang.max() < 1e-11
    for q in (1, 4, 5):
        assert np.abs(got[q][ok] - want[q][ok]).max() < 1e-11, FIELDS[q]
    for q in (2, 3):
        assert (np.abs(got[q][ok] - want[q][ok]) / np.abs(want[q][ok])).max() < 1e-12, FIELDS[q]
    for q in (6, 7, 8):
        assert np.abs(got[q][ok] - want[q][ok]).max() < 1e-12, FIELDS[q]


@pytest.mark.gpu
def test_gpu_ephemeris_from_iod_results(eph_env):
    """The (element_kind, epoch, elem) of the IOD results feed the ephemeris entry directly."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from outfit_b200 import IODParams, OutfitB200
    O, synth, et = eph_env["O"], eph_env["synth"], eph_env["et"]
    ctx = OutfitB200(0)
    ctx.load_ephemeris(eph_env["table"])
    batch = synth.make_trajectories(200, 12, seed=31, table=eph_env["table"], max_triplets=10, n_noise=1)
    res = ctx.fit_full_iod(batch, IODParams.builder(n_noise_realizations=0, max_triplets=10))
    ok = res["status"] == 0
    kind = np.ascontiguousarray(res["element_kind"][ok].astype(np.int32))
    epoch = np.ascontiguousarray(res["epoch"][ok])
    elem = np.ascontiguousarray(res["elem"][ok].T)
    tt, ut1, bf = synth.make_ephemeris_epochs(10, mjd0=float(np.median(batch["mjd_tt"])))
    got, gst = ctx.ephemeris_twobody(kind, epoch, elem, tt, ut1, bf)
    want, wst = O.ephemeris_twobody_batch(et, kind, epoch, elem, tt, ut1, bf)
    assert np.array_equal(gst, wst)
    m = wst == 0
    assert m.sum() > 0.8 * m.size
    assert np.abs((got[0][m] - want[0][m] + np.pi) % (2 * np.pi) - np.pi).max() < 1e-10
    assert np.abs(got[1][m] - want[1][m]).max() < 1e-10


@pytest.mark.gpu
def test_compute_ephemerides_over_fit_results(eph_env):
    """FullOrbitResultExt::compute_ephemerides (ephemeris/batch.rs:134-183): one call over the whole result array of
    fit_full_iod / fit_lsq; failed fits are InvalidConversion entries, the others equal the per-orbit entry."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from outfit_b200 import DifferentialCorrectionConfig, IODParams, OutfitB200
    synth = eph_env["synth"]
    ctx = OutfitB200(0)
    ctx.load_ephemeris(eph_env["table"])
    batch = synth.make_trajectories(300, (3, 14), seed=33, table=eph_env["table"], max_triplets=10, n_noise=1)
    p = IODParams.builder(n_noise_realizations=0, max_triplets=10)
    res = ctx.fit_full_iod(batch, p)
    assert (res["status"] != 0).any() and (res["status"] == 0).sum() > 100
    tt, ut1, bf = synth.make_ephemeris_epochs(6, mjd0=float(np.median(batch["mjd_tt"])))
    observers = [(bf, tt, ut1), (np.zeros(3), tt[:2].copy(), ut1[:2].copy())]
    out, st = ctx.compute_ephemerides(res, observers)
    assert out.shape == (9, 8, 300) and st.shape == (8, 300)
    bad = res["status"] != 0
    assert (st[:, bad] == 9).all() and np.isnan(out[:, :, bad]).all()
    ok = ~bad
    kind = np.ascontiguousarray(res["element_kind"][ok].astype(np.int32))
    ref, rst = ctx.ephemeris_request(kind, np.ascontiguousarray(res["epoch"][ok]), np.ascontiguousarray(res["elem"][ok].T), observers)
    assert np.array_equal(st[:, ok], rst) and np.array_equal(out[:, :, ok], ref, equal_nan=True)
    # the LSQ records: corrected orbits are equinoctial, fallbacks take the IOD orbit and its element type
    lsq, _ = ctx.fit_lsq(batch, p, DifferentialCorrectionConfig.default(), initial_orbits=res)
    out2, st2 = ctx.compute_ephemerides(lsq, observers, iod_results=res)
    fb = lsq["kind"] == 2
    assert fb.any() and np.array_equal(out2[:, :, fb], out[:, :, fb], equal_nan=True) and np.array_equal(st2[:, fb], st[:, fb])
    cor = lsq["kind"] == 1
    good = (st2[:, cor] == 0) & (st[:, cor] == 0)
    d = np.abs((out2[0][:, cor][good] - out[0][:, cor][good] + np.pi) % (2 * np.pi) - np.pi)
    # the corrected orbit is another orbit through the same observations: close on the sky at the arc's epochs, not equal
    assert cor.sum() > 50 and good.mean() > 0.9 and 0.0 < np.median(d) < 0.05


@pytest.mark.gpu
def test_gpu_second_order_aberration_matches_oracle(eph_env):
    """EphemerisConfig::aberration = AberrationOrder::Second (ephemeris/aberration.rs:60-75, 195-234): the line of sight
    from two Keplerian back-propagations by the light time.  GPU == oracle at the first-order tolerances; the two
    orders differ by O((v/c)^2) ~ milliarcseconds, far above that tolerance, so the switch is observable."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from outfit_b200 import EphemerisConfig, OutfitB200, OutfitError
    O, synth, et = eph_env["O"], eph_env["synth"], eph_env["et"]
    ctx = OutfitB200(0)
    ctx.load_ephemeris(eph_env["table"])
    kind, epoch, elem = synth.make_ephemeris_orbits(4000, seed=77, mixed_kinds=True)
    tt, ut1, bf = synth.make_ephemeris_epochs(41, site_idx=2)
    first, fst = ctx.ephemeris_twobody(kind, epoch, elem, tt, ut1, bf)
    ctx.set_ephemeris_config(EphemerisConfig(aberration=2))
    got, gst = ctx.ephemeris_twobody(kind, epoch, elem, tt, ut1, bf)
    want, wst = O.ephemeris_twobody_batch(et, kind, epoch, elem, tt, ut1, bf, aberration_order=2)
    assert np.array_equal(gst, wst)
    ok = wst == 0
    assert ok.mean() > 0.9 and np.isnan(got[:, ~ok]).all()
    dang = np.abs((got[0][ok] - want[0][ok] + np.pi) % (2 * np.pi) - np.pi)
    assert dang.max() < 1e-11
    for q in (1, 4, 5):
        assert np.abs(got[q][ok] - want[q][ok]).max() < 1e-11, FIELDS[q]
    for q in (2, 3):  # distances do not depend on the aberration order
        assert np.array_equal(got[q][ok], first[q][ok])
    for q in (6, 7, 8):
        assert np.abs(got[q][ok] - want[q][ok]).max() < 1e-11 * np.maximum(1.0, np.abs(want[q][ok]).max()), FIELDS[q]
    d12 = np.abs((got[0][ok] - first[0][ok] + np.pi) % (2 * np.pi) - np.pi)
    assert 1e-10 < d12.max() < 1e-6  # milliarcsecond-level change, not a rounding difference
    # back to the default; NBody is refused, not silently replaced
    ctx.set_ephemeris_config(EphemerisConfig())
    again, _ = ctx.ephemeris_twobody(kind, epoch, elem, tt, ut1, bf)
    assert again.tobytes() == first.tobytes()
    with pytest.raises(OutfitError):
        ctx.set_ephemeris_config(EphemerisConfig(propagator=1))
