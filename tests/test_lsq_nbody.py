"""FitLSQ with DifferentialCorrectionConfig::propagator = PropagatorKind::NBody (differential_orbit_correction/
single_iteration.rs:186-191, ephemeris/observation_ephemeris.rs:452-486, orbit_type/equinoctial_element.rs:908-968).

The reference's N-body differential-correction tests (tests/test_diff_cor.rs) need DE440, and its DOP853 crate is not
vendored: like every N-body number of this repository the parity is at the TOLERANCE level.  The oracle's N-body
partials are pinned (a) against its own two-body partials when the Sun is the only perturber and (b) against a finite
difference of its own N-body propagation; the device against the oracle."""
import numpy as np
import pytest


def _setup(oracle, T, seed, n_obs=12):
    from outfit_b200 import synth
    table = synth.make_ephemeris_table()
    et = oracle.make_ephem_table(table["cheb"], table["jd_start"], table["block_days"], table["ipt"], table["emrat"])
    batch = synth.make_trajectories(T, n_obs, seed=seed, table=table, max_triplets=10, n_noise=1)
    ob = oracle.from_soa_batch(batch)
    iod = oracle.fit_full_iod(ob, et, oracle.default_iod_params(n_noise_realizations=0, max_triplets=10), n_threads=0)
    return table, et, batch, ob, iod


def _perturbers(oracle, T, seed, bodies):
    rng = np.random.default_rng(seed)
    radius = {0: 0.0, 3: 1.0, 4: 1.52, 5: 5.2, 6: 9.5}
    gm = np.array([oracle.planet_gm(b) for b in bodies])
    pos = np.zeros((len(bodies), 3, T))
    for j, b in enumerate(bodies):
        lon = rng.uniform(0, 2 * np.pi, T)
        pos[j, 0], pos[j, 1], pos[j, 2] = radius[b] * np.cos(lon), radius[b] * np.sin(lon), 0.02 * radius[b] * np.sin(3 * lon)
    return gm, np.ascontiguousarray(pos)


def test_oracle_nbody_lsq_with_the_sun_alone_is_the_twobody_lsq(oracle):
    _, et, _, ob, iod = _setup(oracle, 30, 77)
    cfg = oracle.default_lsq_config()
    a, af = oracle.fit_lsq(ob, et, cfg, iod, n_threads=0)
    gm, pos = _perturbers(oracle, 30, 1, (0,))
    b, bf = oracle.fit_lsq_nbody(ob, et, cfg, iod, gm, pos, n_threads=0)
    assert np.array_equal(a["kind"], b["kind"]) and np.array_equal(a["total_newton_iterations"], b["total_newton_iterations"])
    ok = a["kind"] == 1
    assert ok.sum() >= 10
    assert np.abs(a["elem"][ok] - b["elem"][ok]).max() < 1e-8      # integration tolerance 1e-12 through the normal equations
    assert np.abs(a["normalised_rms"][ok] / b["normalised_rms"][ok] - 1.0).max() < 1e-6
    assert np.array_equal(af["selection"], bf["selection"])
    # and a massive perturber moves the solution, a little
    gm2, pos2 = _perturbers(oracle, 30, 2, (0, 5))
    c, _ = oracle.fit_lsq_nbody(ob, et, cfg, iod, gm2, pos2, n_threads=0)
    both = ok & (c["kind"] == 1)
    d = np.abs(a["elem"][both] - c["elem"][both]).max()
    assert 1e-7 < d < 1e-1


def test_oracle_nbody_partials_match_a_finite_difference(oracle):
    """d(ra, dec)/d(elements) of compute_obs_and_partials_nbody against central differences of the predicted (ra, dec)."""
    import ctypes as C
    _, et, _, ob, iod = _setup(oracle, 6, 78)
    L = oracle.lib()
    t = int(np.flatnonzero(iod["status"] == 0)[0])
    el_in, eq = oracle.Elements(), oracle.Elements()
    el_in.kind, el_in.epoch = int(iod["element_kind"][t]), float(iod["epoch"][t])
    for j in range(6):
        el_in.e[j] = float(iod["elem"][t][j])
    assert L.oo_to_equinoctial(C.byref(el_in), C.byref(eq)) == 0
    gm, pos = _perturbers(oracle, 1, 3, (0, 5, 3))
    pert = (oracle.Perturber * 3)()
    for j in range(3):
        pert[j].gm = gm[j]
        for c in range(3):
            pert[j].pos[c] = pos[j, c, 0]
    tv = oracle.traj_view(ob, t)
    D6 = C.c_double * 6

    def radec(e):
        ra, dec, dr, dd = C.c_double(), C.c_double(), D6(), D6()
        rc = L.oo_obs_and_partials_nbody(C.byref(tv), C.c_size_t(tv.n - 1), C.byref(et), C.byref(e), pert, C.c_size_t(3),
                                         C.c_double(1e-13), C.c_double(1e-13), C.byref(ra), C.byref(dec), dr, dd)
        assert rc == 0
        return ra.value, dec.value, np.array(dr[:]), np.array(dd[:])

    _, _, dra, ddec = radec(eq)
    for j in range(6):
        h = 1e-6 * max(1.0, abs(eq.e[j]))
        ep, em = oracle.Elements(), oracle.Elements()
        C.memmove(C.byref(ep), C.byref(eq), C.sizeof(eq)); C.memmove(C.byref(em), C.byref(eq), C.sizeof(eq))
        ep.e[j] += h; em.e[j] -= h
        rp, dp, _, _ = radec(ep)
        rm, dm, _, _ = radec(em)
        fd_ra, fd_dec = (rp - rm) / (2 * h), (dp - dm) / (2 * h)
        # 1e-3: the reference's chain rule leaves out d(velocity)/d(elements) in the aberration term
        # (observation_ephemeris.rs:246-258), a ~1e-4 relative effect that a finite difference sees
        scale = max(np.abs(dra).max(), np.abs(ddec).max())
        assert abs(fd_ra - dra[j]) <= 1e-3 * scale, (j, fd_ra, dra[j])
        assert abs(fd_dec - ddec[j]) <= 1e-3 * scale, (j, fd_dec, ddec[j])


@pytest.mark.gpu
def test_gpu_nbody_lsq_matches_oracle(oracle):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from outfit_b200 import DifferentialCorrectionConfig, OutfitB200, RESULT_DTYPE
    T = 240
    table, et, batch, ob, iod = _setup(oracle, T, 401)
    # one outlier per fifth trajectory: the rejection pass (partials at the last linearisation point) runs too
    rng = np.random.default_rng(5)
    off = batch["traj_offset"].astype(np.int64)
    for t in range(0, T, 5):
        i = off[t] + int(rng.integers(0, off[t + 1] - off[t]))
        batch["dec"][i] += 40.0 * batch["sigma_dec"][i]
    ob = oracle.from_soa_batch(batch)
    gm, pos = _perturbers(oracle, T, 9, (0, 5, 6, 3))
    want, wfit = oracle.fit_lsq_nbody(ob, et, oracle.default_lsq_config(), iod, gm, pos, n_threads=0)
    ctx = OutfitB200(0)
    ctx.load_ephemeris(table)
    got, gfit = ctx.fit_lsq_nbody(batch, iod.view(RESULT_DTYPE), gm, pos, DifferentialCorrectionConfig.default())
    assert np.array_equal(got["status"], want["status"])
    same_kind = got["kind"] == want["kind"]
    assert same_kind.mean() >= 0.97, np.flatnonzero(~same_kind)
    ok = same_kind & (want["kind"] == 1)
    assert ok.sum() > 60
    d = np.abs(got["elem"][ok] - want["elem"][ok]).max(axis=1)
    r = np.abs(got["normalised_rms"][ok] / want["normalised_rms"][ok] - 1.0)
    print("nbody lsq: corrected", int(ok.sum()), "median elem diff", np.median(d), "max", d.max(), "rms rel", r.max())
    assert np.median(d) < 1e-9 and np.quantile(d, 0.98) < 1e-6 and np.quantile(r, 0.98) < 1e-5
    fb = same_kind & (want["kind"] == 2)
    assert np.array_equal(got["elem"][fb], want["elem"][fb]) and np.array_equal(got["fallback_cause"][fb], want["fallback_cause"][fb])
    o_ok = np.repeat(ok, np.diff(off))
    assert (gfit["selection"][o_ok] == wfit["selection"][o_ok]).mean() > 0.995
    assert (wfit["selection"] == 1).sum() > 10, "no observation was rejected: the rejection pass is not covered"
    iters_equal = got["total_newton_iterations"][ok] == want["total_newton_iterations"][ok]
    assert iters_equal.mean() > 0.95


@pytest.mark.gpu
def test_gpu_nbody_lsq_with_the_sun_alone_is_the_twobody_lsq(oracle):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from outfit_b200 import DifferentialCorrectionConfig, OutfitB200, RESULT_DTYPE
    T = 400
    table, et, batch, ob, iod = _setup(oracle, T, 402)
    ctx = OutfitB200(0)
    ctx.load_ephemeris(table)
    cfg = DifferentialCorrectionConfig.default()
    a, af = ctx.fit_lsq(batch, None, cfg, initial_orbits=iod.view(RESULT_DTYPE))
    gm, pos = _perturbers(oracle, T, 1, (0,))
    b, bf = ctx.fit_lsq_nbody(batch, iod.view(RESULT_DTYPE), gm, pos, cfg)
    assert np.array_equal(a["status"], b["status"]) and (a["kind"] == b["kind"]).mean() > 0.99
    ok = (a["kind"] == 1) & (b["kind"] == 1)
    assert ok.sum() > 100
    assert np.quantile(np.abs(a["elem"][ok] - b["elem"][ok]).max(axis=1), 0.98) < 1e-7
    # the failed-IOD trajectories pass through untouched in both
    bad = a["kind"] == 0
    assert np.array_equal(a[bad], b[bad])


@pytest.mark.gpu
def test_group_nbody_lsq_equals_one_context(oracle):
    """outfit_b200_group_fit_lsq_nbody over several contexts (all GPUs of the box, or three contexts on one GPU): the
    records at their global indices, bit-identical to the single-context call."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from outfit_b200 import DifferentialCorrectionConfig, NBodyConfig, OutfitB200, OutfitGroup, RESULT_DTYPE, synth
    T = 150
    table = synth.make_ephemeris_table()
    batch = synth.make_trajectories(T, (8, 20), seed=403, table=table, max_triplets=10, n_noise=1)
    ctx = OutfitB200(0)
    ctx.load_ephemeris(table)
    from outfit_b200 import IODParams
    iod = ctx.fit_full_iod(batch, IODParams.builder(n_noise_realizations=0, max_triplets=10))
    gm, pos = _perturbers(oracle, T, 4, (0, 5, 3))
    cfg, nb = DifferentialCorrectionConfig.default(), NBodyConfig(n_perturbers=3, max_steps=2000)
    one, ofit = ctx.fit_lsq_nbody(batch, iod, gm, pos, cfg, nb)
    n_dev = torch.cuda.device_count()
    grp = OutfitGroup(list(range(n_dev)) if n_dev > 1 else [0, 0, 0])
    grp.load_ephemeris(table)
    many, mfit = grp.fit_lsq_nbody(batch, iod, gm, pos, cfg, nb)
    assert (one["kind"] == 1).sum() > 30
    assert one.tobytes() == many.tobytes() and ofit.tobytes() == mfit.tobytes()
    grp.close()


@pytest.mark.gpu
def test_gpu_nbody_lsq_with_on_device_observer_geometry(oracle):
    """Body-fixed observer coordinates + UT1 (pvobs on the device) instead of a precomputed cache: parity with the oracle
    fed by the oracle's own pvobs / Earth positions, from the same initial orbits."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from outfit_b200 import DifferentialCorrectionConfig, IODParams, NBodyConfig, OutfitB200, synth
    from parity_util import oracle_observer_cache
    T = 120
    table = synth.make_ephemeris_table()
    et = oracle.make_ephem_table(table["cheb"], table["jd_start"], table["block_days"], table["ipt"], table["emrat"])
    batch = synth.make_trajectories(T, 12, seed=404, table=table, max_triplets=10, n_noise=1)
    ctx = OutfitB200(0)
    ctx.load_ephemeris(table)
    iod = ctx.fit_full_iod(batch, IODParams.builder(n_noise_realizations=0, max_triplets=10), use_body_fixed=True)
    gm, pos = _perturbers(oracle, T, 6, (0, 5))
    got, gfit = ctx.fit_lsq_nbody(batch, iod, gm, pos, DifferentialCorrectionConfig.default(), NBodyConfig(n_perturbers=2),
                                  use_body_fixed=True)
    hel, geo = oracle_observer_cache(oracle, et, batch)
    ob = oracle.from_soa_batch(batch)
    ob["helio_equ"], ob["geo_ecl"] = hel, geo
    oiod = np.ascontiguousarray(iod.view(oracle.IOD_RESULT_DTYPE))
    want, wfit = oracle.fit_lsq_nbody(ob, et, oracle.default_lsq_config(), oiod, gm, pos, n_threads=0)
    assert np.array_equal(got["status"], want["status"]) and (got["kind"] == want["kind"]).mean() > 0.97
    ok = (got["kind"] == 1) & (want["kind"] == 1)
    assert ok.sum() > 40
    d = np.abs(got["elem"][ok] - want["elem"][ok]).max(axis=1)
    assert np.median(d) < 1e-9 and np.quantile(d, 0.95) < 1e-6, (np.median(d), d.max())
