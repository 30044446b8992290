"""Reference unit / property tests that need neither DE440 nor UT1, restated against the oracle (CPU) and the
device:
  * HorizonRecord::interpolate      src/jpl_ephem/horizon/horizon_records.rs:356-520
  * Newton / Brent-Dekker solvers   src/kepler/params.rs:264-587 (unit tests + proptests; the reference holds no
                                    test module inside brent_dekker_solver.rs itself)
"""
import ctypes as C

import numpy as np
import pytest

MU_SUN = 2.959122082855911e-4  # params.rs:209


# ------------------------------------------------------------------------------------------------
# HorizonRecord::interpolate
# ------------------------------------------------------------------------------------------------
def cheb_basis(tc, n):
    t = [1.0, tc]
    for _ in range(2, n):
        t.append(2.0 * tc * t[-1] - t[-2])
    return np.array(t[:n])


def tc_from_tau(tau, n_sub):  # horizon_records.rs:349-353
    dt1 = int(tau)
    temp = n_sub * tau
    return 2.0 * ((temp % 1.0) + dt1) - 1.0


def record(oracle, x, y, z, start, end, tau, n_sub, with_vel=True):
    co = np.ascontiguousarray(np.stack([x, y, z]).astype(np.float64))
    pos, vel = oracle.D3(), oracle.D3()
    oracle.lib().oo_cheb_record(co.ctypes.data, co.shape[1], float(tau), int(n_sub), float(end - start), 1 if with_vel else 0, pos, vel)
    return np.array(list(pos)), np.array(list(vel))


def test_record_zero_coefficients(oracle):  # :385-398
    z = np.zeros(6)
    for tau, ns in ((0.0, 1), (0.42, 5)):
        p, v = record(oracle, z, z, z, 1000.0, 1001.0, tau, ns)
        assert (p == 0).all() and (v == 0).all()


def test_record_position_matches_explicit_basis(oracle):  # :400-419
    x, y, z = [1.0, 0.5, -0.25, 0.125], [0.0, -2.0, 0.0, 1.0], [3.0, 0.0, 0.0, 0.0]
    for tau in (0.0, 0.1, 0.33, 0.5, 0.73, 0.999999):
        tc = tc_from_tau(tau, 4)
        p, _ = record(oracle, x, y, z, 2000.0, 2002.0, tau, 4, with_vel=False)
        b = cheb_basis(tc, 4)
        assert abs(p[0] - np.dot(x, b)) <= 1e-14 and abs(p[1] - np.dot(y, b)) <= 1e-14 and abs(p[2] - np.dot(z, b)) <= 1e-14


def test_record_velocity_flag_does_not_change_position(oracle):  # :421-441
    a = ([0.1, -0.2, 0.3, -0.4, 0.5], [1.0, 0.0, -1.0, 0.0, 1.0], [2.0, -1.0, 0.0, 0.0, 0.0])
    p0, _ = record(oracle, *a, 1234.5, 1236.5, 0.37, 3, with_vel=False)
    p1, _ = record(oracle, *a, 1234.5, 1236.5, 0.37, 3, with_vel=True)
    assert np.array_equal(p0, p1)


def test_record_pure_t1_gives_constant_velocity(oracle):  # :443-470
    a = 123.456789
    vfac = 2.0 * 8 / 4.0
    for tau in (0.0, 0.2, 0.4, 0.6, 0.9):
        _, v = record(oracle, [0.0, a, 0.0], [0.0, -a, 0.0], [0.0, 2 * a, 0.0], 1000.0, 1004.0, tau, 8)
        assert np.abs(v - vfac * np.array([a, -a, 2 * a])).max() <= 1e-12


def test_record_scaling_invariance_and_edges(oracle):  # :472-520
    s = 5.5
    co = (np.array([0.3, -0.2, 0.5, 0.0, 0.1]), np.array([1.0, 2.0, 3.0, 4.0, 5.0]), np.array([-1.0, 0.0, 1.0, -1.0, 0.0]))
    pa, va = record(oracle, *co, 5000.0, 5002.0, 0.73, 7)
    pb, vb = record(oracle, *(s * c for c in co), 5000.0, 5002.0, 0.73, 7)
    assert np.abs(pb - s * pa).max() <= 1e-12 and np.abs(vb - s * va).max() <= 1e-12
    for tau in (0.0, 1.0, np.finfo(float).eps, 1.0 - 1e-15):
        p, v = record(oracle, [1.0, -0.5, 0.25, -0.125], [0.0, 2.0, 0.0, -1.0], [3.0, -2.0, 1.0, 0.0], 70000.0, 70001.0, tau, 4)
        assert np.isfinite(p).all() and np.isfinite(v).all()


def test_record_linear_in_coefficients_property(oracle):  # proptest :522-560
    rng = np.random.default_rng(11)
    for _ in range(200):
        n = int(rng.integers(3, 12))
        tau, ns = float(rng.uniform(0, 1)), int(rng.integers(1, 9))
        a, b = rng.uniform(-1, 1, (3, n)), rng.uniform(-1, 1, (3, n))
        pa, _ = record(oracle, *a, 0.0, 2.0, tau, ns)
        pb, _ = record(oracle, *b, 0.0, 2.0, tau, ns)
        pab, _ = record(oracle, *(a + b), 0.0, 2.0, tau, ns)
        assert np.abs(pab - (pa + pb)).max() <= 1e-12


@pytest.mark.gpu
def test_device_chebyshev_equals_the_record_evaluation(oracle):
    """The device table read (dev_geometry.cuh / dev_ephemeris.cuh) on a one-block table whose EMB body is a random
    record and whose Moon / Sun are zero: Earth position and velocity == the record evaluation / AU."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from outfit_b200 import OutfitB200
    rng = np.random.default_rng(5)
    nc, ns, days = 13, 2, 32.0
    emb = rng.uniform(-1e8, 1e8, (ns, 3, nc)) * (0.5 ** np.arange(nc))
    blk = np.zeros((1, ns * 3 * nc + 2 * 3 * 3))
    blk[0, :ns * 3 * nc] = emb.ravel()
    ipt = np.array([[0, nc, ns], [ns * 3 * nc, 3, 1], [ns * 3 * nc + 9, 3, 1]], dtype=np.uint32)
    table = dict(cheb=blk, jd_start=2459000.5, block_days=days, ipt=ipt, emrat=81.3)
    ctx = OutfitB200(0)
    ctx.load_ephemeris(table)
    n = 257
    mjd = 59000.0 + np.sort(rng.uniform(0.0, days - 1e-6, n))
    mjd[0] = 59000.0
    kind = np.zeros(1, dtype=np.int32)
    out, st = ctx.ephemeris_twobody(kind, np.array([59000.0]), np.array([[2.5], [0.1], [0.1], [1.0], [1.0], [1.0]]), mjd, mjd,
                                    np.zeros(3))
    assert (st == 0).all()
    au = 149597870.7
    dev = torch.device("cuda", 0)
    geo = torch.empty(3 * n, dtype=torch.float64, device=dev)
    hel = torch.empty(3 * n, dtype=torch.float64, device=dev)
    ctx.observer_cache_device(n, torch.from_numpy(mjd).to(dev), torch.from_numpy(mjd).to(dev), torch.zeros(3 * n, dtype=torch.float64, device=dev), geo, hel)
    torch.cuda.synchronize()
    hel = hel.cpu().numpy().reshape(3, n)
    for i in range(n):
        et_jd = 2400000.5 + np.trunc(mjd[i])
        tau = ((et_jd - 2459000.5) + (mjd[i] - np.trunc(mjd[i]))) / days
        sub = min(int(np.floor(tau * ns)), ns - 1)
        p, _ = record(oracle, *emb[sub], 0.0, days, tau, ns)
        # a zero body-fixed observer sits at the geocentre: heliocentric observer = Earth (Moon = Sun = 0: EMB)
        assert np.abs(hel[:, i] - p / au).max() <= 4e-16 * np.abs(p / au).max() + 1e-300


# ------------------------------------------------------------------------------------------------
# universal Kepler solvers (params.rs tests)
# ------------------------------------------------------------------------------------------------
def kp(oracle, dt, r0, sig0, alpha, e0, kind):
    p = oracle.KeplerParams()
    oracle.lib().oo_kepler_params_default_solver(C.byref(p))
    p.dt, p.r0, p.sig0, p.mu, p.alpha, p.e0, p.kind = dt, r0, sig0, MU_SUN, alpha, e0, kind
    return p


def elliptic(oracle, dt, a, kind):  # params.rs:222-241
    return kp(oracle, dt, a, 0.0, -1.0 / a, 0.01, kind)


def hyperbolic(oracle, dt, c3, kind):  # :247-262
    return kp(oracle, dt, 1.5, 0.001, c3 / MU_SUN, 1.5, kind)


def solve(oracle, p):
    sol = oracle.KeplerSolution()
    rc = oracle.lib().oo_kepler_solve(C.byref(p), C.byref(sol))
    return rc, sol


def residual(oracle, sol, p):  # :214-219
    s = (C.c_double * 4)()
    oracle.lib().oo_s_funct(sol.psi, p.alpha, s)
    return abs(p.r0 * s[1] + p.sig0 * s[2] + s[3] - np.sqrt(p.mu) * p.dt)


@pytest.mark.parametrize("kind", [0, 1])
def test_solver_unit_cases(oracle, kind):  # :268-343 (Newton), :353-397 (Brent)
    for mk, a, b, tol in ((elliptic, 10.0, 1.0, 1e-10), (elliptic, 365.0, 1.0, 1e-9), (elliptic, -30.0, 1.0, 1e-10),
                          (hyperbolic, 5.0, 1e-5, 1e-10)):
        p = mk(oracle, a, b, kind)
        rc, sol = solve(oracle, p)
        assert rc == 0, (mk.__name__, a, b)
        assert residual(oracle, sol, p) < tol


def test_newton_warm_start_consistent_with_cold_start(oracle):  # :321-348
    p = elliptic(oracle, 50.0, 2.0, 0)
    rc, cold = solve(oracle, p)
    assert rc == 0
    p.has_psi_guess, p.psi_guess = 1, cold.psi
    rc, warm = solve(oracle, p)
    assert rc == 0 and abs(cold.psi - warm.psi) <= 1e-10


def test_solvers_consistent_unit_cases(oracle):  # :407-458, PSI_CONSISTENCY_TOL = 1e-12
    for mk, a, b in ((elliptic, 10.0, 1.0), (elliptic, 200.0, 2.5), (elliptic, -45.0, 1.5), (hyperbolic, 5.0, 1e-5),
                     (elliptic, 1000.0, 5.2)):
        rn, sn = solve(oracle, mk(oracle, a, b, 0))
        rb, sb = solve(oracle, mk(oracle, a, b, 1))
        assert rn == 0 and rb == 0
        assert residual(oracle, sn, mk(oracle, a, b, 0)) < 1e-10 and residual(oracle, sb, mk(oracle, a, b, 1)) < 1e-10
        assert abs(sn.psi - sb.psi) <= 1e-12


def test_solver_properties(oracle):  # proptests :483-560
    rng = np.random.default_rng(2026)
    for _ in range(2000):
        a, dt = rng.uniform(0.3, 10.0), rng.uniform(1.0, 500.0)
        sols = []
        for kind in (0, 1):
            p = elliptic(oracle, dt, a, kind)
            rc, sol = solve(oracle, p)
            if rc == 0:
                assert residual(oracle, sol, p) < 1e-9
                sols.append(sol.psi)
        if len(sols) == 2:
            assert abs(sols[0] - sols[1]) < 1e-12
    for _ in range(2000):
        c3, dt = rng.uniform(1e-6, 1e-3), rng.uniform(1.0, 100.0)
        rn, sn = solve(oracle, hyperbolic(oracle, dt, c3, 0))
        rb, sb = solve(oracle, hyperbolic(oracle, dt, c3, 1))
        if rn == 0 and rb == 0:
            assert abs(sn.psi - sb.psi) < 1e-12


@pytest.mark.gpu
def test_device_newton_and_brent_agree_like_the_reference_requires(oracle):
    """The same scenarios on the device: circular states of radius a (r0 = a, sig0 = 0, alpha = -1/a) and the
    hyperbolic template, SolverKind Newton vs BrentDecker through outfit_b200_propagate_universal: psi within
    1e-12 of each other (params.rs:403), residual below 1e-9, and both equal to the oracle's psi."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from outfit_b200 import OutfitB200, SolverType
    ctx = OutfitB200(0)
    rng = np.random.default_rng(7)
    n = 20000
    a = rng.uniform(0.3, 10.0, n)
    dt = rng.uniform(1.0, 500.0, n) * np.where(rng.uniform(size=n) < 0.2, -1.0, 1.0)
    rv = np.zeros((6, n))
    rv[0], rv[4] = a, np.sqrt(MU_SUN / a)
    # hyperbolic half: r0 = 1.5, sig0 = 0.001, alpha = c3 / mu
    h = slice(n // 2, n)
    c3 = rng.uniform(1e-6, 1e-3, n - n // 2)
    dt[h] = rng.uniform(1.0, 100.0, n - n // 2)
    r0 = 1.5
    vr = 0.001 * np.sqrt(MU_SUN) / r0
    v2 = c3 + 2.0 * MU_SUN / r0
    rv[0, h], rv[3, h], rv[4, h] = r0, vr, np.sqrt(v2 - vr * vr)
    t0 = np.full(n, 60000.0)
    t1 = t0 + dt
    on, sn = ctx.propagate_universal(rv, t0, t1, SolverType(kind=0))
    ob, sb = ctx.propagate_universal(rv, t0, t1, SolverType(kind=1))
    both = (sn == 0) & (sb == 0)
    assert both.mean() > 0.999
    assert np.abs(on[10][both] - ob[10][both]).max() < 1e-12
    want_n, wsn = oracle.propagate_universal_batch(rv, t0, t1, 0)
    want_b, wsb = oracle.propagate_universal_batch(rv, t0, t1, 1)
    assert np.array_equal(wsn, sn) and np.array_equal(wsb, sb)
    assert np.abs(on[10][both] - want_n[10][both]).max() < 1e-12 and np.abs(ob[10][both] - want_b[10][both]).max() < 1e-12
    # residual of the universal Kepler equation at the device's psi
    r0n = np.linalg.norm(rv[0:3], axis=0)
    sig0 = (rv[0:3] * rv[3:6]).sum(0) / np.sqrt(MU_SUN)
    alpha = ((rv[3:6] ** 2).sum(0) - 2 * MU_SUN / r0n) / MU_SUN
    s = (C.c_double * 4)()
    for i in np.flatnonzero(both)[::50]:
        oracle.lib().oo_s_funct(float(ob[10][i]), float(alpha[i]), s)
        assert abs(r0n[i] * s[1] + sig0[i] * s[2] + s[3] - np.sqrt(MU_SUN) * dt[i]) < 1e-9
