"""N-body propagator: propagator/nbody.rs + EquinoctialElements::propagate_nbody
(orbit_type/equinoctial_element.rs:908-968).  The reference's own N-body tests need DE440 (tests/test_ephemeris.rs,
tests/test_diff_cor.rs) and its integrator is an un-vendored crate, so the oracle is pinned against scipy's DOP853 -- a
published implementation of the same method -- and by physical properties; the device against the oracle."""
import numpy as np
import pytest

GM_SUN_K2 = 0.01720209895 ** 2


def _orbits(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.uniform(1.2, 3.5, n)
    e = rng.uniform(0.0, 0.4, n)
    inc = np.abs(rng.rayleigh(np.radians(8.0), n))
    node, argp, M = (rng.uniform(0, 2 * np.pi, n) for _ in range(3))
    elem = np.ascontiguousarray(np.stack([a, e, inc, node, argp, M]))
    kind = np.zeros(n, dtype=np.int32)
    epoch = np.full(n, 59000.0)
    t1 = epoch + rng.uniform(-60.0, 120.0, n)
    return kind, epoch, elem, t1


def _perturbers(oracle, n, seed, bodies=(0, 5, 6, 3)):
    """Sun at the origin + a few planets at plausible heliocentric positions (frozen snapshots), per orbit."""
    rng = np.random.default_rng(seed)
    radius = {0: 0.0, 1: 0.39, 2: 0.72, 3: 1.0, 4: 1.52, 5: 5.2, 6: 9.5, 7: 19.2, 8: 30.0}
    gm = np.array([oracle.planet_gm(b) for b in bodies])
    pos = np.zeros((len(bodies), 3, n))
    for j, b in enumerate(bodies):
        lon = rng.uniform(0, 2 * np.pi, n)
        pos[j, 0], pos[j, 1], pos[j, 2] = radius[b] * np.cos(lon), radius[b] * np.sin(lon), 0.02 * radius[b] * np.sin(3 * lon)
    return gm, np.ascontiguousarray(pos)


def test_planet_gm_table(oracle):
    # planet_gm.rs tests: GM_SUN within 1e-4 of k^2 (in fact 5e-12), Jupiter >> Mars
    assert abs(oracle.planet_gm(0) - GM_SUN_K2) / GM_SUN_K2 < 1e-9
    assert oracle.planet_gm(5) > 100 * oracle.planet_gm(4) and np.isnan(oracle.planet_gm(11))


def test_rhs_matches_a_numpy_statement(oracle):
    rng = np.random.default_rng(3)
    y = rng.normal(size=42)
    y[0:3] = [1.9, -0.7, 0.2]
    gm = np.array([oracle.planet_gm(0), oracle.planet_gm(5)])
    pos = np.array([[0.0, 0.0, 0.0], [4.1, 3.0, -0.1]])
    dy = oracle.nbody_rhs(y, gm, pos)
    acc, G = np.zeros(3), np.zeros((3, 3))
    for g, p in zip(gm, pos):
        d = y[0:3] - p
        r = np.linalg.norm(d)
        acc += -g / r ** 3 * d + (g / np.linalg.norm(p) ** 3 * p if np.linalg.norm(p) > 1e-10 else 0.0)
        G += -g * (np.eye(3) / r ** 3 - 3 * np.outer(d, d) / r ** 5)
    A = np.zeros((6, 6))
    A[0:3, 3:6], A[3:6, 0:3] = np.eye(3), G
    want = np.concatenate([y[3:6], acc, (A @ y[6:].reshape(6, 6).T).T.ravel()])
    assert np.allclose(dy, want, rtol=1e-14, atol=1e-18)


def test_oracle_dop853_against_scipy(oracle):
    """The same right-hand side through scipy.integrate's DOP853 (rtol = atol = 1e-12): states agree to 1e-11, and the
    oracle takes the same number of steps (its controller is scipy's, up to summation order)."""
    from scipy.integrate import solve_ivp
    kind, epoch, elem, t1 = _orbits(12, seed=5)
    gm, pos = _perturbers(oracle, 12, seed=6)
    state, stm, status, steps = oracle.propagate_nbody(kind, epoch, elem, t1, gm, pos)
    assert (status == 0).all()
    zero = np.zeros_like(pos)
    s0, _, st0, _ = oracle.propagate_nbody(kind, epoch, elem, epoch, gm, zero)  # span 0: the initial state
    assert (st0 == 0).all()
    for i in range(12):
        y0 = np.concatenate([s0[:, i], np.eye(6).T.ravel()])
        sol = solve_ivp(lambda t, y: oracle.nbody_rhs(y, gm, pos[:, :, i]), (0.0, t1[i] - epoch[i]), y0, method="DOP853",
                        rtol=1e-12, atol=1e-12)
        assert sol.success
        assert np.abs(sol.y[0:6, -1] - state[:, i]).max() < 1e-11
        assert np.abs(sol.y[6:, -1] - stm[:, i]).max() < 1e-8 * max(1.0, np.abs(stm[:, i]).max())
        assert abs(int(steps[i]) - (len(sol.t) - 1)) <= 1


def test_sun_only_equals_two_body_and_stm_is_the_state_jacobian(oracle):
    """NBodyConfig::default() (perturbers = [Sun], mod.rs:139-150): DOP853 on the two-body problem == the analytic
    propagator at the tolerance level (GM_SUN of DE440 differs from k^2 by 5e-12); Phi(t1, t0) == finite differences."""
    import ctypes as C
    kind, epoch, elem, t1 = _orbits(20, seed=8)
    gm = np.array([oracle.planet_gm(0)])
    pos = np.zeros((1, 3, 20))
    state, stm, status, steps = oracle.propagate_nbody(kind, epoch, elem, t1, gm, pos)
    assert (status == 0).all() and (steps > 2).all() and (steps < 400).all()
    L = oracle.lib()
    for i in range(20):
        el = oracle.Elements()
        el.kind, el.epoch = 0, float(epoch[i])
        for q in range(6):
            el.e[q] = float(elem[q, i])
        eq = oracle.Elements()
        assert L.oo_to_equinoctial(C.byref(el), C.byref(eq)) == 0
        p, v = oracle.D3(), oracle.D3()
        assert L.oo_propagate_twobody(C.byref(eq), 0.0, float(t1[i] - epoch[i]), p, v) == 0
        assert np.abs(np.array(list(p)) - state[0:3, i]).max() < 2e-9 and np.abs(np.array(list(v)) - state[3:6, i]).max() < 2e-10
    # STM by central differences of the initial Cartesian state through the same integrator
    i = 3
    s0, _, _, _ = oracle.propagate_nbody(kind[i:i + 1], epoch[i:i + 1], elem[:, i:i + 1], epoch[i:i + 1], gm, pos[:, :, i:i + 1])
    import ctypes
    fd = np.zeros((6, 6))
    L.oo_dop853_nbody.argtypes = [C.c_void_p, C.c_double, C.POINTER(oracle.Perturber), C.c_size_t, C.c_double, C.c_double,
                                  C.c_uint32, C.c_void_p, C.c_void_p]
    per = oracle._perturbers(gm, pos[:, :, i])
    for c in range(6):
        ys = []
        for sgn in (+1, -1):
            y = np.concatenate([s0[:, 0], np.eye(6).ravel()])
            y[c] += sgn * 1e-6
            assert L.oo_dop853_nbody(y.ctypes.data, float(t1[i] - epoch[i]), per, 1, 1e-13, 1e-13, 100000, None, None) == 0
            ys.append(y[0:6].copy())
        fd[:, c] = (ys[0] - ys[1]) / 2e-6
    phi = stm[:, i].reshape(6, 6).T  # column-major storage
    assert np.abs(phi - fd).max() < 1e-5 * max(1.0, np.abs(phi).max())


def test_perturbers_move_the_orbit_and_failures_are_values(oracle):
    kind, epoch, elem, t1 = _orbits(8, seed=9)
    t1 = epoch + 300.0
    gm, pos = _perturbers(oracle, 8, seed=10, bodies=(0, 5))
    a, _, sa, _ = oracle.propagate_nbody(kind, epoch, elem, t1, gm, pos)
    b, _, sb, _ = oracle.propagate_nbody(kind, epoch, elem, t1, gm[:1], pos[:1])
    assert (sa == 0).all() and (sb == 0).all()
    d = np.linalg.norm(a[0:3] - b[0:3], axis=0)
    assert (d > 1e-7).all() and (d < 1e-2).all()  # Jupiter over 300 d: 1e-6 .. 1e-3 AU
    # a parabolic cometary orbit cannot be converted: InvalidConversion as a value
    k2 = kind.copy(); k2[0] = 2
    e2 = elem.copy(); e2[1, 0] = 1.0
    _, _, s2, _ = oracle.propagate_nbody(k2, epoch, e2, t1, gm, pos)
    assert s2[0] == 9 and (s2[1:] == 0).all()


@pytest.mark.gpu
def test_gpu_nbody_matches_oracle(oracle):
    """outfit_b200_propagate_nbody (eight lanes per orbit, DOP853 on [r, v, Phi]) against the oracle: state within 1e-10,
    STM within 1e-8 relative, the same status values, step counts within a few (the error norm is reduced in another
    order); mixed element kinds, forward and backward spans, a zero span, failures as values."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from outfit_b200 import NBodyConfig, OutfitB200, planet_gm, synth
    ctx = OutfitB200(0)
    n = 3000
    kind, epoch, elem = synth.make_ephemeris_orbits(n, seed=31, mixed_kinds=True)
    rng = np.random.default_rng(32)
    t1 = epoch + rng.uniform(-80.0, 150.0, n)
    t1[:8] = epoch[:8]                    # zero span: the initial state, Phi = I
    gm, pos = _perturbers(oracle, n, seed=33, bodies=(0, 5, 6, 3, 2))
    assert [planet_gm(b) for b in (0, 5, 6, 3, 2)] == list(gm)
    got, gstm, gst, gsteps = ctx.propagate_nbody(kind, epoch, elem, t1, gm, pos)
    want, wstm, wst, wsteps = oracle.propagate_nbody(kind, epoch, elem, t1, gm, pos)
    assert np.array_equal(gst, wst)
    ok = wst == 0
    assert ok.mean() > 0.9 and np.isnan(got[:, ~ok]).all()
    assert np.abs(got[:, ok] - want[:, ok]).max() < 1e-10
    scale = np.maximum(1.0, np.abs(wstm[:, ok]).max(axis=0))
    assert (np.abs(gstm[:, ok] - wstm[:, ok]).max(axis=0) / scale).max() < 1e-8
    assert np.abs(gsteps[ok].astype(int) - wsteps[ok].astype(int)).max() <= 3
    z = np.flatnonzero(ok[:8])
    assert len(z) > 0 and (gsteps[z] == 0).all() and all(np.array_equal(gstm[:, i].reshape(6, 6), np.eye(6)) for i in z)
    # NBodyConfig::default(): the Sun alone == the analytic two-body propagation at the tolerance level
    sun, _, sst, _ = ctx.propagate_nbody(kind, epoch, elem, t1, gm[:1], pos[:1], with_stm=False)
    osun, _, _, _ = oracle.propagate_nbody(kind, epoch, elem, t1, gm[:1], pos[:1])
    assert np.array_equal(sst, wst) and np.abs(sun[:, ok] - osun[:, ok]).max() < 1e-10
    # looser tolerances take fewer steps; a step budget that is too small is a per-orbit failure value
    _, _, _, loose = ctx.propagate_nbody(kind[:200], epoch[:200], np.ascontiguousarray(elem[:, :200]), t1[:200], gm,
                                         np.ascontiguousarray(pos[:, :, :200]), NBodyConfig(abs_tol=1e-8, rel_tol=1e-8))
    assert loose[ok[:200]].mean() < gsteps[:200][ok[:200]].mean()
    _, _, bst, _ = ctx.propagate_nbody(kind[:200], epoch[:200], np.ascontiguousarray(elem[:, :200]), t1[:200], gm,
                                       np.ascontiguousarray(pos[:, :, :200]), NBodyConfig(max_steps=2))
    assert (bst[ok[:200] & (gsteps[:200] > 2)] == 21).all()


@pytest.mark.gpu
@pytest.mark.parametrize("aberration", [1, 2])
def test_gpu_nbody_ephemeris_matches_oracle(oracle, aberration):
    """OrbitalElements::compute::<Combined> with PropagatorKind::NBody: every (orbit, epoch) entry integrated on its own
    from the orbit's epoch, then the same observer / aberration / geometry code as the two-body entries."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from outfit_b200 import EphemerisConfig, OutfitB200, synth
    table = synth.make_ephemeris_table()
    et = oracle.make_ephem_table(table["cheb"], table["jd_start"], table["block_days"], table["ipt"], table["emrat"])
    ctx = OutfitB200(0)
    ctx.load_ephemeris(table)
    n, E = 400, 9
    kind, epoch, elem = synth.make_ephemeris_orbits(n, seed=41, mixed_kinds=True)
    tt, ut1, bf = synth.make_ephemeris_epochs(E, step=7.0, site_idx=2)
    gm, pos = _perturbers(oracle, n, seed=42, bodies=(0, 5, 3))
    ctx.set_ephemeris_config(EphemerisConfig(aberration=aberration))
    got, gst = ctx.ephemeris_nbody(kind, epoch, elem, [(bf, tt, ut1)], gm, pos)
    want, wst = oracle.ephemeris_nbody_batch(et, kind, epoch, elem, tt, ut1, bf, gm, pos, aberration_order=aberration)
    assert np.array_equal(gst, wst)
    ok = wst == 0
    assert ok.mean() > 0.9 and np.isnan(got[:, ~ok]).all()
    dang = np.abs((got[0][ok] - want[0][ok] + np.pi) % (2 * np.pi) - np.pi)
    assert dang.max() < 1e-10
    for q in (1, 4, 5):
        assert np.abs(got[q][ok] - want[q][ok]).max() < 1e-10, FIELDS[q]
    for q in (2, 3):
        assert (np.abs(got[q][ok] - want[q][ok]) / np.abs(want[q][ok])).max() < 1e-10, FIELDS[q]
    for q in (6, 7, 8):
        assert np.abs(got[q][ok] - want[q][ok]).max() < 1e-10, FIELDS[q]
    # the perturbers are visible against the two-body ephemeris, far above the tolerance: in the sky position with the
    # first-order aberration; with the second-order one the line of sight comes from two-body back-propagations whatever
    # the main propagator is (reference behaviour, aberration.rs:176-178), so only the distances carry the N-body state
    two, _ = ctx.ephemeris_twobody(kind, epoch, elem, tt, ut1, bf)
    d = np.abs((got[0][ok] - two[0][ok] + np.pi) % (2 * np.pi) - np.pi)
    if aberration == 1:
        assert d.max() > 1e-8
    else:
        assert d.max() < 1e-11 and np.abs(got[3][ok] - two[3][ok]).max() > 1e-8


FIELDS = ("ra", "dec", "geocentric_dist", "heliocentric_dist", "phase_angle", "solar_elongation", "radial_velocity", "d_ra_dt", "d_dec_dt")
