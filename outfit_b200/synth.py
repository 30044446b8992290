"""Seeded synthetic inputs for the batched-IOD hot path (numpy only; no CUDA).

What is generated:
  * a DE440-SHAPED Chebyshev table (EMB 13 coefficients x 2 sub-intervals, Moon 13 x 8, Sun 11 x 2
    per 32-day block -- the IPT shape pinned by reference horizon_data.rs:863-893) fitted to an
    analytic Sun / Earth-Moon-barycentre / Moon model, because the real DE440 file is not in the
    image (the reference downloads it);
  * trajectories: truth orbits -> observation times -> observer positions -> RA/Dec through a
    two-body forward model with first-order aberration -> Gaussian astrometric noise;
  * random heliocentric states + times of flight for the bulk Kepler propagator, following the
    distributions of the reference's own proptest strategies (kepler/propagation.rs:939-1000).

All arrays are C-contiguous float64 / uint64, in the layout the C-ABI takes (include/outfit_b200.h):
per-observation 3-vectors are plane-major SoA, shape (3, n_obs).
"""
import numpy as np

GAUSS_GRAV = 0.01720209895
MU = GAUSS_GRAV * GAUSS_GRAV
AU_KM = 149597870.7
VLIGHT_AU = 2.99792458e5 / AU_KM * 86400.0
EMRAT = 81.30056822149722
OBL = 0.40909280422232897  # mean obliquity J2000 (rad)
COS_OBL, SIN_OBL = 9.174820620691818e-1, 3.977771559319137e-1  # reference constants.rs:93-121
ARCSEC = np.pi / 648000.0

# DE440 IPT shape for the three bodies the observer position needs; offsets are 0-based into OUR
# compact 456-double block (EMB | Moon | Sun), not into the 1018-double DE record.
IPT = np.array([[0, 13, 2], [78, 13, 8], [390, 11, 2]], dtype=np.uint32)
BLOCK_DOUBLES = 456
BLOCK_DAYS = 32.0


def ecl_to_equ(v):
    """Rotate (3, n) ecliptic J2000 -> equatorial J2000."""
    x, y, z = v
    return np.stack([x, COS_OBL * y - SIN_OBL * z, SIN_OBL * y + COS_OBL * z])


def equ_to_ecl(v):
    x, y, z = v
    return np.stack([x, COS_OBL * y + SIN_OBL * z, -SIN_OBL * y + COS_OBL * z])


# ------------------------------------------------------------------------------------------------
# analytic solar-system model (barycentric, equatorial J2000, km)
# ------------------------------------------------------------------------------------------------
def _kepler_E(M, e, iters=12):
    E = M + e * np.sin(M)
    for _ in range(iters):
        E = E - (E - e * np.sin(E) - M) / (1.0 - e * np.cos(E))
    return E


def _model_bodies(jd):
    """Return (emb, moon_geocentric, sun) each (3, n) in km, equatorial J2000, at JD (TDB~TT)."""
    t = jd - 2451545.0
    # Sun's barycentric wobble (Jupiter + Saturn terms, in the ecliptic plane)
    lj = 0.600 + 2 * np.pi * t / 4332.589
    ls = 0.870 + 2 * np.pi * t / 10759.22
    sun_ecl = np.stack([-7.43e5 * np.cos(lj) - 4.08e5 * np.cos(ls),
                        -7.43e5 * np.sin(lj) - 4.08e5 * np.sin(ls),
                        1.5e4 * np.sin(lj)])
    # EMB: heliocentric Kepler ellipse
    a, e, varpi = 1.00000011 * AU_KM, 0.01671022, 1.796767
    n = GAUSS_GRAV / (1.00000011 ** 1.5)
    M = (6.2400601 + n * t) % (2 * np.pi)
    E = _kepler_E(M, e)
    xp, yp = a * (np.cos(E) - e), a * np.sqrt(1 - e * e) * np.sin(E)
    emb_ecl = np.stack([xp * np.cos(varpi) - yp * np.sin(varpi),
                        xp * np.sin(varpi) + yp * np.cos(varpi),
                        np.zeros_like(t)]) + sun_ecl
    # geocentric Moon: inclined circle with a monthly + evection-like modulation
    lm = 3.8103 + 2 * np.pi * t / 27.321582
    node = 2.1824 - 2 * np.pi * t / 6798.38
    inc = 0.08980
    r = 384400.0 * (1.0 - 0.0549 * np.cos(lm - 1.4))
    u = lm - node
    moon_ecl = np.stack([r * (np.cos(node) * np.cos(u) - np.sin(node) * np.sin(u) * np.cos(inc)),
                         r * (np.sin(node) * np.cos(u) + np.cos(node) * np.sin(u) * np.cos(inc)),
                         r * np.sin(u) * np.sin(inc)])
    return ecl_to_equ(emb_ecl), ecl_to_equ(moon_ecl), ecl_to_equ(sun_ecl)


def _cheb_fit(func_samples, n_coeff):
    """func_samples: (..., n_nodes) values at Chebyshev-Gauss nodes (n_nodes >= n_coeff)."""
    n_nodes = func_samples.shape[-1]
    k = np.arange(n_nodes)
    theta = np.pi * (k + 0.5) / n_nodes
    j = np.arange(n_coeff)[:, None]
    Tm = np.cos(j * theta[None, :])  # (n_coeff, n_nodes)
    c = 2.0 / n_nodes * func_samples @ Tm.T
    c[..., 0] *= 0.5
    return c


def make_ephemeris_table(mjd_start=58000.0, n_blocks=125):
    """Synthetic DE440-shaped table.  Returns dict(cheb[n_blocks,456], jd_start, block_days, ipt, emrat)."""
    jd_start = 2400000.5 + mjd_start
    cheb = np.zeros((n_blocks, BLOCK_DOUBLES))
    n_nodes = 24
    xn = np.cos(np.pi * (np.arange(n_nodes) + 0.5) / n_nodes)  # nodes in [-1, 1]
    for b, (off, nc, nsub) in enumerate(IPT):
        off, nc, nsub = int(off), int(nc), int(nsub)
        sub_len = BLOCK_DAYS / nsub
        blk = np.arange(n_blocks)[:, None, None]
        sub = np.arange(nsub)[None, :, None]
        jd = jd_start + blk * BLOCK_DAYS + (sub + 0.5 * (xn[None, None, :] + 1.0)) * sub_len
        bodies = _model_bodies(jd.reshape(-1))
        vals = bodies[b].reshape(3, n_blocks, nsub, n_nodes)
        coef = _cheb_fit(vals, nc)  # (3, n_blocks, nsub, nc)
        # DE layout inside a body: [sub][axis][coeff]
        lay = np.transpose(coef, (1, 2, 0, 3)).reshape(n_blocks, nsub * 3 * nc)
        cheb[:, off:off + nsub * 3 * nc] = lay
    return dict(cheb=np.ascontiguousarray(cheb), jd_start=jd_start, block_days=BLOCK_DAYS,
                ipt=IPT.copy(), emrat=EMRAT, mjd_start=mjd_start, mjd_end=mjd_start + n_blocks * BLOCK_DAYS)


def earth_position_np(table, mjd_tt):
    """Vectorised numpy evaluation of the table (for building inputs only): (3, n) AU equatorial."""
    mjd_tt = np.asarray(mjd_tt, dtype=np.float64)
    jd = 2400000.5 + mjd_tt
    nr = np.floor((jd - table["jd_start"]) / table["block_days"]).astype(np.int64)
    tau = (jd - (table["jd_start"] + nr * table["block_days"])) / table["block_days"]
    out = []
    for b in range(3):
        off, nc, nsub = (int(x) for x in table["ipt"][b])
        sub = np.minimum(np.floor(tau * nsub), nsub - 1).astype(np.int64)
        tc = 2.0 * (tau * nsub - sub) - 1.0
        T = np.zeros((nc,) + tc.shape)
        T[0] = 1.0
        T[1] = tc
        for i in range(2, nc):
            T[i] = 2.0 * tc * T[i - 1] - T[i - 2]
        base = off + sub * 3 * nc
        pos = np.zeros((3,) + tc.shape)
        for ax in range(3):
            idx = base[None, :] + ax * nc + np.arange(nc)[:, None]
            c = table["cheb"][nr[None, :], idx]
            pos[ax] = (c * T).sum(axis=0)
        out.append(pos)
    emb, moon, sun = out
    return ((emb - moon / (1.0 + table["emrat"])) - sun) / AU_KM


# ------------------------------------------------------------------------------------------------
# trajectories
# ------------------------------------------------------------------------------------------------
def _elements_to_state_ecl(a, e, inc, node, argp, M):
    E = _kepler_E(M, e, iters=30)
    n = np.sqrt(MU / a ** 3)
    xp, yp = a * (np.cos(E) - e), a * np.sqrt(1 - e * e) * np.sin(E)
    r = a * (1 - e * np.cos(E))
    vxp, vyp = -a * a * n * np.sin(E) / r, a * a * n * np.sqrt(1 - e * e) * np.cos(E) / r
    cO, sO, ci, si, cw, sw = np.cos(node), np.sin(node), np.cos(inc), np.sin(inc), np.cos(argp), np.sin(argp)
    P = np.stack([cO * cw - sO * sw * ci, sO * cw + cO * sw * ci, sw * si])
    Q = np.stack([-cO * sw - sO * cw * ci, -sO * sw + cO * cw * ci, cw * si])
    return P * xp + Q * yp, P * vxp + Q * vyp


SITES = np.array([  # (longitude rad, rho cos phi, rho sin phi): geocentre + 4 observatories
    [0.0, 0.0, 0.0],
    [np.radians(204.5278), 0.94171, 0.33725],    # Haleakala-like
    [np.radians(289.2634), 0.86310, -0.50331],   # Cerro-Pachon-like
    [np.radians(243.1396), 0.83640, 0.54640],    # Palomar-like
    [np.radians(17.8800), 0.74730, 0.66230],     # mid-latitude Europe
])
ERAU = (6378137.0 / 1000.0) / AU_KM


def _site_geo_ecl(site_idx, mjd_tt):
    lon, rc, rs = SITES[site_idx].T
    theta = 4.894961212789145 + 2 * np.pi * 1.00273790934 * (mjd_tt - 51544.5) + lon
    equ = np.stack([ERAU * rc * np.cos(theta), ERAU * rc * np.sin(theta), ERAU * rs])
    return equ_to_ecl(equ)


def make_trajectories(n_traj, n_obs=12, seed=20261018, table=None, max_triplets=30, n_noise=10,
                      with_noise=True):
    """Synthetic batch.  n_obs: int (fixed) or (lo, hi) inclusive range per trajectory.

    Returns dict with traj_offset (uint64 [T+1]), mjd_tt, ra, dec, sigma_ra, sigma_dec ([n]),
    helio_equ, geo_ecl ([3, n] plane-major), body_fixed ([3, n]), mjd_ut1 ([n]),
    noise_z ([T, max_triplets, n_noise, 6]) or None, truth ([T, 6]).
    """
    if table is None:
        table = make_ephemeris_table()
    rng = np.random.default_rng(seed)
    T = int(n_traj)
    if isinstance(n_obs, (tuple, list)):
        counts = rng.integers(n_obs[0], n_obs[1] + 1, size=T)
    else:
        counts = np.full(T, int(n_obs))
    offs = np.zeros(T + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(counts)
    n = int(offs[-1])
    traj_of = np.repeat(np.arange(T), counts)
    first = np.asarray(offs[:-1], dtype=np.int64)
    k_in_traj = np.arange(n) - first[traj_of]

    # truth orbits (main belt / NEO like)
    a = rng.uniform(1.2, 3.5, T)
    e = rng.uniform(0.0, 0.4, T)
    inc = np.abs(rng.rayleigh(np.radians(8.0), T))
    node, argp, M0 = (rng.uniform(0, 2 * np.pi, T) for _ in range(3))
    epoch0 = rng.uniform(table["mjd_start"] + 200.0, table["mjd_end"] - 400.0, T)

    # observation times: nights of 3 exposures over a 20-60 d arc, irregular offsets (no exact ties)
    arc = rng.uniform(20.0, 60.0, T)
    n_nights = np.maximum(2, (counts + 2) // 3)
    night = k_in_traj // 3
    expo = k_in_traj % 3
    night_frac = night / np.maximum(1, n_nights[traj_of] - 1)
    night_jit = rng.uniform(-0.35, 0.35, (T, 16))[traj_of, np.minimum(night, 15)]
    t = epoch0[traj_of] + night_frac * arc[traj_of] + night_jit + expo * 0.021 + rng.uniform(0.0, 0.009, n)
    # sort within each trajectory (the boundary takes time-sorted observations)
    order = np.lexsort((t, traj_of))
    t = t[order]

    site_traj = rng.integers(0, len(SITES), T)
    site = site_traj[traj_of]
    geo_ecl = _site_geo_ecl(site, t)
    lon, rc, rs = SITES[site].T
    body_fixed = np.stack([ERAU * rc * np.cos(lon), ERAU * rc * np.sin(lon), ERAU * rs])
    earth = earth_position_np(table, t)
    helio_equ = earth + ecl_to_equ(geo_ecl)

    # forward model
    n_mot = np.sqrt(MU / a ** 3)
    Mt = M0[traj_of] + n_mot[traj_of] * (t - epoch0[traj_of])
    pos_ecl, vel_ecl = _elements_to_state_ecl(a[traj_of], e[traj_of], inc[traj_of], node[traj_of],
                                              argp[traj_of], Mt)
    rel = ecl_to_equ(pos_ecl) - helio_equ
    dist = np.sqrt((rel ** 2).sum(axis=0))
    cor = rel - dist / VLIGHT_AU * ecl_to_equ(vel_ecl)
    ra = np.arctan2(cor[1], cor[0]) % (2 * np.pi)
    dec = np.arctan2(cor[2], np.hypot(cor[0], cor[1]))
    sig = rng.uniform(0.1, 0.5, T)[traj_of] * ARCSEC
    dec = dec + rng.normal(0.0, 1.0, n) * sig
    ra = (ra + rng.normal(0.0, 1.0, n) * sig / np.maximum(np.cos(dec), 0.05)) % (2 * np.pi)

    noise = None
    if with_noise and n_noise > 0:
        noise = np.ascontiguousarray(rng.standard_normal((T, max_triplets, n_noise, 6)))
    c = np.ascontiguousarray
    return dict(traj_offset=offs, mjd_tt=c(t), ra=c(ra), dec=c(dec), sigma_ra=c(sig.copy()),
                sigma_dec=c(sig.copy()), helio_equ=c(helio_equ), geo_ecl=c(geo_ecl),
                body_fixed=c(body_fixed), mjd_ut1=c(t - 69.184 / 86400.0), noise_z=noise,
                truth=np.stack([a, e, inc, node, argp, M0, epoch0], axis=1), table=table,
                max_triplets=max_triplets, n_noise=n_noise)


# ------------------------------------------------------------------------------------------------
# bulk Kepler propagation inputs (reference kepler/propagation.rs:939-1000 strategies)
# ------------------------------------------------------------------------------------------------
def make_propagation_states(n, seed=20261018):
    """Random heliocentric states (elliptic and hyperbolic) and times of flight.

    Returns rv (6, n) plane-major [rx, ry, rz, vx, vy, vz], t0 (n), t1 (n).
    """
    rng = np.random.default_rng(seed)
    out_rv = np.empty((6, n))
    filled = 0
    while filled < n:
        m = int((n - filled) * 1.3) + 16
        q = rng.uniform(0.05, 150.0, m)
        e = rng.uniform(0.0, 5.0, m)
        frac = rng.uniform(-1.0, 1.0, m)
        inc = rng.uniform(0.0, np.pi, m)
        node, argp = rng.uniform(0, 2 * np.pi, m), rng.uniform(0, 2 * np.pi, m)
        nu_max = np.where(e < 1.0, np.pi, np.arccos(np.clip(-1.0 / np.maximum(e, 1.0 + 1e-9), -1, 1)) * 0.9)
        nu = frac * nu_max
        p = q * (1.0 + e)
        r = p / (1.0 + e * np.cos(nu))
        keep = (r >= 0.02) & (r <= 500.0) & (np.abs(e - 1.0) > 1e-3)
        h = np.sqrt(MU * p)
        xp, yp = r * np.cos(nu), r * np.sin(nu)
        vxp, vyp = -MU / h * np.sin(nu), MU / h * (e + np.cos(nu))
        cO, sO, ci, si, cw, sw = np.cos(node), np.sin(node), np.cos(inc), np.sin(inc), np.cos(argp), np.sin(argp)
        P = np.stack([cO * cw - sO * sw * ci, sO * cw + cO * sw * ci, sw * si])
        Q = np.stack([-cO * sw - sO * cw * ci, -sO * sw + cO * cw * ci, cw * si])
        rv = np.concatenate([P * xp + Q * yp, P * vxp + Q * vyp])[:, keep]
        take = min(rv.shape[1], n - filled)
        out_rv[:, filled:filled + take] = rv[:, :take]
        filled += take
    # time-of-flight mixture: short arcs, typical gaps, long gaps, either sign
    u = rng.uniform(0, 1, n)
    dt = np.where(u < 0.4, rng.uniform(0.01, 5.0, n), np.where(u < 0.8, rng.uniform(5.0, 60.0, n),
                                                               rng.uniform(60.0, 400.0, n)))
    dt *= np.where(rng.uniform(0, 1, n) < 0.3, -1.0, 1.0)
    t0 = np.full(n, 60000.0)
    return np.ascontiguousarray(out_rv), t0, t0 + dt


# ------------------------------------------------------------------------------------------------
# two-body ephemeris inputs (BASELINE configs[4]: orbits x daily epochs, one observer)
# ------------------------------------------------------------------------------------------------
def make_ephemeris_orbits(n, seed=20261018, epoch0=59000.0, mixed_kinds=False):
    """n elliptic main-belt / NEO-like orbits as (kind, epoch, elem[6, n]) in the OutfitIodResult
    convention: kind 0 Keplerian (a, e, i, Omega, omega, M).  With mixed_kinds a tenth of the orbits is
    given as Equinoctial (kind 1), a few as hyperbolic Cometary (kind 2: rejected, e >= 1) and one as
    parabolic Cometary (conversion error) to exercise the per-orbit error paths."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(1.2, 3.5, n)
    e = rng.uniform(0.0, 0.4, n)
    inc = np.abs(rng.rayleigh(np.radians(8.0), n)) % np.pi
    node, argp, M = (rng.uniform(0, 2 * np.pi, n) for _ in range(3))
    kind = np.zeros(n, dtype=np.int32)
    epoch = epoch0 + rng.uniform(-50.0, 50.0, n)
    elem = np.ascontiguousarray(np.stack([a, e, inc, node, argp, M]))
    if mixed_kinds and n >= 40:
        sel = np.arange(0, n, 10)
        dig = node[sel] + argp[sel]
        th = np.tan(inc[sel] / 2.0)
        elem[:, sel] = np.stack([a[sel], e[sel] * np.sin(dig), e[sel] * np.cos(dig), th * np.sin(node[sel]),
                                 th * np.cos(node[sel]), (dig + M[sel]) % (2 * np.pi)])
        kind[sel] = 1
        hyp = np.arange(5, n, 37)
        elem[0, hyp] = rng.uniform(0.5, 2.0, hyp.size)      # q
        elem[1, hyp] = rng.uniform(1.05, 2.5, hyp.size)     # e > 1
        elem[5, hyp] = rng.uniform(-1.0, 1.0, hyp.size)     # nu
        kind[hyp] = 2
        elem[1, 7] = 1.0                                     # parabolic cometary: InvalidConversion
        kind[7] = 2
    return kind, np.ascontiguousarray(epoch), elem


def make_ephemeris_epochs(n_epochs, mjd0=59000.25, step=1.0, site_idx=1):
    """Daily epochs (MJD TT), the matching UT1 argument (dUT1 = 0: UT1 = TT - 69.184 s here) and the
    Earth-fixed observer position (AU) of one of the synthetic sites."""
    mjd_tt = mjd0 + step * np.arange(n_epochs, dtype=np.float64)
    mjd_ut1 = mjd_tt - 69.184 / 86400.0
    lon, rc, rs = SITES[site_idx]
    body_fixed = np.array([ERAU * rc * np.cos(lon), ERAU * rc * np.sin(lon), ERAU * rs])
    return np.ascontiguousarray(mjd_tt), np.ascontiguousarray(mjd_ut1), body_fixed
