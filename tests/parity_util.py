"""Shared helpers of the parity tests (GPU path vs oracle)."""
import numpy as np

# north_star tolerances: orbital elements within 1e-10 relative, RMS within 1e-9 relative.
ELEM_TOL = 1e-10
RMS_TOL = 1e-9
# Gauss IOD is ill-conditioned for slow movers / short arcs: moving the INPUT RA/Dec by one ulp moves
# the oracle's own answer by up to ~1e-9 (elements) on a few per cent of the trajectories.  Where the
# oracle's 1-ulp sensitivity exceeds the plain tolerance, the bound is FLOOR_FACTOR x that sensitivity.
FLOOR_FACTOR = 256.0

INT_FIELDS = ("status", "cause", "attempts", "corrected", "element_kind", "triplet_idx", "triplet_rank",
              "realization")


def elem_err(a, b):
    """a, e (or q, e): relative; i, node, argp, anomaly: wrapped absolute angle difference (rad)."""
    out = np.empty(a.shape[0])
    d = np.abs(a - b)
    rel = d[:, 0:2] / np.maximum(np.abs(b[:, 0:2]), 1e-300)
    ang = np.abs((a[:, 2:6] - b[:, 2:6] + np.pi) % (2 * np.pi) - np.pi)
    out[:] = np.maximum(rel.max(axis=1), ang.max(axis=1))
    return out


def oracle_floor(O, synth, batch, et, op, base):
    """Per-trajectory sensitivity of the ORACLE to +-1 ulp on RA/Dec: (elem_floor, rms_floor)."""
    ef = np.zeros(len(base))
    rf = np.zeros(len(base))
    for sgn_ra, sgn_dec in ((np.inf, -np.inf), (-np.inf, np.inf), (np.inf, np.inf)):
        ob = O.from_soa_batch(batch)
        ob["ra"] = np.nextafter(ob["ra"], sgn_ra)
        ob["dec"] = np.nextafter(ob["dec"], sgn_dec)
        pert = O.fit_full_iod(ob, et, op, n_threads=0)
        same = (pert["status"] == 0) & (base["status"] == 0) & (pert["realization"] == base["realization"]) & \
               (pert["triplet_idx"] == base["triplet_idx"]).all(axis=1) & \
               (np.abs(pert["epoch"] - base["epoch"]) <= 1e-8)  # an f-g loop that commits or not (epoch 0.0 quirk)
        e = np.where(same, elem_err(pert["elem"], base["elem"]), np.inf)
        r = np.where(same, np.abs(pert["rms"] - base["rms"]) / np.maximum(np.abs(base["rms"]), 1e-300), np.inf)
        ef = np.maximum(ef, e)
        rf = np.maximum(rf, r)
    return ef, rf


def assert_iod_parity(got, want, elem_floor=None, rms_floor=None, min_plain_fraction=None, max_outlier_fraction=0.0):
    """Integer / index / status fields bit-exact; floats within the tolerance rule above."""
    # Integer / index fields: equal on every trajectory -- except, at scale, trajectories whose ORACLE
    # answer itself changes selection when RA/Dec move by one ulp (floor = inf: a candidate that exists or
    # not depending on the last bit, e.g. a Kepler solve on the edge of its Newton budget).  Those may
    # land on the oracle's other answer; they must stay below 5 per 10 000 and need the floors.
    mism = np.zeros(len(want), dtype=bool)
    for f in INT_FIELDS:
        d = got[f] != want[f]
        mism |= d if d.ndim == 1 else d.any(axis=1)
    if mism.any():
        assert elem_floor is not None, f"integer fields differ on {np.argwhere(mism)[:5].ravel()}"
        unstable = ~np.isfinite(elem_floor)
        assert (mism & ~unstable).sum() == 0, f"integer fields differ on stable trajectories {np.argwhere(mism & ~unstable)[:5].ravel()}"
        assert mism.sum() <= int(5e-4 * len(want)), f"{mism.sum()} selection flips in {len(want)} trajectories"
        got, want = got[~mism], want[~mism]
        elem_floor, rms_floor = elem_floor[~mism], rms_floor[~mism]
    bad = want["status"] != 0
    # payloads of the error variants
    assert np.array_equal(got["span"][bad], want["span"][bad])
    cv_g, cv_w = got["cause_value"][bad], want["cause_value"][bad]
    assert np.array_equal(np.isnan(cv_g), np.isnan(cv_w)) and np.array_equal(cv_g[~np.isnan(cv_g)], cv_w[~np.isnan(cv_w)])
    ok = ~bad
    if not ok.any():
        return dict(n_ok=0)
    ee = elem_err(got["elem"][ok], want["elem"][ok])
    er = np.abs(got["rms"][ok] - want["rms"][ok]) / np.maximum(np.abs(want["rms"][ok]), 1e-300)
    # absolute, in days (light-time corrected epoch; it is 0.0 when no f-g iteration ever committed,
    # gauss.rs:1299,1417 -- reference behaviour)
    ep = np.abs(got["epoch"][ok] - want["epoch"][ok])
    etol = np.full(ok.sum(), ELEM_TOL)
    rtol = np.full(ok.sum(), RMS_TOL)
    if elem_floor is not None:
        etol = np.maximum(etol, FLOOR_FACTOR * elem_floor[ok])
        rtol = np.maximum(rtol, FLOOR_FACTOR * rms_floor[ok])
    # Near-parabolic candidates (e > 0.99) send the reference's elliptic initial guess
    # (prelim_elliptic.rs:113-134: Newton on Kepler's equation from u = M) to |psi| ~ 1e6, after which
    # the universal-variable Newton (50 steps) converges or not depending on last-bit details: a
    # chaotic branch of the REFERENCE algorithm.  Such trajectories may differ in the floats (never in
    # the integer fields) but must stay below 0.5 % of the batch.
    near_parab = (want["elem"][ok][:, 1] > 0.99) | (got["elem"][ok][:, 1] > 0.99)
    exempt = near_parab & ((ee > etol) | (er > rtol))
    assert exempt.sum() <= max(1, int(0.005 * ok.sum())), exempt.sum()
    etol = np.where(exempt, np.inf, etol)
    rtol = np.where(exempt, np.inf, rtol)
    # (the floors come from three 1-ulp probes per trajectory: on 10^4+ trajectories a few have a
    # sensitivity the probes missed -- `max_outlier_fraction` bounds them instead of forbidding them)
    n_out = int(((ee > etol) | (er > rtol)).sum())
    assert n_out <= int(max_outlier_fraction * ok.sum()), \
        f"{n_out} float outliers: element error {ee.max():.3e} on {np.argwhere(ee > etol)[:5].ravel()}, rms error {er.max():.3e}"
    # trajectories whose ORACLE answer is itself discontinuous under a 1-ulp move of the inputs
    # (floor = inf: a different branch, e.g. an f-g loop that never commits) only have to agree on
    # the integer / index fields; they must stay rare
    chaotic = np.zeros(ok.sum(), dtype=bool) if elem_floor is None else ~np.isfinite(elem_floor[ok])
    assert chaotic.mean() <= 0.02, chaotic.mean()
    chaotic = chaotic | exempt
    assert (ep[~chaotic] > 1e-8).sum() <= int(max_outlier_fraction * ok.sum()), f"epoch error {ep[~chaotic].max():.3e} d"
    if min_plain_fraction is not None:
        assert (ee <= ELEM_TOL).mean() >= min_plain_fraction, (ee <= ELEM_TOL).mean()
        assert (er <= RMS_TOL).mean() >= min_plain_fraction, (er <= RMS_TOL).mean()
    return dict(n_ok=int(ok.sum()), elem_p50=float(np.median(ee)), elem_max=float(ee.max()),
                rms_p50=float(np.median(er)), rms_max=float(er.max()),
                plain_elem_fraction=float((ee <= ELEM_TOL).mean()), plain_rms_fraction=float((er <= RMS_TOL).mean()))


def oracle_observer_cache(O, et, batch):
    """OutfitCache::build with the ORACLE (pvobs + helio_position) for a body-fixed batch:
    returns (helio_equ [n,3], geo_ecl [n,3]) as the oracle's fit_full_iod expects them."""
    import ctypes as C
    n = len(batch["ra"])
    geo = np.zeros((n, 3))
    hel = np.zeros((n, 3))
    L = O.lib()
    for i in range(n):
        dx, dv, vb, h = O.D3(), O.D3(), O.D3(0, 0, 0), O.D3()
        L.oo_pvobs(float(batch["mjd_tt"][i]), float(batch["mjd_ut1"][i]), O.d3(batch["body_fixed"][:, i]), vb, dx, dv)
        geo[i] = list(dx)
        assert L.oo_helio_position(C.byref(et), float(batch["mjd_tt"][i]), dx, h) == 0
        hel[i] = list(h)
    return np.ascontiguousarray(hel), np.ascontiguousarray(geo)


# ---- differential orbit correction (FitLSQ) -------------------------------------------------------
LSQ_INT_FIELDS = ("status", "kind", "fallback_cause", "total_newton_iterations", "num_measurements")
LSQ_ELEM_TOL = 1e-10      # a relative; h, k, p, q absolute; lambda wrapped (rad)
LSQ_RMS_TOL = 1e-9        # relative
LSQ_COV_TOL = 1e-6        # relative Frobenius error of the 6x6 covariance (an inverse: cond(N) * eps)
LSQ_RES_TOL = 1e-7        # per-observation residuals, in units of the observation's sigma


def lsq_elem_err(a, b):
    d = np.abs(a - b)
    d[:, 0] /= np.maximum(np.abs(b[:, 0]), 1e-300)
    d[:, 5] = np.abs((a[:, 5] - b[:, 5] + np.pi) % (2 * np.pi) - np.pi)
    return d.max(axis=1)


def _lsq_errs(got, want, gfit, wfit, off, sig):
    ee = lsq_elem_err(got["elem"], want["elem"])
    er = np.abs(got["normalised_rms"] - want["normalised_rms"]) / np.maximum(np.abs(want["normalised_rms"]), 1e-300)
    cn = np.linalg.norm(want["covariance"], axis=1)
    ec = np.linalg.norm(got["covariance"] - want["covariance"], axis=1) / np.maximum(cn, 1e-300)
    ro = np.maximum(np.abs(gfit["residual_ra"] - wfit["residual_ra"]) / sig[0],
                    np.abs(gfit["residual_dec"] - wfit["residual_dec"]) / sig[1])
    eo = np.maximum.reduceat(np.concatenate([ro, [0.0]]), off[:-1].astype(np.int64))
    eo = np.where(off[1:] > off[:-1], eo, 0.0)
    return ee, er, ec, eo


def lsq_int_mismatch(got, want, gfit, wfit, off, diagnostics=False):
    """Trajectories whose OUTCOME differs: status, kind and -- for corrected orbits -- the iteration count,
    the number of measurements and the per-observation selection flags.  When the loop fails the reference
    returns the IOD orbit and drops the error (mod.rs:113): `fallback_cause` / `total_newton_iterations` of
    a fallback are diagnostics this implementation adds, compared separately (diagnostics=True)."""
    both_fb = (got["kind"] == 2) & (want["kind"] == 2)
    if diagnostics:
        return both_fb & ((got["fallback_cause"] != want["fallback_cause"]) |
                          (got["total_newton_iterations"] != want["total_newton_iterations"]))
    mism = (got["status"] != want["status"]) | (got["kind"] != want["kind"])
    for f in ("total_newton_iterations", "num_measurements", "fallback_cause"):
        mism |= (got[f] != want[f]) & ~both_fb
    sel = (gfit["selection"] != wfit["selection"]).astype(np.int64)
    per = np.add.reduceat(np.concatenate([sel, [0]]), off[:-1].astype(np.int64))
    return mism | ((off[1:] > off[:-1]) & (per > 0))


def oracle_lsq_floor(O, ob, et, cfg, iod, base, bfit):
    """Sensitivity of the ORACLE's LSQ answer to +-1 ulp on RA/Dec (same initial orbits): per-trajectory
    floors (elements, rms, covariance, residuals) and the trajectories whose integer outcome itself flips."""
    off = ob["traj_offset"]
    sig = (ob["sigma_ra"], ob["sigma_dec"])
    fl = [np.zeros(len(base)) for _ in range(4)]
    unstable = np.zeros(len(base), dtype=bool)
    for s_ra, s_dec in ((np.inf, -np.inf), (-np.inf, np.inf), (np.inf, np.inf)):
        pb = dict(ob)
        pb["ra"] = np.nextafter(ob["ra"], s_ra)
        pb["dec"] = np.nextafter(ob["dec"], s_dec)
        pert, pfit = O.fit_lsq(pb, et, cfg, iod, n_threads=0)
        unstable |= lsq_int_mismatch(pert, base, pfit, bfit, off)
        for i, e in enumerate(_lsq_errs(pert, base, pfit, bfit, off, sig)):
            fl[i] = np.maximum(fl[i], e)
    return fl, unstable


def assert_lsq_parity(got, want, gfit, wfit, ob, floors=None, unstable=None, max_flip_fraction=2e-3,
                      max_outlier_fraction=2e-3, sigma_tol=None):
    """Outcome / counters / selection flags equal; floats within the tolerances above or FLOOR_FACTOR x the
    oracle's own 1-ulp sensitivity.  Trajectories whose ORACLE outcome flips under a 1-ulp move of the
    inputs (a convergence test on the edge) may land on the other outcome; they must stay rare."""
    off = ob["traj_offset"]
    sig = (ob["sigma_ra"], ob["sigma_dec"])
    mism = lsq_int_mismatch(got, want, gfit, wfit, off)
    if mism.any():
        # (three probes sample the +-1 ulp neighbourhood: a fourth direction flips a few trajectories the
        # probes left "stable" -- measured on the oracle itself, ~1e-3 of a batch -- so the flips are
        # bounded in number, and `unstable` must show the oracle has such flips at all)
        assert unstable is not None, f"LSQ outcome differs on {np.argwhere(mism)[:5].ravel()}"
        assert mism.sum() <= max(3, int(max_flip_fraction * len(want))), \
            f"{mism.sum()} outcome flips in {len(want)} trajectories ({unstable.sum()} oracle-unstable), e.g. {np.argwhere(mism)[:5].ravel()}"
    # failed loops whose failure differs (cause or iteration count): the orbit returned is the same IOD orbit;
    # these are the starts with absurd residuals (IOD rms ~1e5) whose Newton steps wander chaotically
    diag = lsq_int_mismatch(got, want, gfit, wfit, off, diagnostics=True)
    assert diag.sum() <= max(3, int(0.03 * (want["kind"] == 2).sum())), (diag.sum(), (want["kind"] == 2).sum())
    ok = ~mism & (want["kind"] == 1)
    fb = ~mism & (want["kind"] == 2)
    # fallbacks return the IOD orbit untouched
    assert np.array_equal(got["elem"][fb], want["elem"][fb]) and np.array_equal(got["epoch"][fb], want["epoch"][fb])
    assert np.array_equal(got["normalised_rms"][fb], want["normalised_rms"][fb])
    assert np.array_equal(got["epoch"][ok], want["epoch"][ok])
    ee, er, ec, eo = _lsq_errs(got, want, gfit, wfit, off, sig)
    tol = [np.full(len(want), t) for t in (LSQ_ELEM_TOL, LSQ_RMS_TOL, LSQ_COV_TOL, LSQ_RES_TOL)]
    if floors is not None:
        tol = [np.maximum(t, FLOOR_FACTOR * f) for t, f in zip(tol, floors)]
    # the covariance is the inverse of a normal matrix whose entries carry ~1 ulp of libm noise:
    # relative error up to cond(N) * eps, whatever the implementation
    nm = np.where(np.isfinite(want["normal_matrix"]), want["normal_matrix"], 0.0).reshape(-1, 6, 6)
    with np.errstate(all="ignore"):
        cond = np.where(ok, np.linalg.cond(np.where(ok[:, None, None], nm, np.eye(6))), 1.0)
    tol[2] = np.maximum(tol[2], 64 * 2.220446049250313e-16 * np.where(np.isfinite(cond), cond, 1e300))
    if sigma_tol is not None:
        # a loop stopped BEFORE convergence (iteration cap) returns x0 + dx, and dx = Gamma G^T W xi carries
        # the same cond(N) * eps of the inverse; a converged loop does not (its last dx is ~0).  The 1-ulp
        # probes on RA/Dec move xi, not G, so they do not see it: such runs are compared in units of the
        # fit's own 1-sigma uncertainty instead.
        d = np.abs(got["elem"] - want["elem"])
        d[:, 5] = np.abs((got["elem"][:, 5] - want["elem"][:, 5] + np.pi) % (2 * np.pi) - np.pi)
        with np.errstate(all="ignore"):
            within = np.all(d <= sigma_tol * np.where(want["sigma"] > 0, want["sigma"], np.inf), axis=1)
        ee = np.where(within, 0.0, ee)
        tol[3] = np.maximum(tol[3], np.where(within, np.inf, 0.0))
    bad = ok & ((ee > tol[0]) | (er > tol[1]) | (ec > tol[2]) | (eo > tol[3]))
    if unstable is not None:
        bad &= ~unstable
    assert bad.sum() <= int(max_outlier_fraction * max(ok.sum(), 1)), \
        f"{bad.sum()} float outliers of {ok.sum()}: elem {ee[ok].max():.3e} rms {er[ok].max():.3e} cov {ec[ok].max():.3e} res {eo[ok].max():.3e} at {np.argwhere(bad)[:5].ravel()}"
    # sigma = sqrt(diag(covariance)) on the device too
    d = got["covariance"][ok][:, ::7]
    assert np.allclose(got["sigma"][ok], np.sqrt(d), rtol=1e-15, atol=0, equal_nan=True)
    return dict(n_corrected=int(ok.sum()), n_fallback=int(fb.sum()), n_flips=int(mism.sum()), n_diag=int(diag.sum()),
                elem_p50=float(np.median(ee[ok])) if ok.any() else 0.0, elem_max=float(ee[ok].max()) if ok.any() else 0.0,
                plain_fraction=float(((ee <= LSQ_ELEM_TOL) & (er <= LSQ_RMS_TOL))[ok].mean()) if ok.any() else 1.0,
                cov_max=float(ec[ok].max()) if ok.any() else 0.0)
