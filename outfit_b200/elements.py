"""Host-side element conversions of the result records (vectorised numpy over the batch).

`FitLSQ` returns equinoctial elements with their covariance (`OrbitalElements::Equinoctial{elements,
uncertainty, covariance}`, differential_orbit_correction/diff_cor.rs:232-244); callers of the reference turn
them into Keplerian elements with `OrbitalElements::to_keplerian`, which propagates the covariance through
the analytic Jacobian.  This module mirrors that step for whole batches of `LSQ_RESULT_DTYPE` records:

  equinoctial_to_keplerian   KeplerianElements::from_equinoctial_internal   orbit_type/keplerian_element.rs:185-233
  jacobian_to_keplerian      EquinoctialElements::jacobian_to_keplerian     orbit_type/equinoctial_element.rs:1049-1140
  propagate_covariance       OrbitalCovariance::propagate (J C J^T)         orbit_type/uncertainty.rs:412-416
  keplerian_sigmas           KeplerianUncertainty::from_covariance          orbit_type/uncertainty.rs:244-254
"""
import numpy as np

TWO_PI = 6.283185307179586476925286766559
_EPS = 1.0e-12


def equinoctial_to_keplerian(eq):
    """(n, 6) equinoctial (a, h, k, p, q, lambda) -> (n, 6) Keplerian (a, e, i, Omega, omega, M)."""
    eq = np.asarray(eq, dtype=np.float64).reshape(-1, 6)
    a, h, k, p, q, lam = (eq[:, j] for j in range(6))
    ecc = np.sqrt(h * h + k * k)
    dig = np.where(ecc < _EPS, 0.0, np.arctan2(h, k))
    t = np.sqrt(p * p + q * q)
    node = np.where(t < _EPS, 0.0, np.arctan2(p, q))
    inc = 2.0 * np.arctan(t)
    return np.stack([a, ecc, inc, node, np.mod(dig - node, TWO_PI), np.mod(lam - dig, TWO_PI)], axis=1)


def jacobian_to_keplerian(eq):
    """(n, 6) -> (n, 6, 6): J[t, row, col] = d(a, e, i, Omega, omega, M)[row] / d(a, h, k, p, q, lambda)[col]."""
    eq = np.asarray(eq, dtype=np.float64).reshape(-1, 6)
    h, k, p, q = eq[:, 1], eq[:, 2], eq[:, 3], eq[:, 4]
    e = np.sqrt(h * h + k * k)
    t = np.sqrt(p * p + q * q)
    with np.errstate(divide="ignore", invalid="ignore"):
        e_ok, t_ok = ~(e < _EPS), ~(t < _EPS)
        dvh = np.where(e_ok, k / (e * e), 0.0)
        dvk = np.where(e_ok, -h / (e * e), 0.0)
        denom = t * (1.0 + t * t)
        dip = np.where(t_ok, 2.0 * p / denom, 0.0)
        diq = np.where(t_ok, 2.0 * q / denom, 0.0)
        dnp = np.where(t_ok, q / (t * t), 0.0)
        dnq = np.where(t_ok, -p / (t * t), 0.0)
    em = np.maximum(e, _EPS)
    J = np.zeros((eq.shape[0], 6, 6))
    J[:, 0, 0] = 1.0
    J[:, 1, 1], J[:, 4, 1], J[:, 5, 1] = h / em, dvh, -dvh
    J[:, 1, 2], J[:, 4, 2], J[:, 5, 2] = k / em, dvk, -dvk
    J[:, 2, 3], J[:, 3, 3], J[:, 4, 3] = dip, dnp, -dnp
    J[:, 2, 4], J[:, 3, 4], J[:, 4, 4] = diq, dnq, -dnq
    J[:, 5, 5] = 1.0
    return J


def propagate_covariance(cov, jac):
    """J C J^T per trajectory; cov (n, 6, 6) row/col indexed (a record's column-major `covariance` field
    reshaped (6, 6) and transposed -- or not: it is symmetric up to rounding)."""
    return jac @ np.asarray(cov, dtype=np.float64).reshape(-1, 6, 6) @ np.transpose(jac, (0, 2, 1))


def keplerian_sigmas(cov):
    with np.errstate(invalid="ignore"):
        return np.sqrt(np.einsum("nii->ni", np.asarray(cov).reshape(-1, 6, 6)))


def lsq_to_keplerian(results):
    """LSQ_RESULT_DTYPE records -> dict(elem (n, 6), covariance (n, 6, 6), sigma (n, 6), valid (n,)): the
    Keplerian form of the corrected orbits (kind == 1); other records are NaN with valid = False."""
    n = len(results)
    ok = results["kind"] == 1
    elem = np.full((n, 6), np.nan)
    cov = np.full((n, 6, 6), np.nan)
    if ok.any():
        eq = results["elem"][ok]
        elem[ok] = equinoctial_to_keplerian(eq)
        c = results["covariance"][ok].reshape(-1, 6, 6).transpose(0, 2, 1)  # column-major -> [row, col]
        cov[ok] = propagate_covariance(c, jacobian_to_keplerian(eq))
    return dict(elem=elem, covariance=cov, sigma=keplerian_sigmas(cov), valid=ok, epoch=results["epoch"].copy())
