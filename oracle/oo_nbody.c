/*
 * oo_nbody.c -- CPU ORACLE (test infrastructure only) for the N-body propagator of the reference
 * (paths relative to /root/reference/src):
 *
 *   EquinoctialElements::propagate_nbody      orbit_type/equinoctial_element.rs:908-968
 *   NBodyOde::diff and its helpers            propagator/nbody.rs:127-356 (direct / indirect acceleration of every
 *                                             perturber FROZEN at t0, gravity gradient, variational equations on
 *                                             the 42-dimensional augmented state [r, v, Phi column-major])
 *   integrate_augmented_state                 propagator/nbody.rs:505-523 (DOP853, abs_tol / rel_tol of NBodyConfig)
 *   NBodyConfig::default                      propagator/mod.rs:139-150
 *   planet GMs                                propagator/planet_gm.rs:10-60
 *
 * The integrator of the reference is `ExplicitRungeKutta::dop853()` of the un-vendored crate
 * `differential_equations`; its step-size controller is not available here, so the DOP853 below restates the
 * PUBLISHED method (Hairer, Norsett & Wanner; coefficients and controller as implemented by scipy.integrate.DOP853:
 * initial step of Sec. II.4, safety 0.9, step factors in [0.2, 10], combined 5th / 3rd-order error estimate) and is
 * pinned against scipy itself (tests/test_nbody.py).  PARITY WITH THE CRATE IS UNPINNED: the accepted steps may differ,
 * the propagated state agrees at the level of the tolerances (1e-12).
 * The reference's indirect term carries a + sign (nbody.rs:159-171); kept as is.
 */
#include <math.h>
#include <string.h>

#include "dop853_coeffs.h"
#include "oo.h"

#define NB_DIM 42

/* NBodyOde::diff (nbody.rs:339-356) */
void oo_nbody_rhs(const double *y, const oo_perturber *pert, size_t n_pert, double *dy) {
  double acc[3] = {0.0, 0.0, 0.0}, G[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; /* G row-major: G[3*row + col] */
  for (size_t p = 0; p < n_pert; p++) {
    const double gm = pert[p].gm;
    double d[3] = {y[0] - pert[p].pos[0], y[1] - pert[p].pos[1], y[2] - pert[p].pos[2]};
    /* direct_acceleration (:127-134): -GM / |d|^3 * d, |d|^3 = powi(3) */
    double dist = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    double dist3 = dist * dist * dist;
    double cdir = -gm / dist3;
    double adir[3] = {cdir * d[0], cdir * d[1], cdir * d[2]};
    /* indirect_acceleration (:159-171) */
    double aind[3] = {0.0, 0.0, 0.0};
    double pd = sqrt(pert[p].pos[0] * pert[p].pos[0] + pert[p].pos[1] * pert[p].pos[1] + pert[p].pos[2] * pert[p].pos[2]);
    if (pd > 1e-10) {
      double c = gm / (pd * pd * pd);
      for (int i = 0; i < 3; i++) aind[i] = c * pert[p].pos[i];
    }
    for (int i = 0; i < 3; i++) acc[i] = acc[i] + adir[i] + aind[i];
    /* gravity_gradient_contribution (:194-204): -GM (I / |d|^3 - 3 d d^T / |d|^5) */
    double dist5 = dist * dist * dist * dist * dist;
    double a = 1.0 / dist3, b = 3.0 / dist5;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) G[3 * r + c] = G[3 * r + c] + (-gm) * ((r == c ? 1.0 : 0.0) * a - (d[r] * d[c]) * b);
  }
  dy[0] = y[3]; dy[1] = y[4]; dy[2] = y[5];
  dy[3] = acc[0]; dy[4] = acc[1]; dy[5] = acc[2];
  /* dPhi/dt = A Phi, A = [[0, I], [G, 0]]; Phi column-major: Phi(r, c) = y[6 + 6 c + r] */
  for (int c = 0; c < 6; c++) {
    const double *ph = y + 6 + 6 * c;
    double *dp = dy + 6 + 6 * c;
    dp[0] = ph[3]; dp[1] = ph[4]; dp[2] = ph[5];
    for (int r = 0; r < 3; r++) dp[3 + r] = (G[3 * r + 0] * ph[0] + G[3 * r + 1] * ph[1]) + G[3 * r + 2] * ph[2];
  }
}

static double rms_norm(const double *x, const double *scale, int n) {
  double s = 0.0;
  for (int i = 0; i < n; i++) { double q = x[i] / scale[i]; s += q * q; }
  return sqrt(s) / sqrt((double)n);
}

/* DOP853 from t = 0 to t = span on the augmented state (scipy.integrate.DOP853 / Hairer's dop853.f controller).
 * Returns OO_OK or OO_ERR_NBODY (step size underflow / step budget exhausted / non-finite state). */
int oo_dop853_nbody(double *y, double span, const oo_perturber *pert, size_t n_pert, double atol, double rtol,
                    uint32_t max_steps, uint32_t *n_steps, uint32_t *n_rejected) {
  const int n = NB_DIM;
  const double direction = span >= 0.0 ? 1.0 : -1.0;
  const double interval = fabs(span);
  const double err_exp = -1.0 / 8.0; /* error estimator order 7 */
  double K[DOP853_STAGES + 1][NB_DIM], f[NB_DIM], ynew[NB_DIM], ytmp[NB_DIM], scale[NB_DIM], tmp[NB_DIM];
  uint32_t steps = 0, rejected = 0;
  if (n_steps) *n_steps = 0;
  if (n_rejected) *n_rejected = 0;
  if (interval == 0.0) return OO_OK;
  oo_nbody_rhs(y, pert, n_pert, f);
  /* select_initial_step (Hairer II.4) */
  double h_abs;
  {
    for (int i = 0; i < n; i++) scale[i] = atol + fabs(y[i]) * rtol;
    double d0 = rms_norm(y, scale, n), d1 = rms_norm(f, scale, n);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    if (h0 > interval) h0 = interval;
    for (int i = 0; i < n; i++) ytmp[i] = y[i] + h0 * direction * f[i];
    oo_nbody_rhs(ytmp, pert, n_pert, tmp);
    for (int i = 0; i < n; i++) tmp[i] = tmp[i] - f[i];
    double d2 = rms_norm(tmp, scale, n) / h0;
    double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), 1.0 / 8.0);
    h_abs = fmin(fmin(100.0 * h0, h1), interval);
  }
  double t = 0.0;
  while (direction * (t - span) < 0.0) {
    if (steps >= max_steps) return OO_ERR_NBODY;
    const double min_step = 10.0 * fabs(nextafter(t, direction * INFINITY) - t);
    if (h_abs < min_step) h_abs = min_step;
    int accepted = 0, was_rejected = 0;
    double t_new = t, h = 0.0;
    while (!accepted) {
      if (h_abs < min_step) return OO_ERR_NBODY;
      h = h_abs * direction;
      t_new = t + h;
      if (direction * (t_new - span) > 0.0) t_new = span;
      h = t_new - t;
      h_abs = fabs(h);
      /* rk_step */
      memcpy(K[0], f, sizeof f);
      for (int s = 1; s < DOP853_STAGES; s++) {
        const double *a = DOP853_A + (s * (s - 1)) / 2;
        for (int i = 0; i < n; i++) {
          double acc = 0.0;
          for (int j = 0; j < s; j++) acc += K[j][i] * a[j];
          ytmp[i] = y[i] + acc * h;
        }
        oo_nbody_rhs(ytmp, pert, n_pert, K[s]);
      }
      for (int i = 0; i < n; i++) {
        double acc = 0.0;
        for (int j = 0; j < DOP853_STAGES; j++) acc += K[j][i] * DOP853_B[j];
        ynew[i] = y[i] + h * acc;
      }
      oo_nbody_rhs(ynew, pert, n_pert, K[DOP853_STAGES]);
      /* _estimate_error_norm */
      double e5 = 0.0, e3 = 0.0;
      for (int i = 0; i < n; i++) {
        const double sc = atol + fmax(fabs(y[i]), fabs(ynew[i])) * rtol;
        double a5 = 0.0, a3 = 0.0;
        for (int j = 0; j <= DOP853_STAGES; j++) { a5 += K[j][i] * DOP853_E5[j]; a3 += K[j][i] * DOP853_E3[j]; }
        a5 /= sc; a3 /= sc;
        e5 += a5 * a5; e3 += a3 * a3;
      }
      double err;
      if (e5 == 0.0 && e3 == 0.0) err = 0.0;
      else err = fabs(h) * e5 / sqrt((e5 + 0.01 * e3) * (double)n);
      if (!(err == err)) return OO_ERR_NBODY; /* NaN: the state left the domain */
      if (err < 1.0) {
        double factor = err == 0.0 ? 10.0 : fmin(10.0, 0.9 * pow(err, err_exp));
        if (was_rejected) factor = fmin(1.0, factor);
        h_abs *= factor;
        accepted = 1;
      } else {
        h_abs *= fmax(0.2, 0.9 * pow(err, err_exp));
        was_rejected = 1;
        rejected++;
      }
    }
    t = t_new;
    memcpy(y, ynew, sizeof ynew);
    memcpy(f, K[DOP853_STAGES], sizeof f);
    steps++;
  }
  if (n_steps) *n_steps = steps;
  if (n_rejected) *n_rejected = rejected;
  return OO_OK;
}

/* EquinoctialElements::propagate_nbody (equinoctial_element.rs:908-968) without the element Jacobians: heliocentric
 * position / velocity (ecliptic J2000) at t1 and the state transition matrix Phi(t1, t0), column-major. */
int oo_propagate_nbody(const oo_elements *equi, double t1_mjd_tt, const oo_perturber *pert, size_t n_pert, double atol,
                       double rtol, double pos[3], double vel[3], double stm[36], uint32_t *n_steps) {
  const double span = t1_mjd_tt - equi->epoch;
  double p0[3], v0[3];
  int rc = oo_propagate_twobody(equi, 0.0, 0.0, p0, v0);
  if (rc != OO_OK) return rc;
  double y[NB_DIM];
  memset(y, 0, sizeof y);
  for (int i = 0; i < 3; i++) { y[i] = p0[i]; y[3 + i] = v0[i]; }
  for (int i = 0; i < 6; i++) y[6 + 7 * i] = 1.0;
  if (n_steps) *n_steps = 0;
  if (fabs(span) >= 1e-14) {
    rc = oo_dop853_nbody(y, span, pert, n_pert, atol, rtol, 100000u, n_steps, NULL);
    if (rc != OO_OK) return rc;
  }
  for (int i = 0; i < 3; i++) { pos[i] = y[i]; vel[i] = y[3 + i]; }
  if (stm) memcpy(stm, y + 6, 36 * sizeof(double));
  return OO_OK;
}

/* planet_gm.rs:10-60: GM in AU^3 / day^2; index = 0 Sun, 1 Mercury, 2 Venus, 3 Earth-Moon, 4 Mars, 5 Jupiter, 6 Saturn,
 * 7 Uranus, 8 Neptune, 9 Pluto, 10 Moon */
double oo_planet_gm(int body) {
  static const double km3_s2[11] = {1.32712440041e11, 2.203178e4, 3.2485857e5, 4.03503235e5, 4.28283736e4, 1.267127648e8,
                                    3.79406252e7, 5.7945564e6, 6.8365271e6, 9.755e2, 4.902800066e3};
  const double au_km = 1.495978707e8;
  const double conv = (86400.0 * 86400.0) / (au_km * au_km * au_km);
  return (body >= 0 && body < 11) ? km3_s2[body] * conv : NAN;
}
