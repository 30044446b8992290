/* oo_lsq.c -- CPU ORACLE (test infrastructure only): differential orbit correction, the weighted
 * least-squares Newton-Raphson refinement of an IOD orbit.
 *
 * Restates, operation for operation:
 *   differential_orbit_correction/mod.rs:60-115          differential_correction
 *   differential_orbit_correction/diff_cor.rs:282-442    run_differential_correction
 *   differential_orbit_correction/single_iteration.rs:140-317
 *   differential_orbit_correction/least_square.rs:188-405
 *   differential_orbit_correction/outlier_rejection.rs:118-235
 *   orbit_type/equinoctial_element.rs:258-270 (is_bizarre), :442-637 (compute_derivative),
 *   :639-759, :809-867 (propagate_twobody with derivatives)
 *   ephemeris/observation_ephemeris.rs:204-258 (angles + position partials), :322-342, :418-450
 * and, from the un-vendored nalgebra 0.34.2, the published algorithms of Cholesky::new /
 * Cholesky::inverse (left-looking, column axpy; forward substitution column by column, adjoint
 * substitution with sequential dots) and QR::new / QR::try_inverse (Householder reflections with the
 * double normalisation of the axis).
 *
 * Pins: equinoctial_element.rs:1317-1420 (compute_derivative, exact), least_square.rs:437-724
 * (normal equations, covariance, rescaling, the OrbFit min_sol vector at 1e-10),
 * outlier_rejection.rs:274-540 (selection rules).  The reference's end-to-end LSQ tests need DE440 /
 * UT1 downloads: the full loop is "parity unpinned" beyond those unit KATs.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oo.h"
#include "oo_linalg.h"

#define M6(m, r, c) ((m)[6 * (c) + (r)])

static const double ROT_ECL2EQU_L[9] = {1.0, 0.0, 0.0,
                                        0.0, 0.9174820620691818, 0.3977771559319137,
                                        0.0, -0.3977771559319137, 0.9174820620691818};

void oo_lsq_config_default(oo_lsq_config *c) { /* diff_cor.rs:175-192, outlier_rejection.rs:74-81,
                                                  equinoctial_element.rs:169-179 */
  c->max_newton_iterations = 30;
  c->max_outlier_rejection_passes = 10;
  c->convergence_threshold = 1e-4;
  c->convergence_before_rejection_threshold = 2.0;
  c->rms_stagnation_ratio = 0.98;
  c->rms_divergence_ratio = 1.5;
  c->max_stagnation_iterations = 3;
  c->enable_outlier_rejection = 1;
  c->chi2_rejection_threshold = 25.0;
  c->chi2_recovery_threshold = 9.0;
  c->eccentricity_limit = 1.2;
  c->min_semi_major_axis = 1e-6;
  c->max_semi_major_axis = 1e4;
  c->min_periapsis_distance = 1e-6;
  c->max_apoapsis_distance = 1e4;
  for (int j = 0; j < 6; j++) c->free_elements[j] = 1;
}

/* equinoctial_element.rs:258-270 */
int oo_is_bizarre(const double eq[6], const oo_lsq_config *c) {
  double e = sqrt(eq[1] * eq[1] + eq[2] * eq[2]);
  double peri = eq[0] * (1.0 - e);
  double apo = eq[0] * (1.0 + e);
  return e > c->eccentricity_limit || eq[0] < c->min_semi_major_axis || eq[0] > c->max_semi_major_axis ||
         peri < c->min_periapsis_distance || apo > c->max_apoapsis_distance;
}

/* equinoctial_element.rs:442-637.  dpos / dvel: [3][6] = column c of the reference's Matrix6x3
   (d component c / d element j at [6*c + j]). */
void oo_compute_derivative(const double eq[6], double t0, double t1, double n, double lam1, double F,
                           double inv_u, double beta, double sF, double cF, double xe, double ye,
                           double vxe, double vye, const double fv[3], const double gv[3],
                           const double pos[3], const double vel[3], double dpos[18], double dvel[18]) {
  const double mu = OO_GAUSS_GRAV * OO_GAUSS_GRAV;
  double a = eq[0], h = eq[1], k = eq[2], p = eq[3], q = eq[4];
  double wv[3] = {2.0 * p * inv_u, -2.0 * q * inv_u, (1.0 - p * p - q * q) * inv_u};
  double r = sqrt(xe * xe + ye * ye);
  double inv_r = 1.0 / r;
  double inv_1_beta = 1.0 / (1.0 - beta);
  double b3 = (beta * beta) * beta;
  double tmp1 = lam1 - F;
  double tmp2 = beta + (h * h) * b3 * inv_1_beta;
  double tmp3 = h * k * b3 * inv_1_beta;
  double tmp4 = beta * h - sF;
  double tmp5 = beta * k - cF;
  double tmp6 = beta + (k * k) * b3 * inv_1_beta;
  double tmp7 = 1.0 - r / a;
  double tmp8 = sF - h;
  double tmp9 = cF - k;
  double tmp10 = a * cF * inv_r;
  double tmp11 = a * sF * inv_r;
  double tmp12 = n * (a * a) * inv_r;
  double dt = t1 - t0;
  double col[6][3], colv[6][3];
  for (int c = 0; c < 3; c++) col[0][c] = (pos[c] - 3.0 * vel[c] * dt / 2.0) / a;
  double dx1de2 = -a * (tmp1 * tmp2 + a * cF * tmp4 * inv_r);
  double dx2de2 = a * (tmp1 * tmp3 - 1.0 + a * cF * tmp5 * inv_r);
  for (int c = 0; c < 3; c++) col[1][c] = dx1de2 * fv[c] + dx2de2 * gv[c];
  double dx1de3 = -a * (tmp1 * tmp3 + 1.0 - a * sF * tmp4 * inv_r);
  double dx2de3 = a * (tmp1 * tmp6 - a * sF * tmp5 * inv_r);
  for (int c = 0; c < 3; c++) col[2][c] = dx1de3 * fv[c] + dx2de3 * gv[c];
  for (int c = 0; c < 3; c++) col[3][c] = 2.0 * (q * (ye * fv[c] - xe * gv[c]) - xe * wv[c]) * inv_u;
  for (int c = 0; c < 3; c++) col[4][c] = 2.0 * (p * (-ye * fv[c] + xe * gv[c]) + ye * wv[c]) * inv_u;
  for (int c = 0; c < 3; c++) col[5][c] = vel[c] / n;
  double r3 = (r * r) * r;
  for (int c = 0; c < 3; c++) colv[0][c] = -(vel[c] - 3.0 * mu * pos[c] * dt / r3) / (2.0 * a);
  double ir2 = inv_r * inv_r, a2 = a * a;
  double dx4de2 = tmp12 * (tmp7 * tmp2 + a2 * tmp8 * tmp4 * ir2 + tmp10 * cF);
  double dx5de2 = -tmp12 * (tmp7 * tmp3 + a2 * tmp8 * tmp5 * ir2 - tmp10 * sF);
  for (int c = 0; c < 3; c++) colv[1][c] = dx4de2 * fv[c] + dx5de2 * gv[c];
  double dx4de3 = tmp12 * (tmp7 * tmp3 + a2 * tmp9 * tmp4 * ir2 - tmp11 * cF);
  double dx5de3 = -tmp12 * (tmp7 * tmp6 + a2 * tmp9 * tmp5 * ir2 + tmp11 * sF);
  for (int c = 0; c < 3; c++) colv[2][c] = dx4de3 * fv[c] + dx5de3 * gv[c];
  for (int c = 0; c < 3; c++) colv[3][c] = 2.0 * (q * (vye * fv[c] - vxe * gv[c]) - vxe * wv[c]) * inv_u;
  for (int c = 0; c < 3; c++) colv[4][c] = 2.0 * (p * (-vye * fv[c] + vxe * gv[c]) + vye * wv[c]) * inv_u;
  double ir3 = (inv_r * inv_r) * inv_r, a3 = (a * a) * a;
  for (int c = 0; c < 3; c++) colv[5][c] = -n * a3 * pos[c] * ir3;
  for (int j = 0; j < 6; j++)
    for (int c = 0; c < 3; c++) { dpos[6 * c + j] = col[j][c]; dvel[6 * c + j] = colv[j][c]; }
}

/* equinoctial_element.rs:809-867 + :639-759 with compute_derivatives = true */
int oo_propagate_twobody_partials(const oo_elements *eq, double t0, double t1, double pos[3],
                                  double vel[3], double dpos[18], double dvel[18]) {
  const double mu = OO_GAUSS_GRAV * OO_GAUSS_GRAV;
  double a = eq->e[0], h = eq->e[1], k = eq->e[2], p = eq->e[3], q = eq->e[4];
  double n = sqrt(mu / ((a * a) * a));
  double lam1 = eq->e[5] + n * (t1 - t0);
  double e2 = h * h + k * k;
  double epsilon = OO_EPS * 1e2;
  double lon_peri = 0.0;
  if (e2 > epsilon) lon_peri = oo_rem_euclid(atan2(h, k), OO_DPI);
  lam1 = oo_rem_euclid(lam1, OO_DPI);
  if (lam1 < lon_peri) lam1 += OO_DPI;
  double F;
  int rc = oo_equinoctial_solve_kepler(eq, lam1, lon_peri, &F);
  if (rc != OO_OK) return rc;
  double beta = 1.0 / (1.0 + sqrt(1.0 - e2));
  double bhk = beta * h * k;
  double sF = sin(F), cF = cos(F);
  double xe = a * ((1.0 - beta * (h * h)) * cF + bhk * sF - k);
  double ye = a * ((1.0 - beta * (k * k)) * sF + bhk * cF - h);
  double u = 1.0 + p * p + q * q;
  double inv_u = 1.0 / u;
  double common = 2.0 * p * q * inv_u;
  double fv[3] = {(1.0 - p * p + q * q) * inv_u, common, -2.0 * p * inv_u};
  double gv[3] = {common, (1.0 + p * p - q * q) * inv_u, 2.0 * q * inv_u};
  for (int i = 0; i < 3; i++) pos[i] = xe * fv[i] + ye * gv[i];
  double vconst = n * (a * a) / sqrt(xe * xe + ye * ye);
  double vxe = vconst * (bhk * cF - (1.0 - beta * (h * h)) * sF);
  double vye = vconst * ((1.0 - beta * (k * k)) * cF - bhk * sF);
  for (int i = 0; i < 3; i++) vel[i] = vxe * fv[i] + vye * gv[i];
  oo_compute_derivative(eq->e, t0, t1, n, lam1, F, inv_u, beta, sF, cF, xe, ye, vxe, vye, fv, gv, pos,
                        vel, dpos, dvel);
  return OO_OK;
}

static void topocentric_partials(const double pe[3], const double ve[3], const double dpos[18], const double obs[3],
                                 double *ra, double *dec, double d_ra[6], double d_dec[6]);
/* observation_ephemeris.rs:418-450 (compute_obs_and_partials_2body), :204-258, :322-342 */
int oo_obs_and_partials(const oo_traj_view *tv, size_t i, const oo_ephem_table *tab,
                        const oo_elements *equi, double *ra, double *dec, double d_ra[6],
                        double d_dec[6]) {
  double h = equi->e[1], k = equi->e[2];
  if (sqrt(h * h + k * k) >= 1.0) return OO_ERR_INVALID_ORBIT;
  double dt = tv->mjd_tt[i] - equi->epoch;
  double pe[3], ve[3], dpos[18], dvel[18];
  int rc = oo_propagate_twobody_partials(equi, 0.0, dt, pe, ve, dpos, dvel);
  if (rc != OO_OK) return rc;
  double obs[3];
  if (tv->scorer_obs_equ) {
    memcpy(obs, &tv->scorer_obs_equ[3 * i], sizeof obs);
  } else {
    rc = oo_scorer_observer_position(tab, tv->mjd_tt[i], &tv->geo_ecl[3 * i], obs);
    if (rc != OO_OK) return rc;
  }
  topocentric_partials(pe, ve, dpos, obs, ra, dec, d_ra, d_dec);
  return OO_OK;
}

/* topocentric_radec_and_partials + element_partials_from_position_partials (observation_ephemeris.rs:204-258, :322-342) */
static void topocentric_partials(const double pe[3], const double ve[3], const double dpos[18], const double obs[3],
                                 double *ra, double *dec, double d_ra[6], double d_dec[6]) {
  const double vlight_au = 2.99792458e5 / OO_AU * 86400.0;
  double ap[3], av[3], rel[3], cor[3];
  oo_matvec(ROT_ECL2EQU_L, pe, ap);
  oo_matvec(ROT_ECL2EQU_L, ve, av);
  for (int c = 0; c < 3; c++) rel[c] = ap[c] - obs[c];
  double ltt = oo_norm3(rel) / vlight_au;
  for (int c = 0; c < 3; c++) cor[c] = rel[c] - ltt * av[c];
  double x = cor[0], y = cor[1], z = cor[2];
  double rho = oo_norm3(cor);
  double rho_xy = hypot(x, y);
  double rho_xy_sq = rho_xy * rho_xy;
  *dec = atan2(z, rho_xy);
  *ra = oo_rem_euclid(atan2(y, x), OO_DPI);
  double rho_sq = rho * rho;
  double gra[3] = {-y / rho_xy_sq, x / rho_xy_sq, 0.0};
  double gdec[3] = {-z * x / (rho_xy * rho_sq), -z * y / (rho_xy * rho_sq), rho_xy / rho_sq};
  double rel_norm = oo_norm3(rel);
  double aberr = 1.0 / (rel_norm * vlight_au);
  double sra = oo_dot3(gra, av) * aberr, sdec = oo_dot3(gdec, av) * aberr;
  double drp[3], ddp[3];
  for (int c = 0; c < 3; c++) {
    drp[c] = gra[c] - sra * rel[c];
    ddp[c] = gdec[c] - sdec * rel[c];
  }
  for (int j = 0; j < 6; j++) {
    double de[3] = {dpos[j], dpos[6 + j], dpos[12 + j]}, dq[3];
    oo_matvec(ROT_ECL2EQU_L, de, dq);
    d_ra[j] = oo_dot3(drp, dq);
    d_dec[j] = oo_dot3(ddp, dq);
  }
}

/* PropagatorKind::NBody inside the differential correction (single_iteration.rs:186-191): the frozen perturbers of the
 * trajectory in hand (PerturberSnapshot at the elements' reference epoch, nbody.rs:453-476) and the integrator
 * tolerances.  Thread-local: oo_fit_lsq_nbody sets it per trajectory; n_pert == 0 <=> PropagatorKind::TwoBody. */
static __thread struct { const oo_perturber *pert; size_t n_pert; double atol, rtol; } nb_ctx = {NULL, 0, 0.0, 0.0};
void oo_lsq_set_nbody(const oo_perturber *pert, size_t n_pert, double atol, double rtol) {
  nb_ctx.pert = pert; nb_ctx.n_pert = n_pert; nb_ctx.atol = atol; nb_ctx.rtol = rtol;
}

/* compute_obs_and_partials_nbody (observation_ephemeris.rs:452-486) on EquinoctialElements::propagate_nbody
 * (equinoctial_element.rs:908-968): state and element Jacobian J0 at the reference epoch, DOP853 on [r, v, Phi], then
 * d pos(t1) / d elements = the top three rows of Phi(t1) J0 (nbody.rs:552-604). */
int oo_obs_and_partials_nbody(const oo_traj_view *tv, size_t i, const oo_ephem_table *tab, const oo_elements *equi,
                              const oo_perturber *pert, size_t n_pert, double atol, double rtol, double *ra,
                              double *dec, double d_ra[6], double d_dec[6]) {
  double h = equi->e[1], k = equi->e[2];
  if (sqrt(h * h + k * k) >= 1.0) return OO_ERR_INVALID_ORBIT;
  double p0[3], v0[3], dpos0[18], dvel0[18];
  int rc = oo_propagate_twobody_partials(equi, 0.0, 0.0, p0, v0, dpos0, dvel0);
  if (rc != OO_OK) return rc;
  double span = tv->mjd_tt[i] - equi->epoch;
  double pe[3], ve[3], dpos[18];
  if (fabs(span) < 1e-14) {
    memcpy(pe, p0, sizeof pe); memcpy(ve, v0, sizeof ve); memcpy(dpos, dpos0, sizeof dpos);
  } else {
    double y[42];
    memset(y, 0, sizeof y);
    for (int c = 0; c < 3; c++) { y[c] = p0[c]; y[3 + c] = v0[c]; }
    for (int c = 0; c < 6; c++) y[6 + 7 * c] = 1.0;
    rc = oo_dop853_nbody(y, span, pert, n_pert, atol, rtol, 100000u, NULL, NULL);
    if (rc != OO_OK) return rc;
    for (int c = 0; c < 3; c++) { pe[c] = y[c]; ve[c] = y[3 + c]; }
    /* (Phi J0)[c][j] = sum_k Phi[c][k] J0[k][j], Phi column-major at y[6 + 6 k + c]; J0[k][j] = d state_k / d element j */
    for (int c = 0; c < 3; c++)
      for (int j = 0; j < 6; j++) {
        double acc = y[6 + c] * dpos0[j];
        for (int kk = 1; kk < 6; kk++) {
          double jk = kk < 3 ? dpos0[6 * kk + j] : dvel0[6 * (kk - 3) + j];
          acc = y[6 + 6 * kk + c] * jk + acc;
        }
        dpos[6 * c + j] = acc;
      }
  }
  double obs[3];
  if (tv->scorer_obs_equ) {
    memcpy(obs, &tv->scorer_obs_equ[3 * i], sizeof obs);
  } else {
    rc = oo_scorer_observer_position(tab, tv->mjd_tt[i], &tv->geo_ecl[3 * i], obs);
    if (rc != OO_OK) return rc;
  }
  topocentric_partials(pe, ve, dpos, obs, ra, dec, d_ra, d_dec);
  return OO_OK;
}

/* least_square.rs:188-199 */
double oo_angular_diff(double a, double b) {
  double d = a - b;
  while (d > OO_PI) d -= OO_DPI;
  while (d < -OO_PI) d += OO_DPI;
  return d;
}

/* nalgebra 0.34 Cholesky::new (linalg/cholesky.rs): in-place lower factor; 0 = not positive definite */
static int cholesky6(double m[36]) {
  for (int j = 0; j < 6; j++) {
    for (int k = 0; k < j; k++) {
      double factor = -M6(m, j, k);
      for (int i = j; i < 6; i++) M6(m, i, j) = factor * M6(m, i, k) + M6(m, i, j);
    }
    double diag = M6(m, j, j);
    if (diag == 0.0 || !(diag >= 0.0)) return 0;
    double denom = sqrt(diag);
    M6(m, j, j) = denom;
    for (int i = j + 1; i < 6; i++) M6(m, i, j) = M6(m, i, j) / denom;
  }
  return 1;
}
/* Cholesky::inverse = solve_mut(identity): L then L^T substitution (linalg/solve.rs) */
static void cholesky6_inverse(const double l[36], double inv[36]) {
  for (int c = 0; c < 6; c++) {
    double b[6] = {0, 0, 0, 0, 0, 0};
    b[c] = 1.0;
    for (int i = 0; i < 5; i++) {
      double coeff = b[i] / M6(l, i, i);
      b[i] = coeff;
      for (int r = i + 1; r < 6; r++) b[r] = -coeff * M6(l, r, i) + b[r];
    }
    b[5] = b[5] / M6(l, 5, 5);
    for (int i = 5; i >= 0; i--) {
      double dot = 0.0;
      for (int r = i + 1; r < 6; r++) dot += M6(l, r, i) * b[r];
      b[i] = (b[i] - dot) / M6(l, i, i);
    }
    for (int r = 0; r < 6; r++) M6(inv, r, c) = b[r];
  }
}
/* nalgebra 0.34 QR::new + try_inverse (linalg/qr.rs, householder.rs, geometry/reflection.rs) */
static int qr6_inverse(const double m_in[36], double inv[36]) {
  double m[36], diag[6];
  memcpy(m, m_in, sizeof m);
  for (int ic = 0; ic < 6; ic++) {
    /* reflection_axis_mut on rows ic.. of column ic */
    double sq = 0.0;
    for (int r = ic; r < 6; r++) sq += M6(m, r, ic) * M6(m, r, ic);
    double nrm = sqrt(sq);
    double x0 = M6(m, ic, ic);
    double modulus = x0 >= 0.0 ? x0 : -x0;
    double sgn = x0 >= 0.0 ? 1.0 : -1.0;
    double signed_norm = sgn * nrm;
    double factor = (sq + modulus * nrm) * 2.0;
    M6(m, ic, ic) = x0 + signed_norm;
    if (factor != 0.0) {
      double sf = sqrt(factor);
      for (int r = ic; r < 6; r++) M6(m, r, ic) = M6(m, r, ic) / sf;
      double n2 = 0.0;
      for (int r = ic; r < 6; r++) n2 += M6(m, r, ic) * M6(m, r, ic);
      double nn = sqrt(n2);
      for (int r = ic; r < 6; r++) M6(m, r, ic) = M6(m, r, ic) / nn;
      double rn = -signed_norm;
      diag[ic] = rn;
      double sign = signbit(rn) ? -1.0 : 1.0; /* f64::signum */
      double m_two = sign * -2.0;
      for (int c = ic + 1; c < 6; c++) {
        double dot = 0.0;
        for (int r = ic; r < 6; r++) dot += M6(m, r, ic) * M6(m, r, c);
        double fac = (dot - 0.0) * m_two;
        for (int r = ic; r < 6; r++) M6(m, r, c) = fac * M6(m, r, ic) + sign * M6(m, r, c);
      }
    } else {
      diag[ic] = signed_norm;
    }
  }
  for (int c = 0; c < 6; c++) {
    double b[6] = {0, 0, 0, 0, 0, 0};
    b[c] = 1.0;
    for (int i = 0; i < 6; i++) { /* q_tr_mul: reflect_with_sign(diag[i].signum()) */
      double sign = signbit(diag[i]) ? -1.0 : 1.0;
      double dot = 0.0;
      for (int r = i; r < 6; r++) dot += M6(m, r, i) * b[r];
      double fac = (dot - 0.0) * (sign * -2.0);
      for (int r = i; r < 6; r++) b[r] = fac * M6(m, r, i) + sign * b[r];
    }
    for (int i = 5; i >= 0; i--) { /* solve_upper_triangular_mut */
      double d = fabs(diag[i]);
      if (d == 0.0) return 0;
      double coeff = b[i] / d;
      b[i] = coeff;
      for (int r = 0; r < i; r++) b[r] = -coeff * M6(m, r, i) + b[r];
    }
    for (int r = 0; r < 6; r++) M6(inv, r, c) = b[r];
  }
  return 1;
}
/* least_square.rs:329-342 */
int oo_invert_normal_matrix(const double m[36], double inv[36]) {
  double l[36];
  memcpy(l, m, sizeof l);
  if (cholesky6(l)) { cholesky6_inverse(l, inv); return 1; }
  if (qr6_inverse(m, inv)) return 1;
  memset(inv, 0, 36 * sizeof(double));
  return 0;
}

/* least_square.rs:225-327 */
void oo_solve_weighted_least_squares(size_t n, const oo_obs_equation *eqs, const int32_t free_elements[6],
                                     oo_lsq_solution *out) {
  size_t active = 0;
  for (size_t i = 0; i < n; i++) active += eqs[i].active ? 1 : 0;
  out->num_measurements = 2 * active;
  double *nm = out->normal_matrix;
  double rhs[6] = {0, 0, 0, 0, 0, 0};
  memset(nm, 0, 36 * sizeof(double));
  double q = 0.0;
  for (size_t i = 0; i < n; i++) {
    const oo_obs_equation *e = &eqs[i];
    if (!e->active) continue;
    const double *pr = e->d_ra, *pd = e->d_dec;
    double wr = e->weight_ra, wd = e->weight_dec, wc = e->weight_cross;
    double xr = e->residual_ra, xd = e->residual_dec;
    for (int j = 0; j < 6; j++) {
      for (int k = 0; k < 6; k++)
        M6(nm, j, k) += pr[j] * wr * pr[k] + pd[j] * wd * pd[k] + wc * (pd[j] * pr[k] + pr[j] * pd[k]);
      rhs[j] += (pr[j] * wr + pd[j] * wc) * xr + (pr[j] * wc + pd[j] * wd) * xd;
    }
    q += wr * xr * xr + wd * xd * xd + 2.0 * wc * xr * xd;
  }
  for (int j = 0; j < 6; j++)
    if (!free_elements[j]) {
      for (int k = 0; k < 6; k++) { M6(nm, j, k) = 0.0; M6(nm, k, j) = 0.0; }
      M6(nm, j, j) = 1.0;
      rhs[j] = 0.0;
    }
  out->inversion_succeeded = oo_invert_normal_matrix(nm, out->covariance);
  for (int r = 0; r < 6; r++) out->correction[r] = 0.0;
  if (out->inversion_succeeded) { /* Matrix6 * Vector6: gemv column by column */
    for (int r = 0; r < 6; r++) out->correction[r] = M6(out->covariance, r, 0) * rhs[0];
    for (int c = 1; c < 6; c++)
      for (int r = 0; r < 6; r++) out->correction[r] = M6(out->covariance, r, c) * rhs[c] + out->correction[r];
  }
  for (int j = 0; j < 6; j++)
    if (!free_elements[j]) out->correction[j] = 0.0;
  out->normalised_rms = out->num_measurements > 0 ? sqrt(q / (double)out->num_measurements) : 0.0;
}

/* least_square.rs:371-394 */
void oo_rescale_covariance(double normal_matrix[36], double covariance[36], size_t num_free,
                           size_t num_measurements, double normalised_rms) {
  double mu = 1.0;
  if (num_free < num_measurements) {
    double factor = sqrt((double)num_measurements / (double)(num_measurements - num_free));
    mu = normalised_rms > 1.0 ? normalised_rms * factor : factor;
  }
  double mu2 = mu * mu;
  for (int i = 0; i < 36; i++) { covariance[i] *= mu2; normal_matrix[i] /= mu2; }
}

static void gemv6(const double m[36], const double v[6], double out[6]) {
  for (int r = 0; r < 6; r++) out[r] = M6(m, r, 0) * v[0];
  for (int c = 1; c < 6; c++)
    for (int r = 0; r < 6; r++) out[r] = M6(m, r, c) * v[c] + out[r];
}
static double dot6(const double a[6], const double b[6]) {
  double res = 0.0;
  for (int i = 0; i < 6; i++) res += a[i] * b[i];
  return res;
}

/* outlier_rejection.rs:118-235 ; selection: 0 Active, 1 Rejected, 2 ForcedOut ; returns #changes */
size_t oo_update_observation_selection(size_t n, oo_obs_fit_data *fit, const oo_obs_equation *eqs,
                                       const double covariance[36], double chi2_reject,
                                       double chi2_recover) {
  size_t changes = 0;
  for (size_t i = 0; i < n; i++) {
    oo_obs_fit_data *f = &fit[i];
    const oo_obs_equation *e = &eqs[i];
    if (f->selection == 2) continue;
    double var_ra = f->sigma_ra * f->sigma_ra;
    double var_dec = f->sigma_dec * f->sigma_dec;
    double cov_cross = -f->sigma_ra * f->sigma_dec * e->weight_cross / (e->weight_ra * e->weight_dec);
    double gga[6], ggd[6];
    gemv6(covariance, e->d_ra, gga);
    gemv6(covariance, e->d_dec, ggd);
    double paa = dot6(e->d_ra, gga), pdd = dot6(e->d_dec, ggd), pad = dot6(e->d_ra, ggd);
    double v00 = var_ra - paa, v01 = cov_cross - pad, v11 = var_dec - pdd;
    double det = v00 * v11 - v01 * v01;
    double scale = fmax(fabs(v00), fabs(v11));
    double thr = OO_EPS * scale * scale;
    if (fabs(det) < thr || scale == 0.0) continue;
    double i00 = v11 / det, i01 = -v01 / det, i10 = -v01 / det, i11 = v00 / det;
    /* Matrix2 * Vector2 (column axpy) then Vector2 dot (a0*b0 + a1*b1) */
    double y0 = i00 * f->residual_ra, y1 = i10 * f->residual_ra;
    y0 = i01 * f->residual_dec + y0;
    y1 = i11 * f->residual_dec + y1;
    double chi2 = f->residual_ra * y0 + f->residual_dec * y1;
    if (f->selection == 0 && chi2 > chi2_reject) { f->selection = 1; changes++; }
    else if (f->selection == 1 && chi2 <= chi2_recover) { f->selection = 0; changes++; }
  }
  return changes;
}

/* single_iteration.rs:140-317 (apply_correction = true; PropagatorKind::TwoBody, or NBody when oo_lsq_set_nbody is on) */
static void single_iteration(const oo_traj_view *tv, const oo_ephem_table *tab, const oo_obs_fit_data *fit,
                             const oo_elements *el, const int32_t free_elements[6], oo_obs_equation *eqs,
                             oo_obs_fit_data *fit_out, oo_lsq_solution *sol, oo_elements *corrected,
                             double *correction_norm) {
  for (size_t i = 0; i < tv->n; i++) {
    oo_obs_equation *e = &eqs[i];
    fit_out[i] = fit[i];
    double ra, dec;
    int ok = fit[i].selection == 0 &&
             (nb_ctx.n_pert ? oo_obs_and_partials_nbody(tv, i, tab, el, nb_ctx.pert, nb_ctx.n_pert, nb_ctx.atol, nb_ctx.rtol, &ra,
                                                        &dec, e->d_ra, e->d_dec)
                            : oo_obs_and_partials(tv, i, tab, el, &ra, &dec, e->d_ra, e->d_dec)) == OO_OK;
    if (!ok) {
      memset(e, 0, sizeof *e);
      e->weight_ra = 1.0 / (1.0 * 1.0);
      e->weight_dec = 1.0 / (1.0 * 1.0);
      continue;
    }
    double rra = oo_angular_diff(tv->ra[i] - fit[i].bias_ra, ra);
    double rdec = (tv->dec[i] - fit[i].bias_dec) - dec;
    double ca = rra / fit[i].sigma_ra, cd = rdec / fit[i].sigma_dec;
    e->residual_ra = rra;
    e->residual_dec = rdec;
    e->weight_ra = 1.0 / (fit[i].sigma_ra * fit[i].sigma_ra);
    e->weight_dec = 1.0 / (fit[i].sigma_dec * fit[i].sigma_dec);
    e->weight_cross = 0.0;
    e->active = 1;
    fit_out[i].residual_ra = rra;
    fit_out[i].residual_dec = rdec;
    fit_out[i].chi = sqrt(ca * ca + cd * cd);
  }
  oo_solve_weighted_least_squares(tv->n, eqs, free_elements, sol);
  double cdx[6];
  gemv6(sol->normal_matrix, sol->correction, cdx);
  *correction_norm = sqrt(dot6(sol->correction, cdx));
  *corrected = *el;
  for (int j = 0; j < 6; j++)
    if (free_elements[j]) corrected->e[j] = el->e[j] + sol->correction[j];
}

/* diff_cor.rs:282-442 ; returns OO_OK or OO_ERR_LSQ_* ; out filled on OO_OK */
int oo_run_differential_correction(const oo_traj_view *tv, const oo_ephem_table *tab,
                                   const oo_elements *initial, const oo_lsq_config *cfg,
                                   oo_obs_fit_data *fit /* in: initial, out: final (n) */,
                                   oo_lsq_result *out) {
  size_t n = tv->n;
  size_t num_free = 0;
  for (int j = 0; j < 6; j++) num_free += cfg->free_elements[j] ? 1 : 0;
  oo_elements el = *initial;
  oo_obs_equation *eqs = (oo_obs_equation *)calloc(n ? n : 1, sizeof *eqs);
  oo_obs_equation *last_eqs = (oo_obs_equation *)calloc(n ? n : 1, sizeof *eqs);
  oo_obs_fit_data *fit_new = (oo_obs_fit_data *)calloc(n ? n : 1, sizeof *fit_new);
  size_t last_n_eqs = 0;
  double last_nm[36], last_cov[36];
  memset(last_nm, 0, sizeof last_nm);
  memset(last_cov, 0, sizeof last_cov);
  double last_rms = 1.7976931348623157e308;
  size_t last_nmeas = 0;
  uint64_t total_it = 0;
  int rc = OO_OK;
  for (uint64_t outer = 0; outer <= cfg->max_outlier_rejection_passes; outer++) {
    double prev_rms = 1.7976931348623157e308;
    uint64_t stagnation = 0;
    int converged = 0;
    last_n_eqs = 0; /* `let mut last_equations = vec![]` */
    for (uint64_t inner = 0; inner < cfg->max_newton_iterations; inner++) {
      total_it++;
      oo_lsq_solution sol;
      oo_elements corrected;
      double cnorm;
      single_iteration(tv, tab, fit, &el, cfg->free_elements, eqs, fit_new, &sol, &corrected, &cnorm);
      if (!sol.inversion_succeeded) { rc = OO_ERR_LSQ_INVERSION; goto done; }
      if (oo_is_bizarre(corrected.e, cfg)) { rc = OO_ERR_LSQ_BIZARRE; goto done; }
      double new_rms = sol.normalised_rms;
      if (prev_rms < 1.7976931348623157e308 && new_rms / prev_rms >= cfg->rms_divergence_ratio) {
        rc = OO_ERR_LSQ_DIVERGED; goto done;
      }
      int stagnated = prev_rms < 1.7976931348623157e308 && new_rms / prev_rms >= cfg->rms_stagnation_ratio;
      if (stagnated) {
        stagnation++;
        if (stagnation >= cfg->max_stagnation_iterations) break;
      } else {
        stagnation = 0;
      }
      memcpy(last_eqs, eqs, n * sizeof *eqs);
      last_n_eqs = n;
      memcpy(last_nm, sol.normal_matrix, sizeof last_nm);
      memcpy(last_cov, sol.covariance, sizeof last_cov);
      last_rms = new_rms;
      last_nmeas = sol.num_measurements;
      el = corrected;
      memcpy(fit, fit_new, n * sizeof *fit);
      prev_rms = new_rms;
      if (cnorm < cfg->convergence_threshold) { converged = 1; break; }
    }
    if (!cfg->enable_outlier_rejection) break;
    if (outer == 0 && last_rms < cfg->convergence_before_rejection_threshold) break;
    if (!converged) break;
    /* update_observation_selection asserts equal lengths: an empty last_equations with n > 0 would
       panic in the reference; unreachable (a converged inner loop always advanced once). */
    if (last_n_eqs != n) break;
    size_t changes = oo_update_observation_selection(n, fit, last_eqs, last_cov, cfg->chi2_rejection_threshold,
                                                     cfg->chi2_recovery_threshold);
    if (changes == 0) break;
  }
  oo_rescale_covariance(last_nm, last_cov, num_free, last_nmeas, last_rms);
  out->epoch = el.epoch;
  memcpy(out->elem, el.e, sizeof out->elem);
  memcpy(out->normal_matrix, last_nm, sizeof last_nm);
  memcpy(out->covariance, last_cov, sizeof last_cov);
  for (int j = 0; j < 6; j++) out->sigma[j] = sqrt(M6(last_cov, j, j)); /* uncertainty.rs:261-271 */
  out->normalised_rms = last_rms;
  out->total_newton_iterations = total_it;
  out->num_measurements = last_nmeas;
done:
  out->total_newton_iterations = total_it;
  free(eqs); free(last_eqs); free(fit_new);
  return rc;
}

/* mod.rs:60-115 with initial_orbits = Some(IOD result): `iod` is the trajectory's IOD outcome */
void oo_differential_correction(const oo_traj_view *tv, const oo_ephem_table *tab, const oo_iod_result *iod,
                                const oo_lsq_config *cfg, oo_lsq_result *out, oo_obs_fit_data *fit) {
  memset(out, 0, sizeof *out);
  for (size_t i = 0; i < tv->n; i++) { /* obs_fit_data.rs:105-116 */
    memset(&fit[i], 0, sizeof fit[i]);
    fit[i].sigma_ra = tv->sigma_ra[i];
    fit[i].sigma_dec = tv->sigma_dec[i];
  }
  if (iod->status != OO_OK) { out->status = iod->status; out->kind = OO_LSQ_KIND_NONE; return; }
  oo_elements in, eq;
  in.kind = iod->element_kind;
  in.epoch = iod->epoch;
  memcpy(in.e, iod->elem, sizeof in.e);
  int rc = oo_to_equinoctial(&in, &eq);
  if (rc != OO_OK) { out->status = rc; out->kind = OO_LSQ_KIND_NONE; return; }
  rc = oo_run_differential_correction(tv, tab, &eq, cfg, fit, out);
  out->status = OO_OK;
  out->fallback_cause = rc;
  if (rc == OO_OK) { out->kind = OO_LSQ_KIND_CORRECTED; return; }
  /* Err(_) => Ok(initial_orbit): the IOD result is returned unchanged (mod.rs:113) */
  out->kind = OO_LSQ_KIND_IOD_FALLBACK;
  out->epoch = iod->epoch;
  memcpy(out->elem, iod->elem, sizeof out->elem);
  out->normalised_rms = iod->rms;
  for (size_t i = 0; i < tv->n; i++) { /* the fit data of a failed run is dropped */
    memset(&fit[i], 0, sizeof fit[i]);
    fit[i].sigma_ra = tv->sigma_ra[i];
    fit[i].sigma_dec = tv->sigma_dec[i];
  }
}

/* keplerian_element.rs:185-233 (KeplerianElements::from_equinoctial_internal): (a,h,k,p,q,lambda) ->
   (a,e,i,Omega,omega,M) */
void oo_equinoctial_to_keplerian(const double eq[6], double kep[6]) {
  const double eps = 1.0e-12;
  double h = eq[1], k = eq[2], p = eq[3], q = eq[4];
  double ecc = sqrt(h * h + k * k);
  double dig = ecc < eps ? 0.0 : atan2(h, k);
  double tgi2 = sqrt(p * p + q * q);
  double node = tgi2 < eps ? 0.0 : atan2(p, q);
  kep[0] = eq[0];
  kep[1] = ecc;
  kep[2] = 2.0 * atan(tgi2);
  kep[3] = node;
  kep[4] = oo_rem_euclid(dig - node, OO_DPI);
  kep[5] = oo_rem_euclid(eq[5] - dig, OO_DPI);
}

/* equinoctial_element.rs:1049-1140 (jacobian_to_keplerian): d(a,e,i,Omega,omega,M)/d(a,h,k,p,q,lambda),
   column-major 6x6 */
void oo_jacobian_to_keplerian(const double eq[6], double jac[36]) {
  const double eps = 1.0e-12;
  double h = eq[1], k = eq[2], p = eq[3], q = eq[4];
  double e = sqrt(h * h + k * k), e_sq = e * e;
  double dvh = 0.0, dvk = 0.0;
  if (!(e < eps)) { dvh = k / e_sq; dvk = -h / e_sq; }
  double t = sqrt(p * p + q * q), t_sq = t * t;
  double dip = 0.0, diq = 0.0, dnp = 0.0, dnq = 0.0;
  if (!(t < eps)) {
    double denom = t * (1.0 + t_sq);
    dip = 2.0 * p / denom; diq = 2.0 * q / denom;
    dnp = q / t_sq; dnq = -p / t_sq;
  }
  double em = fmax(e, eps);
  memset(jac, 0, 36 * sizeof(double));
  M6(jac, 0, 0) = 1.0;
  M6(jac, 1, 1) = h / em; M6(jac, 4, 1) = dvh; M6(jac, 5, 1) = -dvh;
  M6(jac, 1, 2) = k / em; M6(jac, 4, 2) = dvk; M6(jac, 5, 2) = -dvk;
  M6(jac, 2, 3) = dip; M6(jac, 3, 3) = dnp; M6(jac, 4, 3) = -dnp;
  M6(jac, 2, 4) = diq; M6(jac, 3, 4) = dnq; M6(jac, 4, 4) = -dnq;
  M6(jac, 5, 5) = 1.0;
}

/* uncertainty.rs:412-416 (OrbitalCovariance::propagate): J * C * J^T, nalgebra's column-by-column gemv */
void oo_propagate_covariance(const double jac[36], const double cov[36], double out[36]) {
  double jc[36], jt[36];
  for (int c = 0; c < 6; c++) gemv6(jac, &cov[6 * c], &jc[6 * c]);
  for (int r = 0; r < 6; r++)
    for (int c = 0; c < 6; c++) M6(jt, r, c) = M6(jac, c, r);
  for (int c = 0; c < 6; c++) gemv6(jc, &jt[6 * c], &out[6 * c]);
}
