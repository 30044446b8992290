/* oo_gauss.c -- ORACLE (test infrastructure only): Gauss preliminary orbit on one observation
 * triplet.  Restates src/initial_orbit_determination/gauss.rs and the behaviour of the
 * un-vendored crate aberth 0.4.1 (Cargo.lock:6-7) that gauss.rs:655,971 calls.
 *
 * aberth restatement (pinned by gauss.rs:1543-1570: root ORDER and bit-exact values):
 *   - monic coefficients; a = -c_{n-1}/n; P(w) by binomial shift; S(w) = |p_n| w^n - sum |p_i| w^i;
 *     r0 = smallest positive integer with S(r0) > 0;  z_k = a + r0 e^{i theta_k},
 *     theta_k = (2 pi / n) k + (pi / 2) / n.
 *   - Jacobi sweeps: w_i = p(z_i) / (p(z_i) * sum_{k!=i} 1/(z_i - z_k) - p'(z_i)), z_i += w_i,
 *     Horner evaluation through num-complex's MulAdd (fused real parts), complex division by the
 *     textbook formula (no Smith scaling).
 *   - Converged when every root moved by < eps in re and im within one sweep; the UPDATED set is
 *     returned; NaN/Inf -> Failed; after max_iter sweeps -> MaxIteration (current set returned). */
#include <math.h>
#include <string.h>
#include "oo.h"
#include "oo_linalg.h"

typedef struct { double re, im; } cx;
static inline cx cx_mul(cx a, cx b) { cx r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; return r; }
static inline cx cx_add(cx a, cx b) { cx r = {a.re + b.re, a.im + b.im}; return r; }
static inline cx cx_sub(cx a, cx b) { cx r = {a.re - b.re, a.im - b.im}; return r; }
static inline cx cx_div(cx a, cx b) {
  double n = b.re * b.re + b.im * b.im;
  cx r = {(a.re * b.re + a.im * b.im) / n, (a.im * b.re - a.re * b.im) / n};
  return r;
}
/* num-complex 0.4 `impl MulAdd for Complex<T>` */
static inline cx cx_mul_add(cx s, cx o, cx a) {
  cx r;
  r.re = fma(s.re, o.re, a.re) - (s.im * o.im);
  r.im = fma(s.re, o.im, fma(s.im, o.re, a.im));
  return r;
}
static cx horner(const double *c, int n_terms, cx x) {
  cx r = {0.0, 0.0};
  for (int i = n_terms - 1; i >= 0; i--) {
    cx ci = {c[i], 0.0};
    r = cx_mul_add(r, x, ci);
  }
  return r;
}

int oo_aberth8(const double poly[9], uint32_t max_iter, double eps, double re[8], double im[8],
               uint32_t *sweeps) {
  enum { N = 8 };
  oo_tls_cnt.aberth_solves++;
  double dydx[N];
  for (int i = 1; i <= N; i++) dydx[i - 1] = poly[i] * (double)i;
  /* initial guesses */
  double monic[N + 1];
  for (int i = 0; i <= N; i++) monic[i] = poly[i] / poly[N];
  double a = -monic[N - 1] / (double)N;
  for (int ci = 0; ci <= N; ci++) {
    double c = monic[ci];
    monic[ci] = 0.0;
    double binom = 1.0; /* Pascal row ci */
    for (int idx = 0; idx <= ci; idx++) {
      int power = ci - idx;
      double ap = 1.0;
      for (int q = 0; q < power; q++) ap *= a; /* a.powi(power) */
      monic[idx] = fma(c, binom * ap, monic[idx]);
      binom = binom * (double)(ci - idx) / (double)(idx + 1);
    }
  }
  double s_of_w[N + 1];
  for (int i = 0; i < N; i++) s_of_w[i] = -fabs(monic[i]);
  s_of_w[N] = fabs(monic[N]);
  double r0 = 1.0;
  for (int guard = 0; guard < 100000; guard++) {
    cx x = {r0, 0.0};
    cx v = horner(s_of_w, N + 1, x);
    if (v.re > 0.0) break;
    r0 += 1.0;
  }
  cx zs[N], nz[N];
  for (int k = 0; k < N; k++) {
    double theta = (OO_DPI / (double)N) * (double)k + (OO_PI / 2.0) / (double)N;
    zs[k].re = a + r0 * cos(theta);
    zs[k].im = r0 * sin(theta);
  }
  int status = 1;
  uint32_t it = 0;
  for (it = 0; it < max_iter; it++) {
    oo_tls_cnt.aberth_sweeps++;
    int converged = 1;
    for (int i = 0; i < N; i++) {
      cx pz = horner(poly, N + 1, zs[i]);
      cx dz = horner(dydx, N, zs[i]);
      cx sum = {0.0, 0.0};
      for (int k = 0; k < N; k++) {
        if (k == i) continue;
        cx one = {1.0, 0.0};
        sum = cx_add(sum, cx_div(one, cx_sub(zs[i], zs[k])));
      }
      cx z = cx_add(zs[i], cx_div(pz, cx_sub(cx_mul(pz, sum), dz)));
      nz[i] = z;
      if (isnan(z.re) || isnan(z.im) || isinf(z.re) || isinf(z.im)) {
        if (sweeps) *sweeps = it + 1;
        for (int k2 = 0; k2 < N; k2++) { re[k2] = zs[k2].re; im[k2] = zs[k2].im; }
        return 2;
      }
      if (!(fabs(z.re - zs[i].re) < eps && fabs(z.im - zs[i].im) < eps)) converged = 0;
    }
    memcpy(zs, nz, sizeof zs);
    if (converged) { status = 0; it++; break; }
  }
  if (sweeps) *sweeps = it;
  for (int k = 0; k < N; k++) { re[k] = zs[k].re; im[k] = zs[k].im; }
  return status;
}

/* gauss.rs:648-667 */
int oo_solve_8poly(const double poly[9], uint32_t max_iter, double aberth_eps, double root_eps,
                   double roots[8], int *n_roots) {
  double re[8], im[8];
  int st = oo_aberth8(poly, max_iter, aberth_eps, re, im, NULL);
  *n_roots = 0;
  if (st == 2) return OO_ERR_POLY_ROOT_FAILED;
  for (int k = 0; k < 8; k++)
    if (re[k] > 0.0 && fabs(im[k]) < root_eps) roots[(*n_roots)++] = re[k];
  return OO_OK;
}

/* gauss.rs:464-502 */
static void unit_matrix(const oo_gauss_obs *g, double s[9]) {
  for (int c = 0; c < 3; c++) {
    double cd = cos(g->dec[c]);
    OO_M(s, 0, c) = cos(g->ra[c]) * cd;
    OO_M(s, 1, c) = sin(g->ra[c]) * cd;
    OO_M(s, 2, c) = sin(g->dec[c]);
  }
}

/* gauss.rs:532-549 */
int oo_gauss_prelim(const oo_gauss_obs *g, double *tau1, double *tau3, double unit[9],
                    double inv_unit[9], double a[3], double b[3]) {
  double t1 = OO_GAUSS_GRAV * (g->t[0] - g->t[1]);
  double t3 = OO_GAUSS_GRAV * (g->t[2] - g->t[1]);
  double t13 = t3 - t1;
  a[0] = t3 / t13; a[1] = -1.0; a[2] = -(t1 / t13);
  b[0] = a[0] * (t13 * t13 - t3 * t3) / 6.0;
  b[1] = 0.0;
  b[2] = a[2] * (t13 * t13 - t1 * t1) / 6.0;
  unit_matrix(g, unit);
  *tau1 = t1;
  *tau3 = t3;
  if (!oo_inverse3(unit, inv_unit)) return OO_ERR_SINGULAR_DIRECTION_MATRIX;
  return OO_OK;
}

/* gauss.rs:585-614 */
void oo_coeff_eight_poly(const oo_gauss_obs *g, const double unit[9], const double inv_unit[9],
                         const double a[3], const double b[3], double c630[3]) {
  double ra[3], rb[3];
  oo_matvec(g->obs_pos, a, ra);
  oo_matvec(g->obs_pos, b, rb);
  double row1[3] = {OO_M(inv_unit, 1, 0), OO_M(inv_unit, 1, 1), OO_M(inv_unit, 1, 2)};
  double a2star = oo_dot3(row1, ra);
  double b2star = oo_dot3(row1, rb);
  /* observer_position_t.row(1) == column 1 of observer_helio_position */
  const double *r2 = &g->obs_pos[3];
  double r22 = ((0.0 + r2[0] * r2[0]) + r2[1] * r2[1]) + r2[2] * r2[2];
  double s2r2 = oo_dot3(&unit[3], r2);
  c630[0] = -(a2star * a2star) - r22 - (2.0 * a2star * s2r2);
  c630[1] = -(2.0 * b2star * (a2star + s2r2));
  c630[2] = -(b2star * b2star);
}

/* gauss.rs:214-240 */
static unsigned descartes_upper_bound(double c0, double c3, double c6, double zero_eps) {
  double v[3] = {c6, c3, c0};
  int last = 1; /* leading +1 */
  unsigned count = 0;
  for (int i = 0; i < 3; i++) {
    int cur;
    if (fabs(v[i]) <= zero_eps) cur = 0;
    else cur = signbit(v[i]) ? -1 : 1;
    if (cur == 0) continue;
    if (last != 0 && cur != last) count++;
    last = cur;
  }
  return count;
}

/* gauss.rs:702-724 */
int oo_position_vector_and_reference_epoch(const oo_gauss_obs *g, const oo_iod_params *p,
                                           const double unit[9], const double inv_unit[9],
                                           const double c[3], double pos[9], double *epoch) {
  const double vlight_au = 2.99792458e5 / OO_AU * 86400.0; /* constants.rs:79 */
  double gcap[3], crhom[3], rho[3];
  oo_matvec(g->obs_pos, c, gcap);
  oo_matvec(inv_unit, gcap, crhom);
  for (int i = 0; i < 3; i++) rho[i] = -(crhom[i] / c[i]);
  if (rho[1] < p->min_rho2_au) return OO_ERR_SPURIOUS_ROOT;
  for (int col = 0; col < 3; col++)
    for (int r = 0; r < 3; r++)
      OO_M(pos, r, col) = OO_M(g->obs_pos, r, col) + rho[col] * OO_M(unit, r, col);
  *epoch = g->t[1] - rho[1] / vlight_au;
  return OO_OK;
}

/* gauss.rs:754-781 */
void oo_gibbs_correction(const double pos[9], double tau1, double tau3, double v[3]) {
  double tau13 = tau3 - tau1;
  double n1 = oo_norm3(&pos[0]), n2 = oo_norm3(&pos[3]), n3 = oo_norm3(&pos[6]);
  double r1m3 = 1.0 / ((n1 * n1) * n1);
  double r2m3 = 1.0 / ((n2 * n2) * n2);
  double r3m3 = 1.0 / ((n3 * n3) * n3);
  double d1 = tau3 * (r1m3 / 12.0 - 1.0 / (tau1 * tau13));
  double d2 = (tau1 + tau3) * (r2m3 / 12.0 - 1.0 / (tau1 * tau3));
  double d3 = -tau1 * (r3m3 / 12.0 + 1.0 / (tau3 * tau13));
  double d[3] = {-d1, d2, d3};
  for (int r = 0; r < 3; r++) {
    double row[3] = {OO_M(pos, r, 0), OO_M(pos, r, 1), OO_M(pos, r, 2)};
    v[r] = OO_GAUSS_GRAV * oo_dot3(row, d);
  }
}

/* gauss.rs:816-870 ; returns 1 (Some) / 0 (None) */
static int accept_root(const oo_gauss_obs *g, const oo_iod_params *p, double root,
                       const double unit[9], const double inv_unit[9], const double a[3],
                       const double b[3], double tau1, double tau3, double pos[9], double vel[3],
                       double *epoch) {
  double r2m3 = 1.0 / ((root * root) * root);
  double c[3] = {a[0] + b[0] * r2m3, -1.0, a[2] + b[2] * r2m3};
  if (oo_position_vector_and_reference_epoch(g, p, unit, inv_unit, c, pos, epoch) != OO_OK) return 0;
  oo_gibbs_correction(pos, tau1, tau3, vel);
  int acc;
  double e, q, en;
  if (!oo_eccentricity_control(&pos[3], vel, p->max_perihelion_au, p->max_ecc, &acc, &e, &q, &en))
    return 0;
  return acc;
}

/* gauss.rs:1284-1418 ; returns 1 (Some) / 0 (None) */
int oo_pos_and_vel_correction(const oo_gauss_obs *g, const oo_iod_params *p, const double pos_in[9],
                              const double vel_in[3], const double unit[9], const double inv_unit[9],
                              double peri_max, double ecc_max, double err_max, uint64_t itmax,
                              double pos[9], double vel[3], double *epoch_out) {
  memcpy(pos, pos_in, 9 * sizeof(double));
  memcpy(vel, vel_in, 3 * sizeof(double));
  double epoch = 0.0;
  int has01 = 0, has21 = 0;
  double chi01 = 0.0, chi21 = 0.0;
  double dt01 = g->t[0] - g->t[1];
  double dt21 = g->t[2] - g->t[1];
  if (fabs(dt01) <= OO_EPS || fabs(dt21) <= OO_EPS) return 0;
  for (uint64_t it = 0; it < itmax; it++) {
    oo_tls_cnt.fg_iterations++;
    const double *r1 = &pos[0], *r2 = &pos[3], *r3 = &pos[6];
    double v1[3], v2[3], f1, g1, c1, f2, g2, c2;
    int rcl = oo_velocity_correction_with_guess(r1, r2, vel, dt01, peri_max, ecc_max, has01, chi01,
                                                p->kepler_eps, v1, &f1, &g1, &c1);
    int rcr = oo_velocity_correction_with_guess(r3, r2, vel, dt21, peri_max, ecc_max, has21, chi21,
                                                p->kepler_eps, v2, &f2, &g2, &c2);
    if (rcl != OO_OK || rcr != OO_OK) continue;
    has01 = 1; chi01 = c1;
    has21 = 1; chi21 = c2;
    if (!isfinite(g1) || !isfinite(g2)) continue;
    double nv[3];
    for (int i = 0; i < 3; i++) nv[i] = (v1[i] + v2[i]) * 0.5;
    double fl = f1 * g2 - f2 * g1;
    if (!isfinite(fl) || fabs(fl) < OO_EPS) continue;
    double inv_f = 1.0 / fl;
    double c[3] = {g2 * inv_f, -1.0, -g1 * inv_f};
    double npos[9], nepoch;
    if (oo_position_vector_and_reference_epoch(g, p, unit, inv_unit, c, npos, &nepoch) != OO_OK)
      continue;
    int acc;
    double e, q, en;
    if (!oo_eccentricity_control(&npos[3], nv, peri_max, ecc_max, &acc, &e, &q, &en)) return 0;
    if (!acc) return 0;
    double denom = oo_matnorm(npos);
    if (!isfinite(denom) || denom <= OO_EPS) continue;
    double diff[9];
    for (int i = 0; i < 9; i++) diff[i] = npos[i] - pos[i];
    double rel = oo_matnorm(diff) / denom;
    memcpy(pos, npos, sizeof npos);
    memcpy(vel, nv, sizeof nv);
    epoch = nepoch;
    if (rel <= err_max) break;
  }
  *epoch_out = epoch;
  return 1;
}

/* gauss.rs:906-923, 1063-1076 */
static void build_result(const double pos[9], const double vel[3], double epoch, int corrected,
                         oo_gauss_result *out) {
  /* constants.rs:93-105 ROT_EQUMJ2000_TO_ECLMJ2000 (row-major literals in Matrix3::new) */
  static const double ROT[9] = {1.0, 0.0, 0.0,
                                0.0, 9.174820620691818e-1, -3.977771559319137e-1,
                                0.0, 3.977771559319137e-1, 9.174820620691818e-1}; /* column-major */
  double ep[3], ev[3];
  oo_matvec(ROT, &pos[3], ep);
  oo_matvec(ROT, vel, ev);
  oo_ccek1(ep, ev, epoch, &out->orbit);
  out->corrected = corrected;
}

/* gauss.rs:1119-1206 */
int oo_prelim_orbit_all(const oo_gauss_obs *g, const oo_iod_params *p, oo_gauss_result out[],
                        int cap, int *n_out) {
  double tau1, tau3, unit[9], inv_unit[9], a[3], b[3];
  *n_out = 0;
  oo_tls_cnt.gauss_solves++;
  int rc = oo_gauss_prelim(g, &tau1, &tau3, unit, inv_unit, a, b);
  if (rc != OO_OK) return rc;
  double c[3];
  oo_coeff_eight_poly(g, unit, inv_unit, a, b, c);
  double c6 = c[0], c3 = c[1], c0 = c[2];
  double poly[9] = {c0, 0.0, 0.0, c3, 0.0, 0.0, c6, 0.0, 1.0};
  if (descartes_upper_bound(c0, c3, c6, 0.0) == 0) return OO_ERR_GAUSS_NO_ROOTS;
  double re[8], im[8];
  int st = oo_aberth8(poly, p->aberth_max_iter, p->aberth_eps, re, im, NULL);
  if (st == 2) return OO_ERR_POLY_ROOT_FAILED;
  uint64_t max_sol = p->max_tested_solutions;
  int n = 0;
  for (int k = 0; k < 8; k++) {
    if (!(re[k] > 0.0 && fabs(im[k]) < p->root_imag_eps)) continue;
    double r2 = re[k];
    if (!(r2 >= p->r2_min_au && r2 <= p->r2_max_au)) continue;
    double pos[9], vel[3], epoch;
    if (!accept_root(g, p, r2, unit, inv_unit, a, b, tau1, tau3, pos, vel, &epoch)) continue;
    oo_tls_cnt.roots_accepted++;
    double cpos[9], cvel[3], cepoch;
    if (oo_pos_and_vel_correction(g, p, pos, vel, unit, inv_unit, p->max_perihelion_au, p->max_ecc,
                                  p->newton_eps, p->newton_max_it, cpos, cvel, &cepoch)) {
      if ((uint64_t)n < max_sol && n < cap) build_result(cpos, cvel, cepoch, 1, &out[n++]);
    } else {
      if ((uint64_t)n < max_sol && n < cap) build_result(pos, vel, epoch, 0, &out[n++]);
    }
    if ((uint64_t)n >= max_sol) break;
  }
  *n_out = n;
  return n == 0 ? OO_ERR_GAUSS_NO_ROOTS : OO_OK;
}

/* gauss.rs:1238-1247 */
int oo_prelim_orbit(const oo_gauss_obs *g, const oo_iod_params *p, oo_gauss_result *out) {
  oo_gauss_result all[8];
  int n;
  int rc = oo_prelim_orbit_all(g, p, all, 8, &n);
  if (rc != OO_OK) return rc;
  for (int i = 0; i < n; i++)
    if (all[i].corrected) { *out = all[i]; return OO_OK; }
  *out = all[0];
  return OO_OK;
}
