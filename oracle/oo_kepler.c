/* oo_kepler.c -- ORACLE (test infrastructure only): universal-variable Kepler machinery.
 * Restates src/kepler/{stumpff,newton_solver,brent_dekker_solver,params,velocity,propagation}.rs
 * and src/kepler/prelim_kepler/{prelim_elliptic,prelim_hyperbolic,prelim_parabolic}.rs of the
 * reference, operation for operation. */
#include <math.h>
#include <string.h>
#include "oo.h"
#include "oo_linalg.h"

/* params.rs:35-73 (SolverParams::default, SolverType::default) */
void oo_kepler_params_default_solver(oo_kepler_params *p) {
  p->kind = OO_SOLVER_NEWTON;
  p->convergency = 100.0 * OO_EPS;
  p->has_psi_guess = 0;
  p->psi_guess = 0.0;
  p->max_iter_prelim_kepuni = 20;
  p->parabolic_method = OO_PARABOLIC_CARDANO;
}

/* ------------------------------------------------------------------ stumpff.rs */
/* stumpff.rs:135-191 */
static void stumpff_power_series(double psi, double psi2, double beta, double alpha, double tol,
                                 double overflow, int max_terms, double s[4]) {
  double s2 = 0.5 * psi2;
  double term2 = s2;
  double s3 = (s2 * psi) / 3.0;
  double term3 = s3;
  double d2lo = 3.0, d2hi = 4.0, d3lo = 4.0, d3hi = 5.0;
  for (int it = 1; it <= max_terms; it++) {
    oo_tls_cnt.sfunct_terms++;
    term2 *= beta / (d2lo * d2hi);
    s2 += term2;
    term3 *= beta / (d3lo * d3hi);
    s3 += term3;
    int neg2 = fabs(term2) < tol, neg3 = fabs(term3) < tol;
    int div2 = fabs(term2) > overflow, div3 = fabs(term3) > overflow;
    if ((neg2 && neg3) || div2 || div3) break;
    d2lo += 2.0; d2hi += 2.0; d3lo += 2.0; d3hi += 2.0;
  }
  s[1] = psi + alpha * s3;
  s[0] = 1.0 + alpha * s2;
  s[2] = s2;
  s[3] = s3;
}

/* stumpff.rs:201-297 */
static void stumpff_halving(double psi, double beta, double alpha, double tol, double overflow,
                            double beta_threshold, int max_halving, int max_terms, double s[4]) {
  double rpsi = psi, rbeta = beta;
  int halvings = 0;
  while (fabs(rbeta) >= beta_threshold && halvings < max_halving) {
    rpsi *= 0.5;
    rbeta *= 0.25;
    halvings++;
  }
  double s0 = 1.0, s1 = rpsi, t0 = 1.0, t1 = rpsi;
  for (int k = 1; k <= max_terms; k++) {
    oo_tls_cnt.sfunct_terms++;
    t0 *= rbeta / ((double)(2 * k - 1) * (double)(2 * k));
    s0 += t0;
    if (fabs(t0) < tol || fabs(t0) > overflow) break;
  }
  for (int k = 1; k <= max_terms; k++) {
    oo_tls_cnt.sfunct_terms++;
    t1 *= rbeta / ((double)(2 * k) * (double)(2 * k + 1));
    s1 += t1;
    if (fabs(t1) < tol || fabs(t1) > overflow) break;
  }
  for (int h = 0; h < halvings; h++) {
    double c = s0, sn = s1;
    s0 = 2.0 * c * c - 1.0;
    s1 = 2.0 * c * sn;
  }
  s[3] = (s1 - psi) / alpha;
  s[2] = (s0 - 1.0) / alpha;
  s[0] = s0;
  s[1] = s1;
}

/* stumpff.rs:78-126 */
void oo_s_funct(double psi, double alpha, double s[4]) {
  const double tol = 100.0 * OO_EPS;
  const double overflow = 1.0 / OO_EPS;
  oo_tls_cnt.sfunct_calls++;
  if (psi == 0.0) {
    s[0] = 1.0; s[1] = 0.0; s[2] = 0.0; s[3] = 0.0;
    return;
  }
  double psi2 = psi * psi;
  double beta = alpha * psi2;
  if (fabs(beta) < 100.0)
    stumpff_power_series(psi, psi2, beta, alpha, tol, overflow, 70, s);
  else
    stumpff_halving(psi, beta, alpha, tol, overflow, 100.0, 30, 70, s);
}

/* ------------------------------------------------------- prelim_elliptic.rs */
static double principal_angle(double a) { return oo_rem_euclid(a, OO_DPI); } /* angles.rs:15 */

/* prelim_elliptic.rs:72-134 */
double oo_prelim_elliptic(const oo_kepler_params *p) {
  oo_tls_cnt.prelim_calls++;
  double contr = p->convergency;
  uint64_t max_iter = p->max_iter_prelim_kepuni;
  double a0 = -1.0 / p->alpha;
  double n = sqrt(p->mu) * sqrt(-((p->alpha * p->alpha) * p->alpha));
  if (p->e0 < contr) return n * p->dt / sqrt(-p->alpha);
  /* initial_eccentric_anomaly_from_geometry :9-31 */
  double cosu = (1.0 - p->r0 / a0) / p->e0;
  double u0;
  if (fabs(cosu) <= 1.0) u0 = acos(cosu);
  else if (cosu >= 1.0) u0 = 0.0;
  else u0 = OO_PI;
  if (p->sig0 < 0.0) u0 = -u0;
  u0 = principal_angle(u0);
  double m0 = principal_angle(u0 - p->e0 * sin(u0));
  double target = m0 + n * p->dt;
  /* solve_elliptic_kepler_equation :113-134 */
  double u = target;
  for (uint64_t i = 0; i < max_iter; i++) {
    oo_tls_cnt.prelim_steps++;
    double res = u - p->e0 * sin(u) - target;
    double der = 1.0 - p->e0 * cos(u);
    double step = -res / der;
    u += step;
    if (fabs(step) < contr * 1e3) break;
  }
  return (u - u0) / sqrt(-p->alpha);
}

/* prelim_hyperbolic.rs:45-141 */
double oo_prelim_hyperbolic(const oo_kepler_params *p) {
  oo_tls_cnt.prelim_calls++;
  double a0 = -1.0 / p->alpha;
  double n = sqrt(p->mu) * sqrt((p->alpha * p->alpha) * p->alpha);
  double coshf = (1.0 - p->r0 / a0) / p->e0;
  double f0;
  if (coshf > 1.0) f0 = log(coshf + sqrt(coshf * coshf - 1.0));
  else f0 = 0.0;
  if (p->sig0 < 0.0) f0 = -f0;
  double m0 = p->e0 * sinh(f0) - f0;
  double target = m0 + n * p->dt;
  double f = 0.0;
  double thr = p->convergency;
  for (uint64_t i = 0; i < p->max_iter_prelim_kepuni; i++) {
    oo_tls_cnt.prelim_steps++;
    if (fabs(f) < 15.0) {
      double res = p->e0 * sinh(f) - f - target;
      double der = p->e0 * cosh(f) - 1.0;
      double step = -res / der;
      double cand = f + step;
      f = (f * cand < 0.0) ? f / 2.0 : cand;
    } else {
      f /= 2.0;
    }
    if (fabs(f) < thr * 1e3) break;
  }
  return (f - f0) / sqrt(p->alpha);
}

/* ------------------------------------------------------ prelim_parabolic.rs */
/* :149-165 */
static void cubic_res_der(double psi, double r0, double sig0, double sdt, double *res, double *der) {
  *res = ((psi * psi) * psi) / 6.0 + sig0 / 2.0 * (psi * psi) + r0 * psi - sdt;
  *der = (psi * psi) / 2.0 + sig0 * psi + r0;
}
/* :438-477 ; min_by keeps the first minimum */
static double closest_to(const double *roots, int n, double target) {
  int best = 0;
  for (int i = 1; i < n; i++)
    if (fabs(roots[i] - target) < fabs(roots[best] - target)) best = i;
  return roots[best];
}
static double select_physical_root(const double *roots, int n, double r0, double sig0, double sdt) {
  double lin = sdt / r0;
  double mono[3];
  int nm = 0;
  for (int i = 0; i < n; i++) {
    double res, der;
    cubic_res_der(roots[i], r0, sig0, sdt, &res, &der);
    if (der >= 0.0) mono[nm++] = roots[i];
  }
  return nm == 0 ? closest_to(roots, n, lin) : closest_to(mono, nm, lin);
}
/* :264-338 */
static double prelim_parabolic_cardano(const oo_kepler_params *p) {
  double r0 = p->r0, sig0 = p->sig0, dt = p->dt;
  double sdt = sqrt(p->mu) * dt;
  if (dt == 0.0) return 0.0;
  double lead = 1.0 / 6.0;
  double b = (sig0 / 2.0) / lead;
  double c = r0 / lead;
  double d = -sdt / lead;
  double shift = b / 3.0;
  double pp = c - b * shift;
  double qq = 2.0 * ((shift * shift) * shift) - c * shift + d;
  double halfq = qq / 2.0;
  double p3 = pp / 3.0;
  double disc = halfq * halfq + (p3 * p3) * p3;
  double roots[3];
  int nr;
  if (disc > 0.0) {
    double sq = sqrt(disc);
    double y = cbrt(-halfq + sq) + cbrt(-halfq - sq);
    roots[0] = y - shift;
    nr = 1;
  } else {
    /* three_real_roots_trigonometric :403-418 */
    double arg = oo_clamp((3.0 * qq) / (2.0 * pp) * sqrt(-3.0 / pp), -1.0, 1.0);
    double base = acos(arg) / 3.0;
    double amp = 2.0 * sqrt(-pp / 3.0);
    roots[0] = amp * cos(base) - shift;
    roots[1] = amp * cos(base - 2.0 * OO_PI / 3.0) - shift;
    roots[2] = amp * cos(base - 4.0 * OO_PI / 3.0) - shift;
    nr = 3;
  }
  double psi = select_physical_root(roots, nr, r0, sig0, sdt);
  /* polish_root_by_newton :371-393 */
  for (int i = 0; i < 2; i++) {
    double res, der;
    cubic_res_der(psi, r0, sig0, sdt, &res, &der);
    if (der == 0.0 || !isfinite(der)) break;
    psi -= res / der;
  }
  return psi;
}
/* :198-246 */
static double prelim_parabolic_newton(const oo_kepler_params *p) {
  double r0 = p->r0, sig0 = p->sig0, dt = p->dt;
  double sdt = sqrt(p->mu) * dt;
  double contr = p->convergency;
  if (dt == 0.0) return 0.0;
  if (sig0 * sig0 > 2.0 * r0) return prelim_parabolic_cardano(p);
  double psi = sdt / r0;
  for (uint64_t i = 0; i < p->max_iter_prelim_kepuni; i++) {
    double res, der;
    cubic_res_der(psi, r0, sig0, sdt, &res, &der);
    if (!isfinite(der) || fabs(der) < 10.0 * OO_EPS) {
      psi *= 0.5;
      continue;
    }
    double raw = -res / der;
    double mx = 2.0 * (1.0 + fabs(psi));
    double step = oo_clamp(raw, -mx, mx);
    psi += step;
    if (fabs(step) < contr) break;
  }
  return psi;
}
double oo_prelim_parabolic(const oo_kepler_params *p) {
  oo_tls_cnt.prelim_calls++;
  return p->parabolic_method == OO_PARABOLIC_CARDANO ? prelim_parabolic_cardano(p)
                                                     : prelim_parabolic_newton(p);
}

/* params.rs:185-191 + orbit_type.rs:38-44 */
int oo_prelim_kepuni(const oo_kepler_params *p, double *psi) {
  if (p->alpha < 0.0) *psi = oo_prelim_elliptic(p);
  else if (p->alpha > 0.0) *psi = oo_prelim_hyperbolic(p);
  else *psi = oo_prelim_parabolic(p);
  return 1;
}

/* ---------------------------------------------------------- newton_solver.rs */
static void set_solution(oo_kepler_solution *o, double psi, const double s[4]) {
  o->psi = psi; o->s0 = s[0]; o->s1 = s[1]; o->s2 = s[2]; o->s3 = s[3];
}
/* newton_solver.rs:151-161, 240-352 */
int oo_solve_kepuni_newton(const oo_kepler_params *p, oo_kepler_solution *out) {
  double tol = 10.0 * OO_EPS * (1.0 + fabs(sqrt(p->mu) * p->dt));
  double psi;
  if (p->has_psi_guess) psi = p->psi_guess;
  else if (!oo_prelim_kepuni(p, &psi)) return 0;
  for (int it = 0; it < 50; it++) {
    oo_tls_cnt.newton_steps++;
    if (!isfinite(psi)) { psi = 0.5; continue; }
    double s[4];
    oo_s_funct(psi, p->alpha, s);
    double res = p->r0 * s[1] + p->sig0 * s[2] + s[3] - sqrt(p->mu) * p->dt;
    double der = p->r0 * s[0] + p->sig0 * s[1] + s[2];
    if (fabs(res) <= tol) { set_solution(out, psi, s); return 1; }
    if (!isfinite(der) || fabs(der) < 10.0 * OO_EPS) { psi *= 0.5; continue; }
    double raw = -res / der;
    double mx = 2.0 * (1.0 + fabs(psi));
    double step = oo_clamp(raw, -mx, mx);
    double cand = psi + step;
    if (cand * psi < 0.0) cand = 0.5 * psi;
    psi = cand;
    double sa = fabs(step);
    if (sa <= p->convergency) { set_solution(out, psi, s); return 1; }
    if (sa <= p->convergency * (1.0 + fabs(psi))) {
      double sf[4];
      oo_s_funct(psi, p->alpha, sf);
      set_solution(out, psi, sf);
      return 1;
    }
  }
  return 0;
}

/* ---------------------------------------------------- brent_dekker_solver.rs */
static double kep_residual(double psi, const oo_kepler_params *p) { /* :59-62 */
  double s[4];
  oo_tls_cnt.brent_evals++;
  oo_s_funct(psi, p->alpha, s);
  return p->r0 * s[1] + p->sig0 * s[2] + s[3] - sqrt(p->mu) * p->dt;
}
/* brent_dekker_solver.rs:469-526 */
int oo_solve_kepuni_brent(const oo_kepler_params *p, oo_kepler_solution *out) {
  const double PHI = 1.618033988749895;
  double psi0;
  if (p->has_psi_guess) psi0 = p->psi_guess;
  else if (!oo_prelim_kepuni(p, &psi0)) return 0;
  /* bracket_kepler_root :150-177 */
  double hw = fabs(psi0) > 1.0 ? fabs(psi0) : 1.0;
  double lo = psi0 - hw, hi = psi0 + hw;
  double flo = kep_residual(lo, p), fhi = kep_residual(hi, p);
  int found = 0;
  for (int it = 0; it < 60; it++) {
    if (flo * fhi <= 0.0) { found = 1; break; }
    double w = hi - lo;
    if (fabs(flo) < fabs(fhi)) {
      lo = lo - PHI * w;
      flo = kep_residual(lo, p);
    } else {
      hi = hi + PHI * w;
      fhi = kep_residual(hi, p);
    }
  }
  if (!found) return 0;
  /* BrentState::from_bracket :214-233 */
  double a = lo, fa = kep_residual(lo, p);
  double b = hi, fb = kep_residual(hi, p);
  if (fabs(fa) < fabs(fb)) {
    double t = a; a = b; b = t;
    t = fa; fa = fb; fb = t;
  }
  double c = a, fc = fa;
  double prev_step = fabs(hi - lo);
  int prev_bis = 1;
  double tol = p->convergency;
  for (int it = 0; it < 100; it++) {
    if (fabs(fb) <= tol || 0.5 * fabs(b - a) <= tol) {
      double s[4];
      oo_s_funct(b, p->alpha, s);
      set_solution(out, b, s);
      return 1;
    }
    /* select_next_candidate :325-345 */
    double interp;
    if (fabs(fa - fc) > OO_EPS && fabs(fb - fc) > OO_EPS) {
      double ta = a * fb * fc / ((fa - fb) * (fa - fc));
      double tb = b * fa * fc / ((fb - fa) * (fb - fc));
      double tc = c * fa * fb / ((fc - fa) * (fc - fb));
      interp = ta + tb + tc;
    } else {
      interp = b - fb * (b - a) / (fb - fa);
    }
    double ref_len = prev_bis ? fabs(b - c) : prev_step;
    double tq = (3.0 * a + b) / 4.0;
    int inside = (tq < b) ? (interp > tq && interp < b) : (interp > b && interp < tq);
    int progress = fabs(interp - b) < 0.5 * ref_len;
    double next;
    int was_bis;
    if (inside && progress) { next = interp; was_bis = 0; }
    else { next = 0.5 * (a + b); was_bis = 1; }
    double fnext = kep_residual(next, p);
    prev_step = fabs(b - c);
    prev_bis = was_bis;
    /* update_bracket :352-369 */
    c = b; fc = fb;
    if (fa * fnext < 0.0) { b = next; fb = fnext; }
    else { a = next; fa = fnext; }
    if (fabs(fa) < fabs(fb)) {
      double t = a; a = b; b = t;
      t = fa; fa = fb; fb = t;
    }
  }
  return 0;
}

/* params.rs:130-142 */
int oo_kepler_solve(const oo_kepler_params *p, oo_kepler_solution *out) {
  oo_tls_cnt.kepler_universal_solves++;
  switch (p->kind) {
    case OO_SOLVER_NEWTON:
      return oo_solve_kepuni_newton(p, out) ? OO_OK : OO_ERR_NEWTON_KEPLER;
    case OO_SOLVER_BRENT:
      return oo_solve_kepuni_brent(p, out) ? OO_OK : OO_ERR_BRENT_KEPLER;
    default:
      if (oo_solve_kepuni_newton(p, out)) return OO_OK;
      return oo_solve_kepuni_brent(p, out) ? OO_OK : OO_ERR_BRENT_KEPLER;
  }
}

/* ---------------------------------------------------------------- velocity.rs */
/* velocity.rs:94-211 */
int oo_velocity_correction_with_guess(const double x1[3], const double x2[3], const double v2[3],
                                      double dt, double peri_max, double ecc_max, int has_guess,
                                      double chi_guess, double eps, double v_out[3], double *f,
                                      double *g, double *chi) {
  const double mu = OO_GAUSS_GRAV * OO_GAUSS_GRAV;
  double r2 = oo_norm3(x2);
  double sig0 = oo_dot3(x2, v2) / sqrt(mu);
  double h[3];
  oo_cross3(x2, v2, h);
  double hn = oo_norm3(h);
  if (!isfinite(hn) || hn <= 1e6 * OO_EPS) return OO_ERR_VELOCITY_CORRECTION;
  int acc;
  double ecc, peri, energy;
  if (!oo_eccentricity_control(x2, v2, peri_max, ecc_max, &acc, &ecc, &peri, &energy))
    return OO_ERR_VELOCITY_CORRECTION;
  oo_kepler_params kp;
  oo_kepler_params_default_solver(&kp);
  kp.dt = dt;
  kp.r0 = r2;
  kp.sig0 = sig0;
  kp.mu = mu;
  kp.alpha = 2.0 * energy / mu;
  kp.e0 = ecc;
  kp.convergency = eps;
  kp.has_psi_guess = has_guess;
  kp.psi_guess = chi_guess;
  oo_kepler_solution sol;
  int rc = oo_kepler_solve(&kp, &sol);
  if (rc != OO_OK) return rc;
  double fc = 1.0 - sol.s2 / r2;
  double gc = dt - sol.s3 / sqrt(mu);
  double ga = fabs(gc);
  double gmin = 100.0 * OO_EPS * (1.0 + fabs(dt));
  if (!isfinite(ga) || ga < gmin) return OO_ERR_VELOCITY_CORRECTION;
  for (int i = 0; i < 3; i++) {
    double t = (-fc) * x2[i] + x1[i]; /* axpy(-f, x2, 1.0) */
    v_out[i] = t / gc;                /* unscale_mut(g)    */
  }
  *f = fc;
  *g = gc;
  *chi = sol.psi;
  return OO_OK;
}

/* ------------------------------------------------------------- propagation.rs */
/* propagation.rs:114-207 */
int oo_propagate_universal(const double r[3], const double v[3], double t0, double t1, int kind,
                           double convergency, double out[11]) {
  oo_tls_cnt.propagate_universal_calls++;
  const double mu = OO_GAUSS_GRAV * OO_GAUSS_GRAV;
  double r0 = oo_norm3(r);
  if (r0 < OO_EPS) return OO_ERR_DEGENERATE_STATE;
  /* initial_orbital_state :190-207 */
  double v2 = oo_dot3(v, v);
  double sig0 = oo_dot3(r, v) / sqrt(mu);
  double alpha = (v2 - 2.0 * mu / r0) / mu;
  double h[3];
  oo_cross3(r, v, h);
  double h2 = oo_dot3(h, h);
  double e0 = sqrt(1.0 + alpha * h2 / mu);
  e0 = (e0 != e0) ? 0.0 : (e0 > 0.0 ? e0 : 0.0); /* f64::max(NaN, 0.0) == 0.0 */
  double tof = t1 - t0;
  double sqrt_mu = sqrt(mu);
  oo_kepler_params kp;
  oo_kepler_params_default_solver(&kp);
  kp.r0 = r0; kp.sig0 = sig0; kp.mu = mu; kp.alpha = alpha; kp.dt = tof; kp.e0 = e0;
  kp.kind = kind;
  kp.convergency = convergency;
  oo_kepler_solution sol;
  int rc = oo_kepler_solve(&kp, &sol);
  if (rc != OO_OK) return rc;
  double r1 = r0 * sol.s0 + sig0 * sol.s1 + sol.s2;
  if (r1 < OO_EPS) return OO_ERR_DEGENERATE_STATE;
  double fl = 1.0 - sol.s2 / r0;
  double gl = (r0 * sol.s1 + sig0 * sol.s2) / sqrt_mu;
  double fd = -(sqrt_mu / (r0 * r1)) * sol.s1;
  double gd = 1.0 - sol.s2 / r1;
  for (int i = 0; i < 3; i++) {
    out[i] = fl * r[i] + gl * v[i];
    out[3 + i] = fd * r[i] + gd * v[i];
  }
  out[6] = fl; out[7] = gl; out[8] = fd; out[9] = gd; out[10] = sol.psi;
  return OO_OK;
}
