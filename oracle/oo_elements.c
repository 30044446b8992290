/* oo_elements.c -- ORACLE (test infrastructure only): state vector <-> orbital elements and the
 * equinoctial two-body propagator.  Restates src/orb_elem.rs, src/ref_system.rs::rotmt,
 * src/orbit_type/{mod,keplerian_element,cometary_element,equinoctial_element}.rs. */
#include <math.h>
#include "oo.h"
#include "oo_linalg.h"

/* ref_system.rs:453-462 -> nalgebra Rotation3::from_axis_angle on a principal axis.
 * angle == 0 yields the identity; otherwise the Rodrigues pattern below (active rotation). */
void oo_rotmt(double alpha, int axis, double m[9]) {
  for (int i = 0; i < 9; i++) m[i] = 0.0;
  OO_M(m, 0, 0) = OO_M(m, 1, 1) = OO_M(m, 2, 2) = 1.0;
  if (alpha == 0.0) return;
  double s = sin(alpha), c = cos(alpha);
  switch (axis) {
    case 0:
      OO_M(m, 1, 1) = c; OO_M(m, 1, 2) = -s;
      OO_M(m, 2, 1) = s; OO_M(m, 2, 2) = c;
      break;
    case 1:
      OO_M(m, 0, 0) = c; OO_M(m, 0, 2) = s;
      OO_M(m, 2, 0) = -s; OO_M(m, 2, 2) = c;
      break;
    default:
      OO_M(m, 0, 0) = c; OO_M(m, 0, 1) = -s;
      OO_M(m, 1, 0) = s; OO_M(m, 1, 1) = c;
      break;
  }
}

/* orb_elem.rs:257-301 */
int oo_eccentricity_control(const double r[3], const double v[3], double peri_max, double ecc_max,
                            int *accepted, double *ecc, double *peri, double *energy) {
  const double mu = OO_GAUSS_GRAV * OO_GAUSS_GRAV;
  oo_tls_cnt.ecc_controls++;
  double v2 = oo_dot3(v, v);
  double dist = oo_norm3(r);
  double h[3];
  oo_cross3(r, v, h);
  double h2 = oo_dot3(h, h);
  if (sqrt(h2) == 0.0) return 0;
  double vxh[3];
  oo_cross3(v, h, vxh);
  double inv_mu = 1.0 / mu, inv_d = 1.0 / dist;
  double lenz[3];
  for (int i = 0; i < 3; i++) lenz[i] = vxh[i] * inv_mu - r[i] * inv_d;
  double e = oo_norm3(lenz);
  double q = h2 / (mu * (1.0 + e));
  double en = v2 / 2.0 - mu / dist;
  *accepted = (e < ecc_max) && (q < peri_max);
  *ecc = e;
  *peri = q;
  *energy = en;
  return 1;
}

static double wrap_0_2pi(double x) { return oo_rem_euclid(x, OO_DPI); }

/* orb_elem.rs:58-226 */
void oo_ccek1(const double pos[3], const double vel[3], double epoch, oo_elements *out) {
  const double EPS_EQ = 1e-15, EPS_PARAB = 1e-12, EPS_E = 5e-15;
  const double mu = OO_GAUSS_GRAV * OO_GAUSS_GRAV;
  oo_tls_cnt.orbits_built++;
  double h[3];
  oo_cross3(pos, vel, h);
  double h2 = oo_dot3(h, h);
  double hn = sqrt(h2);
  double hu[3] = {h[0] / hn, h[1] / hn, h[2] / hn};
  double sin_i = sqrt(hu[0] * hu[0] + hu[1] * hu[1]);
  double inc = wrap_0_2pi(atan2(sin_i, hu[2]));
  double node;
  if (sin_i <= EPS_EQ) {
    inc = 0.0;
    node = 0.0;
  } else {
    node = wrap_0_2pi(atan2(hu[0], -hu[1]));
  }
  double ri[9], rn[9], rit[9], rnt[9], rot[9];
  oo_rotmt(inc, 0, ri);
  oo_rotmt(node, 2, rn);
  oo_transpose(ri, rit);
  oo_transpose(rn, rnt);
  oo_matmul(rit, rnt, rot);
  double po[3], vo[3];
  oo_matvec(rot, pos, po);
  oo_matvec(rot, vel, vo);
  double rv = po[0] * vo[0] + po[1] * vo[1];
  double rd = sqrt(po[0] * po[0] + po[1] * po[1]);
  double vsq = vo[0] * vo[0] + vo[1] * vo[1];
  double inv_a = 2.0 / rd - vsq / mu;

  out->epoch = epoch;
  int parabolic = 0;
  if (inv_a > EPS_PARAB) {
    double a = 1.0 / inv_a;
    double n = sqrt(mu / ((a * a) * a));
    double esin = rv / (n * a * a);
    double ecos = vsq * rd / mu - 1.0;
    double e = sqrt(esin * esin + ecos * ecos);
    if (fabs(e - 1.0) < EPS_E) {
      parabolic = 1;
    } else {
      double ea = atan2(esin, ecos);
      double ma = wrap_0_2pi(ea - e * sin(ea));
      double x1 = cos(ea) - e;
      double rad = sqrt(1.0 - e * e);
      double x2 = rad * sin(ea);
      double nrm = sqrt(x1 * x1 + x2 * x2);
      double x1n = x1 / nrm, x2n = x2 / nrm;
      double sinw = x1n * po[1] - x2n * po[0];
      double cosw = x1n * po[0] + x2n * po[1];
      double argp = wrap_0_2pi(atan2(sinw, cosw));
      out->kind = OO_ELEM_KEPLERIAN;
      out->e[0] = a; out->e[1] = e; out->e[2] = inc; out->e[3] = node; out->e[4] = argp;
      out->e[5] = ma;
      return;
    }
  } else if (fabs(inv_a) <= EPS_PARAB) {
    parabolic = 1;
  } else {
    double p = h2 / mu;
    double ecosv = p / rd - 1.0;
    double esinv = rv * p / (hn * rd);
    double nu = atan2(esinv, ecosv);
    double e = sqrt(ecosv * ecosv + esinv * esinv);
    if (fabs(e - 1.0) < EPS_E) {
      parabolic = 1;
    } else {
      double q = p / (1.0 + e);
      double argp = wrap_0_2pi(atan2(po[1], po[0]) - nu);
      out->kind = OO_ELEM_COMETARY;
      out->e[0] = q; out->e[1] = e; out->e[2] = inc; out->e[3] = node; out->e[4] = argp;
      out->e[5] = nu;
      return;
    }
  }
  if (parabolic) { /* parabolic_solution closure :122-145 */
    double p = h2 / mu;
    double q = p / 2.0;
    double cosv = p / rd - 1.0;
    double sinv = rv * p / (rd * hn);
    double nu = atan2(sinv, cosv);
    double argp = wrap_0_2pi(atan2(po[1], po[0]) - nu);
    out->kind = OO_ELEM_COMETARY;
    out->e[0] = q; out->e[1] = 1.0; out->e[2] = inc; out->e[3] = node; out->e[4] = argp;
    out->e[5] = nu;
  }
}

/* equinoctial_element.rs:285-313 */
static void from_kepler_internal(double epoch, double a, double e, double inc, double node,
                                 double argp, double ma, oo_elements *out) {
  double dig = node + argp;
  double hh = e * sin(dig);
  double kk = e * cos(dig);
  double th = tan(inc / 2.0);
  double pp = th * sin(node);
  double qq = th * cos(node);
  double lam = oo_rem_euclid(dig + ma, OO_DPI);
  out->kind = OO_ELEM_EQUINOCTIAL;
  out->epoch = epoch;
  out->e[0] = a; out->e[1] = hh; out->e[2] = kk; out->e[3] = pp; out->e[4] = qq; out->e[5] = lam;
}

/* orbit_type/mod.rs:399-443 ; cometary_element.rs:224-290,483-504 */
int oo_to_equinoctial(const oo_elements *in, oo_elements *out) {
  if (in->kind == OO_ELEM_EQUINOCTIAL) { *out = *in; return OO_OK; }
  if (in->kind == OO_ELEM_KEPLERIAN) {
    from_kepler_internal(in->epoch, in->e[0], in->e[1], in->e[2], in->e[3], in->e[4], in->e[5], out);
    return OO_OK;
  }
  /* cometary_to_keplerian :264-290 */
  double e = in->e[1];
  if (fabs(e - 1.0) < 1e-12) return OO_ERR_INVALID_CONVERSION;
  double p = in->e[0] * (1.0 + e);
  double a = -p / (e * e - 1.0);
  /* hyperbolic_mean_anomaly :224-240 */
  if (e <= 1.0) return OO_ERR_INVALID_ORBIT;
  const double EPSC = 1e-15;
  double s = sqrt((e - 1.0) / (e + 1.0));
  double t = tan(0.5 * in->e[5]);
  double x = oo_clamp(s * t, -1.0 + EPSC, 1.0 - EPSC);
  double hh = 2.0 * atanh(x);
  double m = e * sinh(hh) - hh;
  from_kepler_internal(in->epoch, a, e, in->e[2], in->e[3], in->e[4], m, out);
  return OO_OK;
}

/* equinoctial_element.rs:326-348 + roots 0.0.8 find_root_newton_raphson / SimpleConvergency */
int oo_equinoctial_solve_kepler(const oo_elements *eq, double lam1, double lon_peri, double *F) {
  const double eps = OO_EPS * 1e2;
  const int max_iter = 25;
  double h = eq->e[1], k = eq->e[2];
  double x = OO_PI + lon_peri;
  int iter = 0;
  for (;;) {
    oo_tls_cnt.scorer_newton++;
    double f = x - k * sin(x) + h * cos(x) - lam1;
    double d = 1.0 - k * cos(x) - h * sin(x);
    if (fabs(f) < eps) { *F = x; return OO_OK; }
    if (fabs(d) < eps) {
      if (iter == 0) { x = x + 1.0; iter = iter + 1; continue; }
      return OO_ERR_ROOT_FINDING;
    }
    double x1 = x - f / d;
    if (fabs(x - x1) < eps) { *F = x1; return OO_OK; }
    x = x1;
    iter = iter + 1;
    if (iter >= max_iter) return OO_ERR_ROOT_FINDING;
  }
}

/* equinoctial_element.rs:809-867 and :639-759 (compute_derivatives = false) */
int oo_propagate_twobody(const oo_elements *eq, double t0, double t1, double pos[3], double vel[3]) {
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
  /* compute_cartesian_position_and_velocity */
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
  return OO_OK;
}

