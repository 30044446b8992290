// dev_elements.cuh -- state vector -> orbital elements, element conversions and the RMS scorer's
// per-observation step (equinoctial two-body propagation + topocentric RA/Dec residual).
//
// Reference behaviour:
//   ccek1                      src/orb_elem.rs:58-226
//   Keplerian/Cometary -> Equ  src/orbit_type/mod.rs:399-443, equinoctial_element.rs:285-313,
//                              cometary_element.rs:224-290
//   propagate_twobody          src/orbit_type/equinoctial_element.rs:326-348, 639-867
//   ephemeris_error            src/ephemeris/observation_ephemeris.rs:222-416, aberration.rs:139
#pragma once
#include "dev_gauss.cuh"

namespace ofb {

// (r, v) ecliptic J2000 -> Keplerian or Cometary elements
__device__ __noinline__ void ccek1(V3 r, V3 v, double epoch, Orbit &o) {
  const V3 h = cross(r, v);
  const double h2 = dot(h, h);
  const double hn = sqrt(h2);
  const V3 hu = V3{h.x / hn, h.y / hn, h.z / hn};
  const double sin_i = sqrt(hu.x * hu.x + hu.y * hu.y);
  double inc = rem_euclid(atan2(sin_i, hu.z), kTwoPi);
  double node;
  if (sin_i <= 1e-15) { inc = 0.0; node = 0.0; }
  else node = rem_euclid(atan2(hu.x, -hu.y), kTwoPi);
  double si, ci, sn, cn;
  sincos(inc, &si, &ci);
  sincos(node, &sn, &cn);
  // rows of R_x(inc)^T R_z(node)^T : orbital-plane frame
  const V3 r0 = V3{cn, sn, 0.0}, r1 = V3{ci * (-sn), ci * cn, si}, r2 = V3{(-si) * (-sn), (-si) * cn, ci};
  const V3 po = V3{dot(r0, r), dot(r1, r), dot(r2, r)};
  const V3 vo = V3{dot(r0, v), dot(r1, v), dot(r2, v)};
  const double rv = po.x * vo.x + po.y * vo.y;
  const double rd = sqrt(po.x * po.x + po.y * po.y);
  const double vsq = vo.x * vo.x + vo.y * vo.y;
  const double inv_a = 2.0 / rd - vsq / kMu;
  o.epoch = epoch;
  bool parabolic = false;
  if (inv_a > 1e-12) {
    const double a = 1.0 / inv_a;
    const double n = sqrt(kMu / ((a * a) * a));
    const double esin = rv / (n * a * a);
    const double ecos = vsq * rd / kMu - 1.0;
    const double e = sqrt(esin * esin + ecos * ecos);
    if (fabs(e - 1.0) < 5e-15) {
      parabolic = true;
    } else {
      const double ea = atan2(esin, ecos);
      double sE, cE;
      sincos(ea, &sE, &cE);
      const double ma = rem_euclid(ea - e * sE, kTwoPi);
      const double x1 = cE - e;
      const double x2 = sqrt(1.0 - e * e) * sE;
      const double nrm = sqrt(x1 * x1 + x2 * x2);
      const double x1n = x1 / nrm, x2n = x2 / nrm;
      const double argp = rem_euclid(atan2(x1n * po.y - x2n * po.x, x1n * po.x + x2n * po.y), kTwoPi);
      o.kind = 0;
      o.e[0] = a; o.e[1] = e; o.e[2] = inc; o.e[3] = node; o.e[4] = argp; o.e[5] = ma;
      return;
    }
  } else if (fabs(inv_a) <= 1e-12) {
    parabolic = true;
  } else {
    const double p = h2 / kMu;
    const double ecv = p / rd - 1.0;
    const double esv = rv * p / (hn * rd);
    const double nu = atan2(esv, ecv);
    const double e = sqrt(ecv * ecv + esv * esv);
    if (fabs(e - 1.0) < 5e-15) {
      parabolic = true;
    } else {
      o.kind = 2;
      o.e[0] = p / (1.0 + e); o.e[1] = e; o.e[2] = inc; o.e[3] = node;
      o.e[4] = rem_euclid(atan2(po.y, po.x) - nu, kTwoPi);
      o.e[5] = nu;
      return;
    }
  }
  if (parabolic) {
    const double p = h2 / kMu;
    const double nu = atan2(rv * p / (rd * hn), p / rd - 1.0);
    o.kind = 2;
    o.e[0] = p / 2.0; o.e[1] = 1.0; o.e[2] = inc; o.e[3] = node;
    o.e[4] = rem_euclid(atan2(po.y, po.x) - nu, kTwoPi);
    o.e[5] = nu;
  }
}

struct Equinoctial {
  double epoch, a, h, k, p, q, lambda;
};
// returns 0 or OUTFIT_ST_INVALID_CONVERSION (9) / OUTFIT_ST_INVALID_ORBIT (10)
__device__ __noinline__ int to_equinoctial(const Orbit &o, Equinoctial &q) {
  double a = o.e[0], e = o.e[1], m = o.e[5];
  if (o.kind != 0) {
    if (fabs(e - 1.0) < 1e-12) return 9;
    const double p = o.e[0] * (1.0 + e);
    a = -p / (e * e - 1.0);
    if (e <= 1.0) return 10;
    const double s = sqrt((e - 1.0) / (e + 1.0));
    const double x = clampd(s * tan(0.5 * o.e[5]), -1.0 + 1e-15, 1.0 - 1e-15);
    const double hh = 2.0 * atanh(x);
    m = e * sinh(hh) - hh;
  }
  const double dig = o.e[3] + o.e[4];
  double sd, cd, sO, cO;
  sincos(dig, &sd, &cd);
  sincos(o.e[3], &sO, &cO);
  const double th = tan(o.e[2] / 2.0);
  q.epoch = o.epoch;
  q.a = a;
  q.h = e * sd;
  q.k = e * cd;
  q.p = th * sO;
  q.q = th * cO;
  q.lambda = rem_euclid(dig + m, kTwoPi);
  return 0;
}

// Everything of the two-body propagation that does not depend on the observation epoch
// (hoisted out of the per-observation loop; the reference recomputes it per call).
struct ScoreOrbit {
  double epoch, n, lambda, lon_peri, h, k, a;
  double ch, ck, bhk;  // 1 - beta h^2, 1 - beta k^2, beta h k
  V3 F, G;             // equinoctial frame vectors rotated to equatorial J2000
  bool elliptic;
};
__device__ __forceinline__ ScoreOrbit make_score_orbit(const Equinoctial &q) {
  ScoreOrbit s;
  const double e2 = q.h * q.h + q.k * q.k;
  s.elliptic = !(sqrt(e2) >= 1.0);
  s.epoch = q.epoch; s.a = q.a; s.h = q.h; s.k = q.k; s.lambda = q.lambda;
  s.n = sqrt(kMu / ((q.a * q.a) * q.a));
  s.lon_peri = (e2 > kEps * 1e2) ? rem_euclid(atan2(q.h, q.k), kTwoPi) : 0.0;
  const double beta = 1.0 / (1.0 + sqrt(1.0 - e2));
  s.bhk = beta * q.h * q.k;
  s.ch = 1.0 - beta * (q.h * q.h);
  s.ck = 1.0 - beta * (q.k * q.k);
  const double u = 1.0 + q.p * q.p + q.q * q.q;
  const double inv_u = 1.0 / u;
  const double common = 2.0 * q.p * q.q * inv_u;
  s.F = ecl_to_equ(V3{(1.0 - q.p * q.p + q.q * q.q) * inv_u, common, -2.0 * q.p * inv_u});
  s.G = ecl_to_equ(V3{common, (1.0 + q.p * q.p - q.q * q.q) * inv_u, 2.0 * q.q * inv_u});
  return s;
}

// Normalised squared residual of one observation; false <=> the reference returns Err.
// COUNT = false compiles the work counters out (w is then untouched).
template <bool COUNT>
__device__ __forceinline__ bool ephemeris_error(const ScoreOrbit &s, double t_obs, double ra_obs,
                                                double dec_obs, double cos_dec_obs, double sig_ra,
                                                double sig_dec, V3 obs_equ, double &chi2, Work &w) {
  if (COUNT) ++w.scorer_evals;
  double lam1 = rem_euclid(s.lambda + s.n * ((t_obs - s.epoch) - 0.0), kTwoPi);
  if (lam1 < s.lon_peri) lam1 += kTwoPi;
  // generalised Kepler equation F - k sin F + h cos F = lambda  (roots 0.0.8 Newton, eps 100 ulp, 25 its)
  const double eps = kEps * 1e2;
  double x = kPi + s.lon_peri;
  double sF, cF;
  int iter = 0;
  bool last = false;
  for (;;) {
    sincos_angle(x, &sF, &cF);  // the only sincos site: also evaluates at the accepted root
    if (last) break;
    if (COUNT) ++w.scorer_newton;
    const double f = x - s.k * sF + s.h * cF - lam1;
    const double d = 1.0 - s.k * cF - s.h * sF;
    if (fabs(f) < eps) break;
    if (fabs(d) < eps) {
      if (iter == 0) { x = x + 1.0; iter = 1; continue; }
      return false;
    }
    const double x1 = x - bf_div(f, d);
    const bool conv = fabs(x - x1) < eps;
    x = x1;
    if (conv) { last = true; continue; }
    if (++iter >= 25) return false;
  }
  const double xe = s.a * (s.ch * cF + s.bhk * sF - s.k);
  const double ye = s.a * (s.ck * sF + s.bhk * cF - s.h);
  const double vc = bf_div(s.n * (s.a * s.a), bf_sqrt(xe * xe + ye * ye));
  const double vxe = vc * (s.bhk * cF - s.ch * sF);
  const double vye = vc * (s.ck * cF - s.bhk * sF);
  const V3 pos = xe * s.F + ye * s.G;
  const V3 vel = vxe * s.F + vye * s.G;
  const V3 rel = pos - obs_equ;
  const double ltt = div_by_const(bf_sqrt(dot(rel, rel)), kVlightAu, 1.0 / kVlightAu);  // RN(x / c), Markstein
  const V3 cor = rel - ltt * vel;
  const double dec = atan2_finite(cor.z, hypot(cor.x, cor.y));
  const double ra = rem_euclid(atan2_finite(cor.y, cor.x), kTwoPi);
  double da = ra_obs - ra;
  if (!(fabs(da) < kTwoPi)) da = fmod(da, kTwoPi);  // |x| < m: fmod(x, m) == x exactly
  if (da > kPi) da -= kTwoPi;  // reference quirk: wraps only the > pi side
  const double a = cos_dec_obs * bf_div(da, sig_ra);
  const double b = bf_div(dec_obs - dec, sig_dec);
  chi2 = a * a + b * b;
  return true;
}

}  // namespace ofb
