// dev_lsq.cuh -- differential orbit correction (weighted least-squares Newton-Raphson on equinoctial
// elements with outlier rejection), four lanes per trajectory (k_lsq.cuh).
//
// Reference: differential_orbit_correction/{mod.rs:60-115, diff_cor.rs:282-442,
// single_iteration.rs:140-317, least_square.rs:188-405, outlier_rejection.rs:118-235},
// orbit_type/equinoctial_element.rs:258-270,442-637,639-867, ephemeris/observation_ephemeris.rs:204-258,
// 418-450; the 6x6 Cholesky / Householder-QR inverses follow the published algorithms of nalgebra 0.34
// (column axpy factorisation, column-by-column substitution, sequential dots).
//
// The arithmetic keeps the reference's operation order (no FMA contraction: the file is compiled with
// -fmad=false), so against a CPU restatement only the libm calls differ (<= 1 ulp each).
// What is NOT kept is the reference's data flow: `last_equations` (12 partials per observation, cloned
// every Newton step) is not stored -- the rejection step re-evaluates the partials at the elements the
// last accepted step was linearised at, which reproduces them bit for bit from 7 doubles of state.
#pragma once
#include "dev_elements.cuh"

namespace ofb {

struct LsqCfgDev {
  unsigned long long max_newton_iterations, max_outlier_rejection_passes, max_stagnation_iterations;
  double convergence_threshold, convergence_before_rejection_threshold, rms_stagnation_ratio, rms_divergence_ratio;
  double chi2_reject, chi2_recover;
  double ecc_limit, min_a, max_a, min_q, max_Q;
  int enable_outlier_rejection;
  int free_el[6];
};

#define OFB_M6(m, r, c) ((m)[6 * (c) + (r)])

// The FitLSQ kernels are bound by instruction fetch (one trip of their state machine sweeps most of the kernel):
// the iterative fmod / remainder of the two range reductions, which garbage input alone reaches, stay out of line.
__device__ __noinline__ double lsq_fmod_cold(double x, double m) { return fmod(x, m); }
__device__ __noinline__ double lsq_remainder_cold(double x, double m) { return remainder(x, m); }
__device__ __forceinline__ double lsq_rem_euclid(double x, double m) {  // rem_euclid (dev_kepler.cuh), same values
  if (x >= 0.0 && x < m) return x;
  if (x < 0.0 && x > -m) return x + m;
  const double r = lsq_fmod_cold(x, m);
  return r < 0.0 ? r + m : r;
}

// equinoctial_element.rs:258-270
__device__ __forceinline__ bool lsq_is_bizarre(const double *e, const LsqCfgDev &c) {
  const double ecc = sqrt(e[1] * e[1] + e[2] * e[2]);
  const double peri = e[0] * (1.0 - ecc);
  const double apo = e[0] * (1.0 + ecc);
  return ecc > c.ecc_limit || e[0] < c.min_a || e[0] > c.max_a || peri < c.min_q || apo > c.max_Q;
}

// propagate_twobody(0.0, dt, true) (equinoctial_element.rs:809-867, 639-759) + compute_derivative (:442-637) of the
// orbit `el` = (epoch, a, h, k, p, q, lambda): heliocentric position and velocity (ecliptic J2000) at epoch + t1 and the
// columns col[j] = d pos / d element j (and, with VEL, colv[j] = d vel / d element j).
// false <=> the reference returns Err (e >= 1, Kepler equation not converged).
template <bool VEL>
__device__ __forceinline__ bool lsq_state_and_columns(const double *el, double t1, V3 &pos_out, V3 &vel_out, V3 (&col)[6],
                                                      V3 (&colv)[6]) {
  const double a = el[1], h = el[2], k = el[3], p = el[4], q = el[5];
  const double e2 = h * h + k * k;
  if (sqrt(e2) >= 1.0) return false;  // check_elliptical_orbit
  const double t0 = 0.0;
  const double n = sqrt(kMu / ((a * a) * a));
  double lam1 = el[6] + n * (t1 - t0);
  double lon_peri = 0.0;
  if (e2 > kEps * 1e2) lon_peri = lsq_rem_euclid(atan2_finite(h, k), kTwoPi);
  lam1 = lsq_rem_euclid(lam1, kTwoPi);
  if (lam1 < lon_peri) lam1 += kTwoPi;
  // solve_kepler_equation (:326-348; roots 0.0.8 Newton, eps 100 ulp, 25 iterations)
  const double eps = kEps * 1e2;
  double F = kPi + lon_peri;
  int iter = 0;
  // Every lane that entered together leaves together: a lane whose Newton iteration has ended (state != 0)
  // idles until the slowest of its group is done.  With per-lane `break`s the lanes that left early ran ahead
  // and the 400 flops below executed once per exit time, at 12 of 32 lanes (profiles/r2h_lsq_lines.txt).
  int state = 0;  // 0 iterating | 1 converged | 2 failed (the reference returns Err)
  for (;;) {
    if (state == 0) {
      double sx, cx;
      sincos_angle(F, &sx, &cx);
      const double f = F - k * sx + h * cx - lam1;
      const double d = 1.0 - k * cx - h * sx;
      if (fabs(f) < eps) {
        state = 1;
      } else if (fabs(d) < eps) {
        if (iter == 0) { F = F + 1.0; iter = 1; }
        else state = 2;
      } else {
        const double x1 = F - f / d;
        if (fabs(F - x1) < eps) state = 1;
        else if (++iter >= 25) state = 2;
        F = x1;
      }
    }
    if (!__any_sync(__activemask(), state == 0)) break;
  }
  if (state == 2) return false;
  // compute_cartesian_position_and_velocity (:639-759)
  const double beta = 1.0 / (1.0 + sqrt(1.0 - e2));
  const double bhk = beta * h * k;
  double sF, cF;
  sincos_angle(F, &sF, &cF);
  const double xe = a * ((1.0 - beta * (h * h)) * cF + bhk * sF - k);
  const double ye = a * ((1.0 - beta * (k * k)) * sF + bhk * cF - h);
  const double u = 1.0 + p * p + q * q;
  const double inv_u = 1.0 / u;
  const double common = 2.0 * p * q * inv_u;
  const V3 fv{(1.0 - p * p + q * q) * inv_u, common, -2.0 * p * inv_u};
  const V3 gv{common, (1.0 + p * p - q * q) * inv_u, 2.0 * q * inv_u};
  const V3 pos = xe * fv + ye * gv;
  const double vconst = n * (a * a) / sqrt(xe * xe + ye * ye);
  const double vxe = vconst * (bhk * cF - (1.0 - beta * (h * h)) * sF);
  const double vye = vconst * ((1.0 - beta * (k * k)) * cF - bhk * sF);
  const V3 vel = vxe * fv + vye * gv;
  // compute_derivative (:442-637); the velocity block only where the N-body Jacobian J0 needs it
  const V3 wv{2.0 * p * inv_u, -2.0 * q * inv_u, (1.0 - p * p - q * q) * inv_u};
  const double r = sqrt(xe * xe + ye * ye);
  const double inv_r = 1.0 / r;
  const double inv_1_beta = 1.0 / (1.0 - beta);
  const double b3 = (beta * beta) * beta;
  const double tmp1 = lam1 - F;
  const double tmp2 = beta + (h * h) * b3 * inv_1_beta;
  const double tmp3 = h * k * b3 * inv_1_beta;
  const double tmp4 = beta * h - sF;
  const double tmp5 = beta * k - cF;
  const double tmp6 = beta + (k * k) * b3 * inv_1_beta;
  const double dt = t1 - t0;
  col[0] = V3{(pos.x - 3.0 * vel.x * dt / 2.0) / a, (pos.y - 3.0 * vel.y * dt / 2.0) / a,
              (pos.z - 3.0 * vel.z * dt / 2.0) / a};
  const double dx1de2 = -a * (tmp1 * tmp2 + a * cF * tmp4 * inv_r);
  const double dx2de2 = a * (tmp1 * tmp3 - 1.0 + a * cF * tmp5 * inv_r);
  col[1] = dx1de2 * fv + dx2de2 * gv;
  const double dx1de3 = -a * (tmp1 * tmp3 + 1.0 - a * sF * tmp4 * inv_r);
  const double dx2de3 = a * (tmp1 * tmp6 - a * sF * tmp5 * inv_r);
  col[2] = dx1de3 * fv + dx2de3 * gv;
  {
    const V3 t = q * (ye * fv - xe * gv) - xe * wv;
    col[3] = V3{2.0 * t.x * inv_u, 2.0 * t.y * inv_u, 2.0 * t.z * inv_u};
    const V3 s = p * ((-ye) * fv + xe * gv) + ye * wv;
    col[4] = V3{2.0 * s.x * inv_u, 2.0 * s.y * inv_u, 2.0 * s.z * inv_u};
  }
  col[5] = V3{vel.x / n, vel.y / n, vel.z / n};
  if (VEL) {
    const double tmp7 = 1.0 - r / a;
    const double tmp8 = sF - h;
    const double tmp9 = cF - k;
    const double tmp10 = a * cF * inv_r;
    const double tmp11 = a * sF * inv_r;
    const double tmp12 = n * (a * a) * inv_r;
    const double r3 = (r * r) * r;
    colv[0] = V3{-(vel.x - 3.0 * kMu * pos.x * dt / r3) / (2.0 * a), -(vel.y - 3.0 * kMu * pos.y * dt / r3) / (2.0 * a),
                 -(vel.z - 3.0 * kMu * pos.z * dt / r3) / (2.0 * a)};
    const double ir2 = inv_r * inv_r, a2 = a * a;
    const double dx4de2 = tmp12 * (tmp7 * tmp2 + a2 * tmp8 * tmp4 * ir2 + tmp10 * cF);
    const double dx5de2 = -tmp12 * (tmp7 * tmp3 + a2 * tmp8 * tmp5 * ir2 - tmp10 * sF);
    colv[1] = dx4de2 * fv + dx5de2 * gv;
    const double dx4de3 = tmp12 * (tmp7 * tmp3 + a2 * tmp9 * tmp4 * ir2 - tmp11 * cF);
    const double dx5de3 = -tmp12 * (tmp7 * tmp6 + a2 * tmp9 * tmp5 * ir2 + tmp11 * sF);
    colv[2] = dx4de3 * fv + dx5de3 * gv;
    const V3 t = q * (vye * fv - vxe * gv) - vxe * wv;
    colv[3] = V3{2.0 * t.x * inv_u, 2.0 * t.y * inv_u, 2.0 * t.z * inv_u};
    const V3 sv = p * ((-vye) * fv + vxe * gv) + vye * wv;
    colv[4] = V3{2.0 * sv.x * inv_u, 2.0 * sv.y * inv_u, 2.0 * sv.z * inv_u};
    const double ir3 = (inv_r * inv_r) * inv_r, a3 = (a * a) * a;
    colv[5] = V3{-n * a3 * pos.x * ir3, -n * a3 * pos.y * ir3, -n * a3 * pos.z * ir3};
  }
  pos_out = pos;
  vel_out = vel;
  return true;
}

// topocentric_radec_and_partials (observation_ephemeris.rs:204-258) + element_partials_from_position_partials: predicted
// (ra, dec) of the heliocentric state (ecliptic J2000) seen from obs_equ, and d(ra, dec)/d(elements) through col[j]
__device__ __forceinline__ void lsq_topocentric(V3 pos, V3 vel, const V3 (&col)[6], V3 obs_equ, double &ra, double &dec,
                                                double *d_ra, double *d_dec) {
  const V3 ap = ecl_to_equ(pos), av = ecl_to_equ(vel);
  const V3 rel = ap - obs_equ;
  const double ltt = norm(rel) / kVlightAu;
  const V3 cor = rel - ltt * av;
  const double x = cor.x, y = cor.y, z = cor.z;
  const double rho = norm(cor);
  const double rho_xy = hypot(x, y);
  const double rho_xy_sq = rho_xy * rho_xy;
  dec = atan2_finite(z, rho_xy);
  ra = lsq_rem_euclid(atan2_finite(y, x), kTwoPi);
  const double rho_sq = rho * rho;
  const V3 gra{-y / rho_xy_sq, x / rho_xy_sq, 0.0};
  const V3 gdec{-z * x / (rho_xy * rho_sq), -z * y / (rho_xy * rho_sq), rho_xy / rho_sq};
  const double rel_norm = norm(rel);
  const double aberr = 1.0 / (rel_norm * kVlightAu);
  const double sra = dot(gra, av) * aberr, sdec = dot(gdec, av) * aberr;
  const V3 drp = gra - sra * rel;
  const V3 ddp = gdec - sdec * rel;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const V3 dq = ecl_to_equ(col[j]);
    d_ra[j] = dot(drp, dq);
    d_dec[j] = dot(ddp, dq);
  }
}

// compute_obs_and_partials_2body (observation_ephemeris.rs:418-450): predicted (ra, dec) of the orbit `el` at one
// observation and d(ra, dec)/d(elements).  false <=> the reference returns Err.
__device__ __noinline__ bool lsq_obs_and_partials(const double *el, double t_obs, V3 obs_equ, double &ra,
                                                  double &dec, double *d_ra, double *d_dec) {
  V3 pos, vel, col[6], colv[6];
  if (!lsq_state_and_columns<false>(el, t_obs - el[0], pos, vel, col, colv)) return false;
  lsq_topocentric(pos, vel, col, obs_equ, ra, dec, d_ra, d_dec);
  return true;
}

// ecl_to_equ above is (kCos*y - kSin*z, kSin*y + kCos*z); the reference multiplies by the full 3x3
// matrix, (0*x + c*y) + (-s)*z: the zero products and the sign placement do not change any rounding.

// S = element stride of the 6x6 matrices (1: a private array; kLsqQuads: one column of the quad kernel's shared block);
// vectors (v, out, b) are always private
#define OFB_MS(m, r, c) ((m)[(6 * (c) + (r)) * S])
// nalgebra Cholesky::new: in-place lower factor; false <=> not positive definite
template <int S = 1>
__device__ __noinline__ bool lsq_cholesky6(double *m) {
#pragma unroll 1
  for (int j = 0; j < 6; ++j) {
    for (int k = 0; k < j; ++k) {
      const double factor = -OFB_MS(m, j, k);
      for (int i = j; i < 6; ++i) OFB_MS(m, i, j) = factor * OFB_MS(m, i, k) + OFB_MS(m, i, j);
    }
    const double diag = OFB_MS(m, j, j);
    if (diag == 0.0 || !(diag >= 0.0)) return false;
    const double denom = sqrt(diag);
    OFB_MS(m, j, j) = denom;
    for (int i = j + 1; i < 6; ++i) OFB_MS(m, i, j) = OFB_MS(m, i, j) / denom;
  }
  return true;
}
// Cholesky::inverse: L then L^T substitution on the identity, column by column
template <int S = 1>
__device__ __noinline__ void lsq_cholesky6_inverse_column(const double *l, double *inv, int c) {
  double b[6];
  for (int r = 0; r < 6; ++r) b[r] = r == c ? 1.0 : 0.0;
  for (int i = 0; i < 5; ++i) {
    const double coeff = b[i] / OFB_MS(l, i, i);
    b[i] = coeff;
    for (int r = i + 1; r < 6; ++r) b[r] = -coeff * OFB_MS(l, r, i) + b[r];
  }
  b[5] = b[5] / OFB_MS(l, 5, 5);
  for (int i = 5; i >= 0; --i) {
    double dot = 0.0;
    for (int r = i + 1; r < 6; ++r) dot += OFB_MS(l, r, i) * b[r];
    b[i] = (b[i] - dot) / OFB_MS(l, i, i);
  }
  for (int r = 0; r < 6; ++r) OFB_MS(inv, r, c) = b[r];
}
template <int S = 1>
__device__ __forceinline__ void lsq_cholesky6_inverse(const double *l, double *inv) {
#pragma unroll 1
  for (int c = 0; c < 6; ++c)  // rolled: six copies of the substitution (66 divisions) are 26 KB of code
    lsq_cholesky6_inverse_column<S>(l, inv, c);
}
// nalgebra QR::new + try_inverse (Householder, doubly normalised axis); `m` is destroyed
template <int S = 1>
__device__ __noinline__ bool lsq_qr6_inverse(double *m, double *inv) {
  double diag[6];
#pragma unroll 1
  for (int ic = 0; ic < 6; ++ic) {  // rolled (the matrix is in memory: dynamic indices cost nothing)
    double sq = 0.0;
    for (int r = ic; r < 6; ++r) sq += OFB_MS(m, r, ic) * OFB_MS(m, r, ic);
    const double nrm = sqrt(sq);
    const double x0 = OFB_MS(m, ic, ic);
    const double modulus = x0 >= 0.0 ? x0 : -x0;
    const double sgn = x0 >= 0.0 ? 1.0 : -1.0;
    const double signed_norm = sgn * nrm;
    const double factor = (sq + modulus * nrm) * 2.0;
    OFB_MS(m, ic, ic) = x0 + signed_norm;
    if (factor != 0.0) {
      const double sf = sqrt(factor);
      for (int r = ic; r < 6; ++r) OFB_MS(m, r, ic) = OFB_MS(m, r, ic) / sf;
      double n2 = 0.0;
      for (int r = ic; r < 6; ++r) n2 += OFB_MS(m, r, ic) * OFB_MS(m, r, ic);
      const double nn = sqrt(n2);
      for (int r = ic; r < 6; ++r) OFB_MS(m, r, ic) = OFB_MS(m, r, ic) / nn;
      const double rn = -signed_norm;
      diag[ic] = rn;
      const double sign = signbit(rn) ? -1.0 : 1.0;
      const double m_two = sign * -2.0;
      for (int c = ic + 1; c < 6; ++c) {
        double dot = 0.0;
        for (int r = ic; r < 6; ++r) dot += OFB_MS(m, r, ic) * OFB_MS(m, r, c);
        const double fac = (dot - 0.0) * m_two;
        for (int r = ic; r < 6; ++r) OFB_MS(m, r, c) = fac * OFB_MS(m, r, ic) + sign * OFB_MS(m, r, c);
      }
    } else {
      diag[ic] = signed_norm;
    }
  }
#pragma unroll 1
  for (int c = 0; c < 6; ++c) {
    double b[6];
    for (int r = 0; r < 6; ++r) b[r] = r == c ? 1.0 : 0.0;
    for (int i = 0; i < 6; ++i) {
      const double sign = signbit(diag[i]) ? -1.0 : 1.0;
      double dot = 0.0;
      for (int r = i; r < 6; ++r) dot += OFB_MS(m, r, i) * b[r];
      const double fac = (dot - 0.0) * (sign * -2.0);
      for (int r = i; r < 6; ++r) b[r] = fac * OFB_MS(m, r, i) + sign * b[r];
    }
    for (int i = 5; i >= 0; --i) {
      const double d = fabs(diag[i]);
      if (d == 0.0) return false;
      const double coeff = b[i] / d;
      b[i] = coeff;
      for (int r = 0; r < i; ++r) b[r] = -coeff * OFB_MS(m, r, i) + b[r];
    }
    for (int r = 0; r < 6; ++r) OFB_MS(inv, r, c) = b[r];
  }
  return true;
}
// least_square.rs:329-342 ; `work` is a 36-double temporary
template <int S = 1>
__device__ __forceinline__ bool lsq_invert_normal_matrix(const double *m, double *inv, double *work) {
  for (int i = 0; i < 36; ++i) work[i * S] = m[i * S];
  if (lsq_cholesky6<S>(work)) { lsq_cholesky6_inverse<S>(work, inv); return true; }
  for (int i = 0; i < 36; ++i) work[i * S] = m[i * S];
  if (lsq_qr6_inverse<S>(work, inv)) return true;
  for (int i = 0; i < 36; ++i) inv[i * S] = 0.0;
  return false;
}
template <int S = 1>
__device__ __forceinline__ void lsq_gemv6(const double *m, const double *v, double *out) {
  for (int r = 0; r < 6; ++r) out[r] = OFB_MS(m, r, 0) * v[0];
  for (int c = 1; c < 6; ++c)
    for (int r = 0; r < 6; ++r) out[r] = OFB_MS(m, r, c) * v[c] + out[r];
}
__device__ __forceinline__ double lsq_dot6(const double *a, const double *b) {
  double res = 0.0;
  for (int i = 0; i < 6; ++i) res += a[i] * b[i];
  return res;
}
// least_square.rs:188-199
__device__ __forceinline__ double lsq_angular_diff(double a, double b) {
  double d = a - b;
  // the reference's subtract-until-in-range loop never ends for an infinite difference and takes |d| / 2 pi
  // trips for a huge one (garbage input: predicted and observed RA both lie in [0, 2 pi) otherwise); a kernel
  // must not hang on it
  if (!(fabs(d) < 1e6)) return lsq_remainder_cold(d, kTwoPi);
  while (d > kPi) d -= kTwoPi;
  while (d < -kPi) d += kTwoPi;
  return d;
}

}  // namespace ofb
