// dev_gauss.cuh -- device-side Gauss preliminary orbit for ONE observation triplet (one lane).
//
// Reference behaviour: /root/reference/src/initial_orbit_determination/gauss.rs
//   gauss_prelim :532  coeff_eight_poly :585  visit_real_positive_roots/aberth :964
//   accept_root :816   position_vector_and_reference_epoch :702  gibbs_correction :754
//   pos_and_vel_correction :1284   prelim_orbit_all :1119   prelim_orbit :1238
//
// Aberth-Ehrlich root finder: the reference's root ORDER decides which orbit a triplet
// contributes, and the order of the two roots that come from one conjugate pair of starting
// points is decided by rounding-level symmetry breaking.  The sweep below therefore reproduces
// the arithmetic of the aberth 0.4.1 crate exactly (explicit IEEE intrinsics, fused Horner steps,
// textbook complex division, index-ordered sums) and takes the unit-circle starting directions
// from the host libm (c_aberth_dir), so the 8 roots are bit-identical to the CPU path.
#pragma once
#include "dev_kepler.cuh"

namespace ofb {

struct IodDevParams {
  double noise_scale, extf, dtmax, dt_min, dt_max_triplet, inv_optimal_interval;
  double max_ecc, max_perihelion_au, min_rho2_au;
  double aberth_eps, kepler_eps, r2_min_au, r2_max_au, newton_eps, root_imag_eps;
  unsigned n_noise, max_triplets, max_obs_for_triplets, aberth_max_iter, max_tested_solutions,
      newton_max_it;
};

// cos/sin of theta_k = (2 pi / 8) k + (pi / 2) / 8, k = 0..7, evaluated by the HOST libm.
__constant__ double c_aberth_dir[16];

struct Cx {
  double re, im;
};
// exact (never contracted) complex helpers mirroring num-complex 0.4
__device__ __forceinline__ Cx cx_sub(Cx a, Cx b) { return Cx{__dsub_rn(a.re, b.re), __dsub_rn(a.im, b.im)}; }
__device__ __forceinline__ Cx cx_add(Cx a, Cx b) { return Cx{__dadd_rn(a.re, b.re), __dadd_rn(a.im, b.im)}; }
__device__ __forceinline__ Cx cx_mul(Cx a, Cx b) {
  return Cx{__dsub_rn(__dmul_rn(a.re, b.re), __dmul_rn(a.im, b.im)),
            __dadd_rn(__dmul_rn(a.re, b.im), __dmul_rn(a.im, b.re))};
}
// Two quotients over one denominator n: one correctly rounded reciprocal y = RN(1/n) and Markstein's
// correction (q0 = a*y; r = fma(-n, q0, a); q = fma(r, y, q0)) give RN(a/n), the bits of the IEEE
// division, whenever 1/n and the quotient neither overflow nor underflow -- |z_i - z_k|^2 and
// |p s - p'|^2 of O(1) iterates are nowhere near that.  A vanished denominator (coincident iterates)
// yields NaN here where the division yields +-inf or NaN: both end the solve as `Failed`
// (the caller tests isfinite on the new iterate), so no guard is needed.
__device__ __forceinline__ Cx div2_same_denominator(double a, double b, double n) {
  const double y = bf_rcp(n);
  const double qa = __dmul_rn(a, y), qb = __dmul_rn(b, y);
  const double ra = __fma_rn(-n, qa, a), rb = __fma_rn(-n, qb, b);
  return Cx{__fma_rn(ra, y, qa), __fma_rn(rb, y, qb)};
}
__device__ __forceinline__ Cx cx_div(Cx a, Cx b) {
  const double n = __dadd_rn(__dmul_rn(b.re, b.re), __dmul_rn(b.im, b.im));
  const double re = __dadd_rn(__dmul_rn(a.re, b.re), __dmul_rn(a.im, b.im));
  const double im = __dsub_rn(__dmul_rn(a.im, b.re), __dmul_rn(a.re, b.im));
  return div2_same_denominator(re, im, n);
}
__device__ __forceinline__ Cx cx_recip(Cx b) {  // (1 + 0i) / b
  const double n = __dadd_rn(__dmul_rn(b.re, b.re), __dmul_rn(b.im, b.im));
  return div2_same_denominator(b.re, -b.im, n);
}
// one Horner step r*x + c with the fused form of num-complex's MulAdd
__device__ __forceinline__ Cx cx_horner_step(Cx r, Cx x, double c) {
  Cx o;
  o.re = __dsub_rn(__fma_rn(r.re, x.re, c), __dmul_rn(r.im, x.im));
  o.im = __fma_rn(r.re, x.im, __fma_rn(r.im, x.re, 0.0));
  return o;
}

// p(z) = z^8 + c6 z^6 + c3 z^3 + c0 and p'(z), dense Horner (zeros included: they round)
__device__ __forceinline__ void poly8_eval(Cx z, double c0, double c3, double c6, Cx &p, Cx &dp) {
  // The first two dense Horner steps are exact: (0*z + 1) = 1 and (1*z + 0) = z (resp. 8 and 8z for
  // the derivative; scaling by 8 is exact), so the recurrences start from z and 8z.  (Only the sign of
  // a zero component can differ, which no later operation observes; a non-finite z stays non-finite.)
  Cx r = z;
  r = cx_horner_step(r, z, c6);
  r = cx_horner_step(r, z, 0.0);
  r = cx_horner_step(r, z, 0.0);
  r = cx_horner_step(r, z, c3);
  r = cx_horner_step(r, z, 0.0);
  r = cx_horner_step(r, z, 0.0);
  r = cx_horner_step(r, z, c0);
  p = r;
  Cx d = Cx{__dmul_rn(8.0, z.re), __dmul_rn(8.0, z.im)};
  d = cx_horner_step(d, z, __dmul_rn(c6, 6.0));
  d = cx_horner_step(d, z, 0.0);
  d = cx_horner_step(d, z, 0.0);
  d = cx_horner_step(d, z, __dmul_rn(c3, 3.0));
  d = cx_horner_step(d, z, 0.0);
  d = cx_horner_step(d, z, 0.0);
  dp = d;
}

// returns 0 converged / 1 max-iter / 2 failed; the 8 iterates (index order k) are left in zsm.
// One Jacobi sweep = (1) the 28 pair reciprocals 1/(z_i - z_k), i < k, each added to sum_i and
// subtracted from sum_k -- IEEE negation is exact, and visiting the pairs in lexicographic order
// delivers every sum its terms in increasing index order, so the sums are bit-identical to the
// crate's 56-reciprocal double loop at half the divisions; (2) per root: dense Horner p, p',
// w = p / (p * sum - p'), z += w.  Fully unrolled on registers (this kernel's code is small).
__device__ __forceinline__ int aberth8(double c0, double c3, double c6, unsigned max_iter, double eps,
                                       volatile double *zsm, unsigned stride, Work &w) {
  // Cauchy-type start radius: smallest integer r0 with S(r0) > 0, S(w) = w^8 - |c6| w^6 - |c3| w^3 - |c0|
  const double s0 = -fabs(c0), s3 = -fabs(c3), s6 = -fabs(c6);
  // The crate walks r0 = 1, 2, 3, ... (at most 100000 steps).  S is negative below its single positive
  // root and positive above it, and rounding can only blur the sign within ~1e-12 relative of that
  // root, far less than the unit spacing of the candidates: the first integer with S > 0 is found by
  // doubling + bisection in <= 34 evaluations instead of up to 1e5 (r0 reaches 1e4..1e5 on a few
  // candidates per 100k trajectories; walked one by one they were the long pole of this kernel).
  auto s_pos = [&](double rr) {
    double r = 0.0;
    r = __dsub_rn(__fma_rn(r, rr, 1.0), 0.0);
    r = __fma_rn(r, rr, -0.0);
    r = __fma_rn(r, rr, s6);
    r = __fma_rn(r, rr, -0.0);
    r = __fma_rn(r, rr, -0.0);
    r = __fma_rn(r, rr, s3);
    r = __fma_rn(r, rr, -0.0);
    r = __fma_rn(r, rr, -0.0);
    r = __fma_rn(r, rr, s0);
    return r > 0.0;
  };
  double r0 = 1.0;
  if (!s_pos(1.0)) {
    double lo = 1.0, hi = 2.0;  // S(lo) <= 0
    while (hi < 100000.0 && !s_pos(hi)) { lo = hi; hi = hi * 2.0; }
    if (hi >= 100000.0) {
      hi = 100000.0;
      if (!s_pos(hi)) { lo = hi; hi = 100001.0; }  // the crate's walk ends at 100001 without a sign change
    }
    while (hi - lo > 1.0) {
      const double mid = floor(0.5 * (lo + hi));
      if (s_pos(mid)) hi = mid; else lo = mid;
    }
    r0 = hi;
  }
  // zsm: this thread's column of a [32][blockDim.x] shared array: z re (0-7), z im (8-15),
  // sum re (16-23), sum im (24-31).  The 28-pair section is unrolled on registers; the per-root
  // update is a ROLLED loop over shared memory: fully unrolled, one sweep was ~40 KB of SASS, more
  // than the 32 KB instruction cache level, and 30 % of the issue slots stalled on instruction fetch
  // (ncu r01q: stall_no_inst).
#pragma unroll 1
  for (int k = 0; k < 8; ++k) {
    zsm[k * stride] = __dadd_rn(-0.0, __dmul_rn(r0, c_aberth_dir[2 * k]));
    zsm[(8 + k) * stride] = __dmul_rn(r0, c_aberth_dir[2 * k + 1]);
  }
#pragma unroll 1
  for (unsigned it = 0; it < max_iter; ++it) {
    ++w.aberth_sweeps;
    {
      double zr[8], zi[8], sr[8], si[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { zr[k] = zsm[k * stride]; zi[k] = zsm[(8 + k) * stride]; sr[k] = 0.0; si[k] = 0.0; }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int k = i + 1; k < 8; ++k) {
          const Cx rec = cx_recip(Cx{__dsub_rn(zr[i], zr[k]), __dsub_rn(zi[i], zi[k])});
          sr[i] = __dadd_rn(sr[i], rec.re);
          si[i] = __dadd_rn(si[i], rec.im);
          sr[k] = __dadd_rn(sr[k], -rec.re);  // 1/(z_k - z_i) = -(1/(z_i - z_k)) exactly
          si[k] = __dadd_rn(si[k], -rec.im);
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { zsm[(16 + k) * stride] = sr[k]; zsm[(24 + k) * stride] = si[k]; }
    }
    bool converged = true;
    bool failed = false;
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
      const Cx z = Cx{zsm[i * stride], zsm[(8 + i) * stride]};
      const Cx sum = Cx{zsm[(16 + i) * stride], zsm[(24 + i) * stride]};
      Cx p, dp;
      poly8_eval(z, c0, c3, c6, p, dp);
      const Cx nz = cx_add(z, cx_div(p, cx_sub(cx_mul(p, sum), dp)));
      if (!(isfinite(nz.re) && isfinite(nz.im))) failed = true;
      if (!(fabs(__dsub_rn(nz.re, z.re)) < eps && fabs(__dsub_rn(nz.im, z.im)) < eps)) converged = false;
      zsm[i * stride] = nz.re;  // in place: no later root of this sweep reads z_i (the sums are complete)
      zsm[(8 + i) * stride] = nz.im;
    }
    if (failed) return 2;  // the caller maps it to PolynomialRootFindingFailed
    if (converged) return 0;
  }
  return 1;
}

struct Triplet {
  double ra[3], dec[3], t[3];
  V3 R[3];  // heliocentric observer positions at the 3 epochs (equatorial J2000), AU
};
struct Orbit {
  int kind;        // 0 Keplerian, 2 Cometary
  int corrected;   // 1 CorrectedOrbit / 0 PrelimOrbit
  double epoch;
  double e[6];
};

struct GaussGeom {
  V3 S[3];      // unit line-of-sight vectors (columns of the unit matrix)
  V3 SiR[3];    // ROWS of the inverse
  double tau1, tau3;
  double a0, a2, b0, b2;
};

__device__ __forceinline__ V3 gibbs_velocity(const V3 (&pos)[3], double tau1, double tau3) {
  const double tau13 = tau3 - tau1;
  const double n1 = norm(pos[0]), n2 = norm(pos[1]), n3 = norm(pos[2]);
  const double r1m3 = 1.0 / ((n1 * n1) * n1), r2m3 = 1.0 / ((n2 * n2) * n2), r3m3 = 1.0 / ((n3 * n3) * n3);
  const double d1 = tau3 * (r1m3 / 12.0 - 1.0 / (tau1 * tau13));
  const double d2 = (tau1 + tau3) * (r2m3 / 12.0 - 1.0 / (tau1 * tau3));
  const double d3 = -tau1 * (r3m3 / 12.0 + 1.0 / (tau3 * tau13));
  const double e1 = -d1;
  return V3{kGaussK * ((pos[0].x * e1 + pos[1].x * d2) + pos[2].x * d3),
            kGaussK * ((pos[0].y * e1 + pos[1].y * d2) + pos[2].y * d3),
            kGaussK * ((pos[0].z * e1 + pos[1].z * d2) + pos[2].z * d3)};
}

}  // namespace ofb
