// dev_correct.cuh -- the f-g correction phase (P2) as one register-resident loop per candidate.
//
// Reference behaviour restated here:
//   accept_root / positions / Gibbs          gauss.rs:702-870
//   pos_and_vel_correction                   gauss.rs:1284-1418
//   velocity_correction_with_guess           kepler/velocity.rs:94-211
//   UniversalKeplerParams::solve (Newton)    kepler/newton_solver.rs:151-352
//   s_funct                                  kepler/stumpff.rs:78-297
//   eccentricity_control                     orb_elem.rs:257-301
//
// Why it looks like this: ncu (profiles/r01d) showed the first version of this phase bound by
// fixed-latency dependency stalls (37 %) and local-memory round trips (12.7 % of the warp
// instructions were LDL/STL: structs handed by reference to un-inlined functions), with FP64
// instructions only 44 % of the instruction stream.  Here
//   * everything is inlined and every per-candidate quantity lives in registers, except the
//     loop-invariant geometry (R, S, S^-1, tau, a, b: 34 doubles) which is staged in shared memory
//     [slot][thread] (conflict-free 8-byte lanes) and read once per f-g iteration;
//   * divisions that share a denominator use one correctly rounded reciprocal + Markstein's exact
//     correction (q0 = a*y; r = fma(-n, q0, a); q = fma(r, y, q0)), which returns RN(a/n) -- the
//     same bits as the division -- at 3 FP64 instructions per extra numerator; divisions by
//     compile-time constants use the same form with a literal reciprocal;
//   * the Stumpff series computes the (loop-independent) term ratios beta/c_k four terms ahead of
//     the running product, so the dependent chain per term is one DMUL + one DADD;
//   * the work counters are compiled in only for the counting instantiation (COUNT = true).
#pragma once
#include "dev_iod.cuh"

namespace ofb {

#ifndef OUTFIT_CORRECT_THREADS
#define OUTFIT_CORRECT_THREADS 128
#endif
constexpr int kCorrectThreads = OUTFIT_CORRECT_THREADS;
// shared-memory slots per thread (doubles)
enum {
  SL_R0 = 0, SL_R1 = 3, SL_R2 = 6, SL_S0 = 9, SL_S1 = 12, SL_S2 = 15, SL_I0 = 18, SL_I1 = 21, SL_I2 = 24,
  SL_T0 = 27, SL_T1 = 28, SL_T2 = 29,
  // loop state that is idle during the Kepler solves: the outer positions, the first side's result
  SL_P0 = 30, SL_P2 = 33, SL_LF = 36, SL_LG = 37, SL_LCHI = 38, SL_LV = 39,
  SL_COUNT = 42
};
constexpr size_t kCorrectSmemBytes = (size_t)SL_COUNT * kCorrectThreads * sizeof(double);

// RN(a / n) from y = RN(1 / n) (Markstein) for finite operands whose reciprocal and quotient neither
// overflow nor underflow -- AU / day scale quantities here.  For a non-finite or vanished n it yields
// NaN where the division may yield 0 or inf; every use below is followed by the reference's own
// finiteness / acceptability tests, which reject both.
__device__ __forceinline__ double div_mk(double a, double n, double y) {
  const double q0 = __dmul_rn(a, y);
  const double r = __fma_rn(-n, q0, a);
  return __fma_rn(r, y, q0);
}
// Correctly rounded reciprocals of the compile-time denominators, in the constant bank: an FP64
// instruction takes c[bank][offset] as a direct operand, whereas a 64-bit literal costs two UMOVs.
enum { RC_3 = 16, RC_GAUSSK = 17, RC_MU = 18, RC_VLIGHT = 19 };
__constant__ double c_rcp[20] = {
    1.0 / 12.0,  1.0 / 20.0,  1.0 / 30.0,  1.0 / 42.0,  1.0 / 56.0,  1.0 / 72.0,  1.0 / 90.0,  1.0 / 110.0,
    1.0 / 132.0, 1.0 / 156.0, 1.0 / 182.0, 1.0 / 210.0, 1.0 / 240.0, 1.0 / 272.0, 1.0 / 306.0, 1.0 / 342.0,
    1.0 / 3.0,   1.0 / kGaussK, 1.0 / kMu, 1.0 / kVlightAu};

struct WorkC {
  unsigned roots_accepted, fg_iterations, kepler_solves, newton_steps, sfunct_terms;
  unsigned fg_skipped;  // iterations of the reference's loop that the exact early exits did not execute
};

// ---- s_funct, |beta| >= 100 (rare): halving + duplication, as stumpff.rs:200-297 ----------------
__device__ __noinline__ void s_funct_large(double psi, double alpha, double beta, double *out4, unsigned *terms) {
  const double tol = 100.0 * kEps, big = 1.0 / kEps;
  double rp = psi, rb = beta;
  int halvings = 0;
  while (fabs(rb) >= 100.0 && halvings < 30) { rp *= 0.5; rb *= 0.25; ++halvings; }
  double s0 = 1.0, s1 = rp, t0 = 1.0, t1 = rp;
  unsigned n = 0;
  for (int k = 1; k <= 70; ++k) {
    ++n;
    t0 *= rb / ((double)(2 * k - 1) * (double)(2 * k));
    s0 += t0;
    if (fabs(t0) < tol || fabs(t0) > big) break;
  }
  for (int k = 1; k <= 70; ++k) {
    ++n;
    t1 *= rb / ((double)(2 * k) * (double)(2 * k + 1));
    s1 += t1;
    if (fabs(t1) < tol || fabs(t1) > big) break;
  }
  for (int h = 0; h < halvings; ++h) {
    const double c = s0, sn = s1;
    s0 = 2.0 * c * c - 1.0;
    s1 = 2.0 * c * sn;
  }
  out4[0] = s0; out4[1] = s1; out4[2] = (s0 - 1.0) / alpha; out4[3] = (s1 - psi) / alpha;
  *terms += n;
}

// series tail beyond the 8 unrolled terms (it >= 8): table reciprocals up to kSeriesTable, then IEEE
__device__ __noinline__ void s_funct_tail(double beta, double *st /* s2,t2,s3,t3 in/out */, unsigned *terms) {
  const double tol = 100.0 * kEps, big = 1.0 / kEps;
  double s2 = st[0], t2 = st[1], s3 = st[2], t3 = st[3];
  double d = 3.0 + 2.0 * 8;
  unsigned n = 0;
  for (int it = 8; it < 70; ++it) {
    ++n;
    double q2, q3;
    if (it < kSeriesTable) {
      q2 = div_by_const(beta, d * (d + 1.0), c_series_rcp[2 * it]);
      q3 = div_by_const(beta, (d + 1.0) * (d + 2.0), c_series_rcp[2 * it + 1]);
    } else {
      q2 = beta / (d * (d + 1.0));
      q3 = beta / ((d + 1.0) * (d + 2.0));
    }
    t2 *= q2; s2 += t2;
    t3 *= q3; s3 += t3;
    const double a2 = fabs(t2), a3 = fabs(t3);
    if ((a2 < tol && a3 < tol) || a2 > big || a3 > big) break;
    d += 2.0;
  }
  st[0] = s2; st[1] = t2; st[2] = s3; st[3] = t3;
  *terms += n;
}

// denominators of the first 8 series terms: (2k+3)(2k+4) and (2k+4)(2k+5)
#define OFB_SERIES_TERM(C2, C3)                                                     \
  {                                                                                 \
    t2 = t2 * q2_##C2; s2 = s2 + t2;                                                \
    t3 = t3 * q3_##C3; s3 = s3 + t3;                                                \
    if (COUNT) ++terms;                                                             \
    const double a2 = fabs(t2), a3 = fabs(t3);                                      \
    if ((a2 < tol && a3 < tol) || a2 > big || a3 > big) goto series_done;           \
  }
#define OFB_SERIES_Q(K, C2, C3)                                                     \
  double q2_##C2 = div_mk(beta, (double)C2, c_rcp[2 * K]);                          \
  double q3_##C3 = div_mk(beta, (double)C3, c_rcp[2 * K + 1]);

template <bool COUNT>
__device__ __forceinline__ void s_funct_fast(double psi, double alpha, double &s0, double &s1, double &s2o,
                                             double &s3o, unsigned &terms) {
  const double tol = 100.0 * kEps, big = 1.0 / kEps;
  if (psi == 0.0) { s0 = 1.0; s1 = 0.0; s2o = 0.0; s3o = 0.0; return; }
  const double psi2 = psi * psi;
  const double beta = alpha * psi2;
  if (!(fabs(beta) < 100.0)) {
    double o[4];
    unsigned n = 0;
    s_funct_large(psi, alpha, beta, o, &n);
    if (COUNT) terms += n;
    s0 = o[0]; s1 = o[1]; s2o = o[2]; s3o = o[3];
    return;
  }
  double s2 = 0.5 * psi2, t2 = s2;
  double s3 = div_mk(s2 * psi, 3.0, c_rcp[RC_3]), t3 = s3;
  {
    OFB_SERIES_Q(0, 12, 20) OFB_SERIES_Q(1, 30, 42) OFB_SERIES_Q(2, 56, 72) OFB_SERIES_Q(3, 90, 110)
    // scheduling fence: the eight independent ratio chains are issued back to back BEFORE the running
    // products start (the compiler otherwise sinks each ratio into its term, serialising 5 FP64 ops)
    asm volatile("" : "+d"(q2_12), "+d"(q3_20), "+d"(q2_30), "+d"(q3_42), "+d"(q2_56), "+d"(q3_72), "+d"(q2_90), "+d"(q3_110));
    OFB_SERIES_TERM(12, 20) OFB_SERIES_TERM(30, 42) OFB_SERIES_TERM(56, 72) OFB_SERIES_TERM(90, 110)
  }
  {
    OFB_SERIES_Q(4, 132, 156) OFB_SERIES_Q(5, 182, 210) OFB_SERIES_Q(6, 240, 272) OFB_SERIES_Q(7, 306, 342)
    asm volatile("" : "+d"(q2_132), "+d"(q3_156), "+d"(q2_182), "+d"(q3_210), "+d"(q2_240), "+d"(q3_272), "+d"(q2_306), "+d"(q3_342));
    OFB_SERIES_TERM(132, 156) OFB_SERIES_TERM(182, 210) OFB_SERIES_TERM(240, 272) OFB_SERIES_TERM(306, 342)
  }
  {
    double st[4] = {s2, t2, s3, t3};
    unsigned n = 0;
    s_funct_tail(beta, st, &n);
    if (COUNT) terms += n;
    s2 = st[0]; s3 = st[2];
  }
series_done:
  s1 = psi + alpha * s3;
  s0 = 1.0 + alpha * s2;
  s2o = s2;
  s3o = s3;
}
#undef OFB_SERIES_TERM
#undef OFB_SERIES_Q

// cold-start guess (first f-g iteration of a root only): scalars by value, no stack traffic in the caller
__device__ __noinline__ double prelim_kepuni_v(double dt, double r0, double sig0, double alpha, double e0,
                                               double convergency, unsigned max_iter_prelim = 20,
                                               int parabolic_newton = 0) {
  KepIn kp;
  kp.dt = dt; kp.r0 = r0; kp.sig0 = sig0; kp.alpha = alpha; kp.e0 = e0;
  kp.convergency = convergency; kp.max_iter_prelim = max_iter_prelim; kp.parabolic_newton = parabolic_newton;
  return prelim_kepuni(kp);
}

// Newton on the universal Kepler equation (newton_solver.rs:240-352); one s_funct site.
// Returns ok; on success psi and (s2, s3) of the accepted evaluation.
template <bool COUNT>
__device__ __forceinline__ bool kepuni_newton_fast(double dt, double r0, double sig0, double alpha, double convergency,
                                                   double &psi, double &s2, double &s3, WorkC &w, double *s01 = nullptr) {
  const double sdt = kGaussK * dt;
  const double tol = 10.0 * kEps * (1.0 + fabs(sdt));
  bool final_eval = false;
  int it = 0;
  for (;;) {
    if (!final_eval) {
      if (it >= 50) return false;
      ++it;
      if (COUNT) ++w.newton_steps;
      if (!isfinite(psi)) { psi = 0.5; continue; }
    }
    double s0, s1;
    s_funct_fast<COUNT>(psi, alpha, s0, s1, s2, s3, w.sfunct_terms);
    if (s01) { s01[0] = s0; s01[1] = s1; }
    if (final_eval) return true;
    const double res = r0 * s1 + sig0 * s2 + s3 - sdt;
    const double der = r0 * s0 + sig0 * s1 + s2;
    if (fabs(res) <= tol) return true;
    if (!isfinite(der) || fabs(der) < 10.0 * kEps) { psi *= 0.5; continue; }
    const double mx = 2.0 * (1.0 + fabs(psi));
    const double step = clampd(bf_div(-res, der), -mx, mx);
    double cand = psi + step;
    if (cand * psi < 0.0) cand = 0.5 * psi;
    psi = cand;
    const double sa = fabs(step);
    if (sa <= convergency) return true;  // (s2, s3) of the evaluation before the step, like the reference
    if (sa <= convergency * (1.0 + fabs(psi))) final_eval = true;
  }
}

// middle state shared by the two sides of one f-g iteration (eccentricity_control + mid_state)
struct MidC {
  double r2, inv_r2, sig0, hn, ecc, alpha;
  bool defined, accepted;
};
__device__ __forceinline__ MidC middle_state(V3 r, V3 v, double peri_max, double ecc_max) {
  MidC m;
  const double v2 = dot(v, v);
  const double dist = bf_sqrt(dot(r, r));
  const V3 h = cross(r, v);
  const double h2 = dot(h, h);
  m.hn = bf_sqrt(h2);
  m.defined = !(m.hn == 0.0);
  const V3 vxh = cross(v, h);
  const double inv_mu = c_rcp[RC_MU];
  const double inv_d = bf_rcp(dist);
  const V3 lenz = V3{vxh.x * inv_mu - r.x * inv_d, vxh.y * inv_mu - r.y * inv_d, vxh.z * inv_mu - r.z * inv_d};
  m.ecc = bf_sqrt(dot(lenz, lenz));
  const double peri = bf_div(h2, kMu * (1.0 + m.ecc));
  const double energy = v2 / 2.0 - div_mk(kMu, dist, inv_d);
  m.accepted = (m.ecc < ecc_max) && (peri < peri_max);
  m.r2 = dist;
  m.inv_r2 = inv_d;
  m.sig0 = div_mk(dot(r, v), kGaussK, c_rcp[RC_GAUSSK]);
  m.alpha = div_mk(2.0 * energy, kMu, c_rcp[RC_MU]);
  return m;
}

// shared-memory view of this thread's loop-invariant geometry
// (volatile: keeps the compiler from hoisting these loop-invariant loads into registers, which is
// exactly the register pressure -- and the spills -- the staging exists to avoid)
struct GeoSm {
  volatile double *p;  // &smem[threadIdx.x]
  __device__ __forceinline__ double at(int slot) const { return p[slot * kCorrectThreads]; }
  __device__ __forceinline__ V3 v3(int slot) const { return V3{at(slot), at(slot + 1), at(slot + 2)}; }
  __device__ __forceinline__ void put(int slot, double v) const { p[slot * kCorrectThreads] = v; }
  __device__ __forceinline__ void put3(int slot, V3 v) const { put(slot, v.x); put(slot + 1, v.y); put(slot + 2, v.z); }
};

struct SideC {
  bool ok;
  V3 v;
  double f, g, chi;
};
template <bool COUNT>
__device__ __forceinline__ SideC correction_side(const GeoSm &G, int x1_slot, V3 x2, const MidC &m, double dt, bool has_guess,
                                                 double chi_guess, double eps, WorkC &w) {
  SideC o;
  o.ok = false;
  if (COUNT) ++w.kepler_solves;
  double psi = has_guess ? chi_guess : prelim_kepuni_v(dt, m.r2, m.sig0, m.alpha, m.ecc, eps);
  double s2, s3;
  if (!kepuni_newton_fast<COUNT>(dt, m.r2, m.sig0, m.alpha, eps, psi, s2, s3, w)) return o;
  const double f = 1.0 - div_mk(s2, m.r2, m.inv_r2);
  const double g = dt - div_mk(s3, kGaussK, c_rcp[RC_GAUSSK]);
  const double ga = fabs(g);
  if (!isfinite(ga) || ga < 100.0 * kEps * (1.0 + fabs(dt))) return o;
  const V3 x1 = G.v3(x1_slot);
  const double nx = (-f) * x2.x + x1.x, ny = (-f) * x2.y + x1.y, nz = (-f) * x2.z + x1.z;
  const double yg = bf_rcp(g);  // |g| >= 100 eps (1 + |dt|), finite: checked above
  o.v = V3{div_mk(nx, g, yg), div_mk(ny, g, yg), div_mk(nz, g, yg)};
  o.f = f; o.g = g; o.chi = psi;
  o.ok = true;
  return o;
}

// positions_from_c with c1 = -1 (gauss.rs:702): -(x / -1) == x exactly, so rho1 = S^-1 row 1 . gc
__device__ __forceinline__ bool positions_c(const GeoSm &G, double c0, double c2, double min_rho2, V3 &p0, V3 &p1,
                                            V3 &p2, double &epoch) {
  const V3 R0 = G.v3(SL_R0), R1 = G.v3(SL_R1), R2 = G.v3(SL_R2);
  const V3 gc = V3{(R0.x * c0 + R1.x * -1.0) + R2.x * c2, (R0.y * c0 + R1.y * -1.0) + R2.y * c2,
                   (R0.z * c0 + R1.z * -1.0) + R2.z * c2};
  const double rho0 = -bf_div(dot(G.v3(SL_I0), gc), c0);
  const double rho1 = dot(G.v3(SL_I1), gc);
  const double rho2 = -bf_div(dot(G.v3(SL_I2), gc), c2);
  if (rho1 < min_rho2) return false;
  p0 = R0 + rho0 * G.v3(SL_S0);
  p1 = R1 + rho1 * G.v3(SL_S1);
  p2 = R2 + rho2 * G.v3(SL_S2);
  epoch = G.at(SL_T1) - div_mk(rho1, kVlightAu, c_rcp[RC_VLIGHT]);
  return true;
}

// accept_root (gauss.rs:816-870): positions, light-time epoch, Gibbs velocity, acceptability.
// tau / a / b of gauss_prelim (gauss.rs:464-500) are rebuilt from the three epochs (a handful of
// operations per root) instead of occupying six shared-memory slots.  On success the outer positions
// are parked in shared memory (SL_P0, SL_P2); the caller keeps p1, vel, mid in registers.
template <bool COUNT>
__device__ __forceinline__ bool accept_root_fast(const GeoSm &G, const IodDevParams &P, double root, V3 &p1, V3 &vel,
                                                 double &epoch, MidC &mid, WorkC &w) {
  const double t0 = G.at(SL_T0), t1 = G.at(SL_T1), t2 = G.at(SL_T2);
  const double tau1 = kGaussK * (t0 - t1), tau3 = kGaussK * (t2 - t1);
  const double tau13 = tau3 - tau1;
  const double a0 = tau3 / tau13, a2 = -(tau1 / tau13);
  const double b0 = a0 * (tau13 * tau13 - tau3 * tau3) / 6.0;
  const double b2 = a2 * (tau13 * tau13 - tau1 * tau1) / 6.0;
  const double r2m3 = 1.0 / ((root * root) * root);
  V3 p0, p2;
  if (!positions_c(G, a0 + b0 * r2m3, a2 + b2 * r2m3, P.min_rho2_au, p0, p1, p2, epoch)) return false;
  {
    const V3 pos[3] = {p0, p1, p2};
    vel = gibbs_velocity(pos, tau1, tau3);
  }
  mid = middle_state(p1, vel, P.max_perihelion_au, P.max_ecc);
  if (!mid.defined || !mid.accepted) return false;
  if (COUNT) ++w.roots_accepted;
  G.put3(SL_P0, p0);
  G.put3(SL_P2, p2);
  return true;
}

// pos_and_vel_correction (gauss.rs:1284-1418) on an accepted root.  false <=> None (the caller keeps
// the accepted state as a PrelimOrbit); true: (p1, vel, ep) hold the corrected state.
// Register diet: the outer positions and the first side's result are idle while a Kepler solve runs,
// so they live in shared memory; what stays in registers across the solves is p1, vel, the middle
// state and the two chi warm starts.
template <bool COUNT>
__device__ __forceinline__ bool fg_correction_fast(const GeoSm &G, const IodDevParams &P, V3 &p1, V3 &vel, MidC &mid,
                                                   double &ep, WorkC &w) {
  const double dt01 = G.at(SL_T0) - G.at(SL_T1), dt21 = G.at(SL_T2) - G.at(SL_T1);
  if (fabs(dt01) <= kEps || fabs(dt21) <= kEps) return false;
  ep = 0.0;
  bool has_chi = false;
  double chi01 = 0.0, chi21 = 0.0;
#pragma unroll 1
  for (unsigned it = 0; it < P.newton_max_it; ++it) {
    if (COUNT) ++w.fg_iterations;
    // velocity_correction_with_guess guards (velocity.rs:105-123), identical for both sides
    const bool sides_ok = isfinite(mid.hn) && !(mid.hn <= 1e6 * kEps) && mid.defined;
    bool both = false;
    SideC Rr;
    Rr.ok = false;
    if (sides_ok) {
      const SideC L = correction_side<COUNT>(G, SL_P0, p1, mid, G.at(SL_T0) - G.at(SL_T1), has_chi, chi01, P.kepler_eps, w);
      if (L.ok) { G.put(SL_LF, L.f); G.put(SL_LG, L.g); G.put(SL_LCHI, L.chi); G.put3(SL_LV, L.v); }
      const bool l_ok = L.ok;
      Rr = correction_side<COUNT>(G, SL_P2, p1, mid, G.at(SL_T2) - G.at(SL_T1), has_chi, chi21, P.kepler_eps, w);
      both = l_ok && Rr.ok;
    }
    if (!both) {
      // The reference `continue`s here with NOTHING updated (positions, velocity and the chi warm
      // starts are only committed after both sides succeed), so every remaining iteration would
      // repeat this one bit for bit and the loop would end after newton_max_it passes with the
      // current state (gauss.rs:1310-1330).  Leaving now is exact -- and these are the candidates
      // whose every Kepler solve exhausts its 50 Newton steps: left to repeat, a single one of them
      // ran for 15 ms (clock64 probe, profiles/r01n_stragglers.log) and set the duration of the launch.
      if (COUNT) { w.fg_iterations += P.newton_max_it - 1 - it; w.fg_skipped += P.newton_max_it - 1 - it; }
      break;
    }
    const double Lf = G.at(SL_LF), Lg = G.at(SL_LG), Lchi = G.at(SL_LCHI);
    // An iteration that ends in one of the `continue`s below commits only the chi warm starts.  If they
    // come out bit-identical to the ones it started from, the next iteration has exactly the same
    // inputs (positions, velocity, middle state, chi) and therefore the same outcome, and so on until
    // newton_max_it: leaving now is exact.  This is the fate of ~45 % of the accepted roots (the new
    // geocentric distance falls below min_rho2_au on the first or second pass and the warm-started
    // Kepler solves return their guess), which the reference walks through all 50 iterations.
    const bool same_chi = has_chi && Lchi == chi01 && Rr.chi == chi21;
    has_chi = true; chi01 = Lchi; chi21 = Rr.chi;
    const V3 Lv = G.v3(SL_LV);
    const V3 nv = V3{(Lv.x + Rr.v.x) * 0.5, (Lv.y + Rr.v.y) * 0.5, (Lv.z + Rr.v.z) * 0.5};
    const double fl = Lf * Rr.g - Rr.f * Lg;
    bool stall = !isfinite(fl) || fabs(fl) < kEps;
    V3 n0, n1, n2;
    double nep = 0.0;
    if (!stall) {
      const double inv_f = bf_rcp(fl);
      stall = !positions_c(G, Rr.g * inv_f, -Lg * inv_f, P.min_rho2_au, n0, n1, n2, nep);
    }
    MidC nm;
    double denom = 0.0;
    if (!stall) {
      nm = middle_state(n1, nv, P.max_perihelion_au, P.max_ecc);
      if (!nm.defined || !nm.accepted) return false;
      denom = bf_sqrt((dot(n0, n0) + dot(n1, n1)) + dot(n2, n2));
      stall = !isfinite(denom) || denom <= kEps;
    }
    if (stall) {
      if (same_chi) {
        if (COUNT) { w.fg_iterations += P.newton_max_it - 1 - it; w.fg_skipped += P.newton_max_it - 1 - it; }
        break;
      }
      continue;
    }
    const V3 d0 = n0 - G.v3(SL_P0), d1 = n1 - p1, d2 = n2 - G.v3(SL_P2);
    const double rel = bf_div(bf_sqrt((dot(d0, d0) + dot(d1, d1)) + dot(d2, d2)), denom);
    G.put3(SL_P0, n0);
    G.put3(SL_P2, n2);
    p1 = n1;
    vel = nv;
    ep = nep;
    mid = nm;
    if (rel <= P.newton_eps) break;
  }
  return true;
}

}  // namespace ofb
