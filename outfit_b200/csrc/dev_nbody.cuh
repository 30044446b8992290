// dev_nbody.cuh -- N-body propagation with frozen perturbers.
//
// Reference behaviour (paths under /root/reference/src):
//   EquinoctialElements::propagate_nbody   orbit_type/equinoctial_element.rs:908-968
//   NBodyOde::diff + helpers               propagator/nbody.rs:127-356
//   integrate_augmented_state              propagator/nbody.rs:505-523 (DOP853, NBodyConfig abs_tol / rel_tol)
// The reference integrates with the un-vendored crate `differential_equations`; the DOP853 here is the published
// method (coefficients and step-size controller as in scipy.integrate.DOP853, the implementation the CPU checker of
// tests/ is pinned against): PARITY WITH THE CRATE IS UNPINNED, results agree at the tolerance level.
//
// Mapping: EIGHT LANES PER ORBIT.  The augmented state is [r, v, Phi] (42 doubles) and the variational equations
// dPhi/dt = A(r) Phi couple the columns of Phi only through r, so lane 0 of a group integrates (r, v), lanes 1..6 one
// column of Phi each (6 doubles per lane), lane 7 idles.  Per stage lane 0 broadcasts the stage position (3 shuffles),
// every lane rebuilds the 3x3 gravity gradient from the frozen perturbers, and the error norm of the step-size
// controller is reduced over the group (3 xor-shuffles), so all lanes of a group take the same steps.  The stage
// derivatives (10 live slots x 6 doubles per lane) live in shared memory [slot][thread].  The four groups of a warp
// run a warp-uniform number of step attempts (finished groups ride along with h = 0).
#pragma once
#include "dev_elements.cuh"

namespace ofb {

#include "dop853_coeffs.inc"

#ifndef OUTFIT_NB_THREADS
#define OUTFIT_NB_THREADS 64
#endif
#ifndef OUTFIT_NB_BPS
#define OUTFIT_NB_BPS 7
#endif
constexpr int kNbThreads = OUTFIT_NB_THREADS;  // 8 orbits per block
// Stage derivatives per lane.  Of the 12 stages k2 feeds only stage 3 and k3 only stages 4 and 5 (the zeros of
// Hairer's a-matrix), so k11 and k12 take their places: 10 live slots instead of 12 -- with 64-thread blocks 30 KB
// per block, 7 blocks (14 warps) per SM where 13 slots x 128 threads allowed 2 blocks (8 warps) of a kernel that
// waits on FP64 latency (warps_active 9 %, profiles/r2h_ncu_bulk.txt).
constexpr int kNbStageSlots = 10;
constexpr int kNbSlots = kNbStageSlots * 6;
__device__ __forceinline__ int nb_stage_slot(int s) { return s == 0 ? 0 : (s == 1 ? 9 : (s == 2 ? 8 : s - 2)); }
constexpr size_t kNbSmemBytes = (size_t)kNbSlots * kNbThreads * sizeof(double);
constexpr int kNbMaxPert = 12;

struct NbPert {                              // PerturberSnapshot (nbody.rs:17-32), per orbit
  double gm[kNbMaxPert];
  V3 pos[kNbMaxPert];
  V3 aind[kNbMaxPert];                       // nb_prepare: the indirect acceleration of each perturber
  int n;
};
// the indirect acceleration of perturber p (the Sun's own fall towards it, nbody.rs:150-175): a constant of the orbit
__device__ __forceinline__ V3 nb_indirect(const NbPert &P, int p) {
  const double pd = sqrt(dot(P.pos[p], P.pos[p]));
  if (!(pd > 1e-10)) return V3{0.0, 0.0, 0.0};
  const double c = P.gm[p] / (pd * pd * pd);
  return V3{c * P.pos[p].x, c * P.pos[p].y, c * P.pos[p].z};
}
// once per orbit, after gm / pos are filled: the right-hand side then reads the constant instead of recomputing a
// square root and a division per perturber and evaluation (same values)
__device__ __forceinline__ void nb_prepare(NbPert &P) {
  for (int p = 0; p < kNbMaxPert; ++p) P.aind[p] = p < P.n ? nb_indirect(P, p) : V3{0.0, 0.0, 0.0};
}

// acceleration (lane 0 only needs it) and gravity gradient at heliocentric position r (nbody.rs:127-270)
__device__ __forceinline__ void nb_field(const NbPert &P, V3 r, V3 &acc, double (&G)[9]) {
  acc = V3{0.0, 0.0, 0.0};
#pragma unroll
  for (int q = 0; q < 9; ++q) G[q] = 0.0;
  // the perturber of the NEXT trip is fetched (local memory: dynamic index) before the arithmetic of this one
  double gm_n = P.gm[0];
  V3 pos_n = P.pos[0], aind_n = P.aind[0];
#pragma unroll 1
  for (int p = 0; p < P.n; ++p) {
    const double gm = gm_n;
    const V3 pos = pos_n, aind = aind_n;
    if (p + 1 < P.n) { gm_n = P.gm[p + 1]; pos_n = P.pos[p + 1]; aind_n = P.aind[p + 1]; }
    const V3 d = r - pos;
    const double dist = bf_sqrt(dot(d, d));  // branch-free forms (dev_kepler.cuh): the same bits for normal operands
    const double dist3 = dist * dist * dist;
    const double cdir = bf_div(-gm, dist3);
    acc = V3{acc.x + cdir * d.x + aind.x, acc.y + cdir * d.y + aind.y, acc.z + cdir * d.z + aind.z};
    const double dist5 = dist * dist * dist * dist * dist;
    const double a = bf_rcp(dist3), b = bf_div(3.0, dist5);
    const double dv[3] = {d.x, d.y, d.z};
    // the gradient is symmetric to the bit (dv[r] * dv[c] == dv[c] * dv[r], the same formula either side of the
    // diagonal): six entries per perturber, mirrored after the loop
#pragma unroll
    for (int rr = 0; rr < 3; ++rr)
#pragma unroll
      for (int cc = rr; cc < 3; ++cc) G[3 * rr + cc] = G[3 * rr + cc] + (-gm) * ((rr == cc ? 1.0 : 0.0) * a - (dv[rr] * dv[cc]) * b);
  }
  G[3] = G[1]; G[6] = G[2]; G[7] = G[5];
}

// derivative of this lane's 6 components at stage state w (its own 6 values) and stage position rs (from lane 0)
__device__ __forceinline__ void nb_rhs_lane(const NbPert &P, int role, const double (&w)[6], V3 rs, double (&dw)[6]) {
  V3 acc;
  double G[9];
  nb_field(P, rs, acc, G);
  dw[0] = w[3]; dw[1] = w[4]; dw[2] = w[5];
  if (role == 0) {
    dw[3] = acc.x; dw[4] = acc.y; dw[5] = acc.z;
  } else {
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) dw[3 + rr] = (G[3 * rr + 0] * w[0] + G[3 * rr + 1] * w[1]) + G[3 * rr + 2] * w[2];
  }
}

__device__ __forceinline__ double nb_group_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ V3 nb_group_pos(const double (&w)[6], unsigned lane) {
  const int src = (int)(lane & ~7u);
  return V3{__shfl_sync(0xffffffffu, w[0], src), __shfl_sync(0xffffffffu, w[1], src), __shfl_sync(0xffffffffu, w[2], src)};
}

// state of an equinoctial orbit at its own epoch: propagate_twobody(0, 0, ..) (equinoctial_element.rs:809-867, 639-759)
__device__ __forceinline__ bool nb_initial_state(const Equinoctial &q, V3 &pos, V3 &vel) {
  const double a = q.a, h = q.h, k = q.k, p = q.p, qq = q.q;
  const double e2 = h * h + k * k;
  const double n = sqrt(kMu / ((a * a) * a));
  double lam1 = q.lambda + n * (0.0 - 0.0);
  double lon_peri = 0.0;
  if (e2 > kEps * 1e2) lon_peri = rem_euclid(atan2(h, k), kTwoPi);
  lam1 = rem_euclid(lam1, kTwoPi);
  if (lam1 < lon_peri) lam1 += kTwoPi;
  const double eps = kEps * 1e2;
  double F = kPi + lon_peri;
  int iter = 0;
  for (;;) {
    double sx, cx;
    sincos(F, &sx, &cx);
    const double f = F - k * sx + h * cx - lam1;
    const double d = 1.0 - k * cx - h * sx;
    if (fabs(f) < eps) break;
    if (fabs(d) < eps) {
      if (iter == 0) { F = F + 1.0; iter = 1; continue; }
      return false;
    }
    const double x1 = F - f / d;
    if (fabs(F - x1) < eps) { F = x1; break; }
    F = x1;
    if (++iter >= 25) return false;
  }
  const double beta = 1.0 / (1.0 + sqrt(1.0 - e2));
  const double bhk = beta * h * k;
  double sF, cF;
  sincos(F, &sF, &cF);
  const double xe = a * ((1.0 - beta * (h * h)) * cF + bhk * sF - k);
  const double ye = a * ((1.0 - beta * (k * k)) * sF + bhk * cF - h);
  const double u = 1.0 + p * p + qq * qq;
  const double inv_u = 1.0 / u;
  const double common = 2.0 * p * qq * inv_u;
  const V3 fv{(1.0 - p * p + qq * qq) * inv_u, common, -2.0 * p * inv_u};
  const V3 gv{common, (1.0 + p * p - qq * qq) * inv_u, 2.0 * qq * inv_u};
  pos = xe * fv + ye * gv;
  const double vconst = n * (a * a) / sqrt(xe * xe + ye * ye);
  const double vxe = vconst * (bhk * cF - (1.0 - beta * (h * h)) * sF);
  const double vye = vconst * ((1.0 - beta * (k * k)) * cF - bhk * sF);
  vel = vxe * fv + vye * gv;
  return true;
}

// DOP853 from t = 0 to t = span for this lane's 6 components y (in / out).  Every lane of the WARP must call it
// (groups without work pass span = 0).  Returns 0, or OUTFIT_ST_NBODY_FAILED; *steps = accepted steps of the group.
__device__ __forceinline__ int nb_dop853(const NbPert &P, int role, double (&y)[6], double span, double atol, double rtol,
                                         unsigned max_steps, double *ksm /* &smem[threadIdx.x] */, unsigned lane, unsigned *steps) {
  const double direction = span >= 0.0 ? 1.0 : -1.0;
  const double interval = fabs(span);
  const bool dummy = role == 7;
  const double n_comp = 42.0;
  auto K = [&](int s, int c) -> double & { return ksm[(size_t)(nb_stage_slot(s) * 6 + c) * kNbThreads]; };
  double f[6], w[6], ynew[6];
  unsigned nst = 0;
  int rc = 0;
  bool done = !(interval > 0.0);
  // f = rhs(y)
  nb_rhs_lane(P, role, y, nb_group_pos(y, lane), f);
  double h_abs = 0.0;
  {  // select_initial_step (Hairer II.4)
    double s0 = 0.0, s1 = 0.0;
    double sc[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      sc[c] = atol + fabs(y[c]) * rtol;
      if (!dummy) { const double a0 = y[c] / sc[c], a1 = f[c] / sc[c]; s0 += a0 * a0; s1 += a1 * a1; }
    }
    const double d0 = sqrt(nb_group_sum(s0)) / sqrt(n_comp), d1 = sqrt(nb_group_sum(s1)) / sqrt(n_comp);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    if (h0 > interval) h0 = interval;
#pragma unroll
    for (int c = 0; c < 6; ++c) w[c] = y[c] + h0 * direction * f[c];
    double f1[6];
    nb_rhs_lane(P, role, w, nb_group_pos(w, lane), f1);
    double s2 = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c)
      if (!dummy) { const double a2 = (f1[c] - f[c]) / sc[c]; s2 += a2 * a2; }
    const double d2 = h0 > 0.0 ? sqrt(nb_group_sum(s2)) / sqrt(n_comp) / h0 : 0.0 * nb_group_sum(s2);
    const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), 1.0 / 8.0);
    h_abs = fmin(fmin(100.0 * h0, h1), interval);
  }
  double t = 0.0;
  bool was_rejected = false;
  while (__any_sync(0xffffffffu, !done)) {
    // one step ATTEMPT by every lane of the warp (finished groups ride along with h = 0)
    double h = 0.0, t_new = t;
    if (!done) {
      if (nst >= max_steps) { rc = OUTFIT_ST_NBODY_FAILED; done = true; }
      const double min_step = 10.0 * fabs(nextafter(t, direction * INFINITY) - t);
      if (!was_rejected && h_abs < min_step) h_abs = min_step;
      if (!done && h_abs < min_step) { rc = OUTFIT_ST_NBODY_FAILED; done = true; }
      if (!done) {
        h = h_abs * direction;
        t_new = t + h;
        if (direction * (t_new - span) > 0.0) t_new = span;
        h = t_new - t;
        h_abs = fabs(h);
      }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) K(0, c) = f[c];
#pragma unroll 1
    for (int s = 1; s < DOP853_STAGES; ++s) {
      const double *a = DOP853_A + (s * (s - 1)) / 2;
      // y + h sum_j a_sj k_j over the non-zero a_sj (j = 0, then first(s) .. s - 1; a zero coefficient contributes an
      // exact +0): stage-major, the six components of a stage together -- each component still sums in j order
      double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      for (int j = 0; j < s; j = (j == 0 ? (s <= 2 ? 1 : (s <= 4 ? 2 : 3)) : j + 1)) {
        const double aj = a[j];
        const double *kj = ksm + (size_t)(nb_stage_slot(j) * 6) * kNbThreads;
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[c] += kj[(size_t)c * kNbThreads] * aj;
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) w[c] = y[c] + acc[c] * h;
      double dw[6];
      nb_rhs_lane(P, role, w, nb_group_pos(w, lane), dw);
#pragma unroll
      for (int c = 0; c < 6; ++c) K(s, c) = dw[c];
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double acc = 0.0;
      for (int j = 0; j < DOP853_STAGES; ++j)
        if (j == 0 || j >= 5) acc += K(j, c) * DOP853_B[j];
      ynew[c] = y[c] + h * acc;
    }
    double fnew[6];
    nb_rhs_lane(P, role, ynew, nb_group_pos(ynew, lane), fnew);
    double e5 = 0.0, e3 = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const double scl = atol + fmax(fabs(y[c]), fabs(ynew[c])) * rtol;
      double a5 = fnew[c] * DOP853_E5[DOP853_STAGES], a3 = fnew[c] * DOP853_E3[DOP853_STAGES];
      double b5 = 0.0, b3 = 0.0;
      for (int j = 0; j < DOP853_STAGES; ++j)
        if (j == 0 || j >= 5) { b5 += K(j, c) * DOP853_E5[j]; b3 += K(j, c) * DOP853_E3[j]; }
      a5 = (b5 + a5) / scl; a3 = (b3 + a3) / scl;
      if (!dummy) { e5 += a5 * a5; e3 += a3 * a3; }
    }
    e5 = nb_group_sum(e5);
    e3 = nb_group_sum(e3);
    if (!done) {
      double err = (e5 == 0.0 && e3 == 0.0) ? 0.0 : fabs(h) * e5 / sqrt((e5 + 0.01 * e3) * n_comp);
      if (!(err == err)) { rc = OUTFIT_ST_NBODY_FAILED; done = true; }
      else if (err < 1.0) {
        double factor = err == 0.0 ? 10.0 : fmin(10.0, 0.9 * pow(err, -0.125));
        if (was_rejected) factor = fmin(1.0, factor);
        h_abs *= factor;
        was_rejected = false;
        t = t_new;
#pragma unroll
        for (int c = 0; c < 6; ++c) { y[c] = ynew[c]; f[c] = fnew[c]; }
        ++nst;
        if (!(direction * (t - span) < 0.0)) done = true;
      } else {
        h_abs *= fmax(0.2, 0.9 * pow(err, -0.125));
        was_rejected = true;
      }
    }
  }
  *steps = nst;
  return rc;
}

}  // namespace ofb
