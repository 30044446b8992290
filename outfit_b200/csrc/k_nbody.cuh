// k_nbody.cuh -- bulk EquinoctialElements::propagate_nbody (orbit_type/equinoctial_element.rs:908-968): one orbit per
// group of eight lanes, frozen perturbers, DOP853 on [r, v, Phi] (dev_nbody.cuh).
#pragma once
#include "../../include/outfit_b200.h"
#include "dev_nbody.cuh"

using namespace ofb;

struct NbCfgDev {
  double atol, rtol;
  unsigned n_pert, max_steps;
};

// kind / epoch / elem [6][n] as in the ephemeris entries; t1 [n]; gm [n_pert]; pert_pos [n_pert][3][n] = heliocentric
// position (ecliptic J2000, AU) of every perturber at each orbit's reference epoch (build_perturber_snapshots,
// nbody.rs:453-476).  out [6][n] = position, velocity at t1 (ecliptic J2000); stm [36][n] column-major or null;
// status [n]; steps [n] or null.
__global__ void __launch_bounds__(kNbThreads, OUTFIT_NB_BPS)
propagate_nbody_kernel(size_t n, const int *__restrict__ kind, const double *__restrict__ epoch,
                       const double *__restrict__ elem, const double *__restrict__ t1, NbCfgDev cfg,
                       const double *__restrict__ gm, const double *__restrict__ pert_pos, double *__restrict__ out,
                       double *__restrict__ stm, int *__restrict__ status, unsigned *__restrict__ steps) {
  extern __shared__ __align__(16) double nb_sm[];
  const unsigned lane = threadIdx.x & 31u;
  const int role = (int)(lane & 7u);
  const size_t item = ((size_t)blockIdx.x * kNbThreads + threadIdx.x) >> 3;
  const bool live = item < n;
  const size_t ic = live ? item : 0;
  NbPert P;
  P.n = (int)cfg.n_pert;
  for (int p = 0; p < kNbMaxPert; ++p) {
    if (p < P.n) {
      P.gm[p] = gm[p];
      P.pos[p] = V3{pert_pos[((size_t)p * 3 + 0) * n + ic], pert_pos[((size_t)p * 3 + 1) * n + ic], pert_pos[((size_t)p * 3 + 2) * n + ic]};
    } else {
      P.gm[p] = 0.0;
      P.pos[p] = V3{0.0, 0.0, 0.0};
    }
  }
  nb_prepare(P);
  int st = OUTFIT_ST_OK;
  Equinoctial eq;
  {
    const int kd = kind[ic];
    if (kd == 1) {
      eq.epoch = epoch[ic];
      eq.a = elem[ic]; eq.h = elem[n + ic]; eq.k = elem[2 * n + ic];
      eq.p = elem[3 * n + ic]; eq.q = elem[4 * n + ic]; eq.lambda = elem[5 * n + ic];
    } else {
      Orbit orb;
      orb.kind = kd; orb.corrected = 0; orb.epoch = epoch[ic];
#pragma unroll
      for (int q = 0; q < 6; ++q) orb.e[q] = elem[(size_t)q * n + ic];
      const int rq = to_equinoctial(orb, eq);
      if (rq != 0) st = rq;
    }
  }
  if (st == 0 && !(sqrt(eq.h * eq.h + eq.k * eq.k) < 1.0)) st = OUTFIT_ST_INVALID_ORBIT;  // the two-body start needs e < 1
  V3 p0 = V3{0, 0, 0}, v0 = V3{0, 0, 0};
  if (st == 0 && !nb_initial_state(eq, p0, v0)) st = OUTFIT_ST_ROOT_FINDING;
  double y[6] = {0, 0, 0, 0, 0, 0};
  if (role == 0) { y[0] = p0.x; y[1] = p0.y; y[2] = p0.z; y[3] = v0.x; y[4] = v0.y; y[5] = v0.z; }
  else if (role < 7) y[role - 1] = 1.0;
  double span = (live && st == 0) ? t1[ic] - eq.epoch : 0.0;
  if (fabs(span) < 1e-14) span = 0.0;  // equinoctial_element.rs:923-931
  if (!(span == span)) { span = 0.0; if (st == 0) st = OUTFIT_ST_NBODY_FAILED; }
  unsigned nst = 0;
  const int rc = nb_dop853(P, role, y, span, cfg.atol, cfg.rtol, cfg.max_steps, nb_sm + threadIdx.x, lane, &nst);
  if (st == 0 && rc != 0) st = rc;
  if (!live) return;
  if (role == 0) {
#pragma unroll
    for (int c = 0; c < 6; ++c) out[(size_t)c * n + item] = st == 0 ? y[c] : NAN;
    status[item] = st;
    if (steps) steps[item] = nst;
  } else if (role < 7 && stm) {
#pragma unroll
    for (int c = 0; c < 6; ++c) stm[(size_t)(6 * (role - 1) + c) * n + item] = st == 0 ? y[c] : NAN;
  }
}

// PropagatorKind::NBody of the ephemeris path (propagator/mod.rs:93-101): every (epoch e, orbit i) entry is its OWN
// integration from the orbit's reference epoch to mjd_tt[e] -- what the reference does, entry by entry -- by one group
// of eight lanes; the state is rotated to the equatorial frame (ecl_state_to_equ) and written to
// state [6][n_epochs][n_orbits], status [n_epochs][n_orbits] for ephemeris_twobody_kernel<.., NBODY = true>.
__global__ void __launch_bounds__(kNbThreads, OUTFIT_NB_BPS)
ephemeris_nbody_state_kernel(size_t n_orbits, const int *__restrict__ kind, const double *__restrict__ epoch,
                             const double *__restrict__ elem, size_t n_epochs, const double *__restrict__ mjd_tt,
                             NbCfgDev cfg, const double *__restrict__ gm, const double *__restrict__ pert_pos,
                             double *__restrict__ state, int *__restrict__ status) {
  extern __shared__ __align__(16) double nb_sm[];
  const unsigned lane = threadIdx.x & 31u;
  const int role = (int)(lane & 7u);
  const size_t n_ent = n_orbits * n_epochs;
  const size_t item = ((size_t)blockIdx.x * kNbThreads + threadIdx.x) >> 3;  // = e * n_orbits + i
  const bool live = item < n_ent;
  const size_t it = live ? item : 0;
  const size_t e = it / n_orbits, i = it - e * n_orbits;
  NbPert P;
  P.n = (int)cfg.n_pert;
  for (int p = 0; p < kNbMaxPert; ++p) {
    if (p < P.n) {
      P.gm[p] = gm[p];
      P.pos[p] = V3{pert_pos[((size_t)p * 3 + 0) * n_orbits + i], pert_pos[((size_t)p * 3 + 1) * n_orbits + i],
                    pert_pos[((size_t)p * 3 + 2) * n_orbits + i]};
    } else {
      P.gm[p] = 0.0;
      P.pos[p] = V3{0.0, 0.0, 0.0};
    }
  }
  nb_prepare(P);
  int st = OUTFIT_ST_OK;
  Equinoctial eq;
  {
    const int kd = kind[i];
    if (kd == 1) {
      eq.epoch = epoch[i];
      eq.a = elem[i]; eq.h = elem[n_orbits + i]; eq.k = elem[2 * n_orbits + i];
      eq.p = elem[3 * n_orbits + i]; eq.q = elem[4 * n_orbits + i]; eq.lambda = elem[5 * n_orbits + i];
    } else {
      Orbit orb;
      orb.kind = kd; orb.corrected = 0; orb.epoch = epoch[i];
#pragma unroll
      for (int q = 0; q < 6; ++q) orb.e[q] = elem[(size_t)q * n_orbits + i];
      if (to_equinoctial(orb, eq) != 0) st = OUTFIT_ST_INVALID_CONVERSION;  // mod.rs:196-213
    }
  }
  if (st == 0 && !(sqrt(eq.h * eq.h + eq.k * eq.k) < 1.0)) st = OUTFIT_ST_INVALID_CONVERSION;  // check_elliptical_orbit, mod.rs:219-240
  V3 p0 = V3{0, 0, 0}, v0 = V3{0, 0, 0};
  if (st == 0 && !nb_initial_state(eq, p0, v0)) st = OUTFIT_ST_ROOT_FINDING;
  double y[6] = {0, 0, 0, 0, 0, 0};
  if (role == 0) { y[0] = p0.x; y[1] = p0.y; y[2] = p0.z; y[3] = v0.x; y[4] = v0.y; y[5] = v0.z; }
  else if (role < 7) y[role - 1] = 1.0;
  double span = (live && st == 0) ? mjd_tt[e] - eq.epoch : 0.0;
  if (fabs(span) < 1e-14) span = 0.0;
  if (!(span == span)) { span = 0.0; if (st == 0) st = OUTFIT_ST_NBODY_FAILED; }
  unsigned nst = 0;
  const int rc = nb_dop853(P, role, y, span, cfg.atol, cfg.rtol, cfg.max_steps, nb_sm + threadIdx.x, lane, &nst);
  if (st == 0 && rc != 0) st = rc;
  if (!live || role != 0) return;
  const V3 ap = ecl_to_equ(V3{y[0], y[1], y[2]}), av = ecl_to_equ(V3{y[3], y[4], y[5]});
  state[item] = ap.x; state[n_ent + item] = ap.y; state[2 * n_ent + item] = ap.z;
  state[3 * n_ent + item] = av.x; state[4 * n_ent + item] = av.y; state[5 * n_ent + item] = av.z;
  status[item] = st;
}
