// k_lsq_nbody.cuh -- differential orbit correction with DifferentialCorrectionConfig::propagator =
// PropagatorKind::NBody(config) (differential_orbit_correction/single_iteration.rs:186-191): the partials of every
// observation come from compute_obs_and_partials_nbody (ephemeris/observation_ephemeris.rs:452-486), i.e. from
// EquinoctialElements::propagate_nbody (orbit_type/equinoctial_element.rs:908-968): the state and the element Jacobian
// J0 at the elements' reference epoch, DOP853 on [r, v, Phi] under the perturbers frozen at that epoch, and
// d pos(t_obs) / d elements = the top three rows of Phi(t_obs) J0 (propagator/nbody.rs:552-604).
//
// An integration is 2-3 orders of magnitude more work than the rest of a Newton step, and it wants eight lanes per
// observation (dev_nbody.cuh), so the persistent one-kernel state machine of the two-body path (k_lsq.cuh) is cut at
// its one expensive call:
//
//   lsqnb_init_kernel       thread / trajectory   ObsFitData::new, IOD orbit -> equinoctial, state machine at its start
//   loop on the host until no trajectory is active (one 8-byte read-back per trip):
//     lsqnb_partials_kernel 8 lanes / observation the N-body partials of every selected observation of every ACTIVE
//                                                 trajectory at the elements its next trip needs (the current ones
//                                                 for a Newton step, the last accepted linearisation point for a
//                                                 rejection pass) -> one 16-double record per observation
//     lsqnb_step_kernel     thread / trajectory   ONE trip of the state machine of diff_cor.rs:282-442 from the records:
//                                                 a Newton step (normal equations, 6x6 inverse, tests) or a rejection
//                                                 pass, the end-of-loop decisions, the result record
//
// The arithmetic of a trip is the two-body kernel's (same helpers of dev_lsq.cuh, same order); the per-trajectory state
// between trips lives in global memory (LsqNbState).  Parity with the reference is at the tolerance level, like every
// N-body number here (the reference's DOP853 crate is not vendored: dev_nbody.cuh).
#pragma once
#include "k_lsq.cuh"
#include "k_nbody.cuh"

namespace ofb {

struct LsqNbState {
  double el[7], el_lin[7], last_cov[36];
  double last_rms, prev_rms;
  unsigned long long last_nmeas, total_it, outer, inner, stagnation;
  int have_lin, converged, busy, post, fail_code;
  int phase;  // what the next trip needs: 0 nothing | 1 a Newton step at `el` | 2 a rejection pass at `el_lin`
};
struct LsqNbRec {  // compute_obs_and_partials_nbody of one observation
  double ok, ra, dec, pr[6], pd[6], _pad;
};

// the end of a trip (diff_cor.rs:340-428): the next phase, or the result record
__device__ __forceinline__ void lsqnb_decide(LsqNbState &st, const LsqCfgDev &C, unsigned num_free, const OutfitIodResult &iod,
                                             OutfitLsqResult *res, OutfitObsFit *F, unsigned n_obs) {
  st.phase = 0;
  if (!st.post) {
    if (st.inner >= C.max_newton_iterations) st.post = 1;
    else { st.phase = 1; return; }
  }
  const bool finish = st.fail_code != 0 || !C.enable_outlier_rejection ||
                      (st.outer == 0 && st.last_rms < C.convergence_before_rejection_threshold) || !st.converged || !st.have_lin;
  if (!finish) { st.phase = 2; return; }
  res->status = OUTFIT_ST_OK;
  res->total_newton_iterations = st.total_it;
  if (st.fail_code) {  // Err(_) => Ok(initial_orbit) (mod.rs:113)
    res->kind = OUTFIT_LSQ_IOD_FALLBACK;
    res->fallback_cause = st.fail_code;
    res->epoch = iod.epoch;
    for (int j = 0; j < 6; ++j) res->elem[j] = iod.elem[j];
    res->normalised_rms = iod.rms;
    for (int i = 0; i < 36; ++i) res->normal_matrix[i] = 0.0;
    for (unsigned i = 0; i < n_obs; ++i) { F[i].residual_ra = 0.0; F[i].residual_dec = 0.0; F[i].chi = 0.0; F[i].selection = 0; }
  } else {  // rescale_covariance (least_square.rs:371-394)
    double mu = 1.0;
    if (num_free < st.last_nmeas) {
      const double factor = sqrt((double)st.last_nmeas / (double)(st.last_nmeas - num_free));
      mu = st.last_rms > 1.0 ? st.last_rms * factor : factor;
    }
    const double mu2 = mu * mu;
    res->kind = OUTFIT_LSQ_CORRECTED;
    res->epoch = st.el[0];
    for (int j = 0; j < 6; ++j) res->elem[j] = st.el[1 + j];
#pragma unroll 1
    for (int i = 0; i < 36; ++i) {
      res->covariance[i] = st.last_cov[i] * mu2;
      res->normal_matrix[i] = res->normal_matrix[i] / mu2;  // held unscaled since the last accepted step
    }
#pragma unroll 1
    for (int j = 0; j < 6; ++j) res->sigma[j] = sqrt(st.last_cov[7 * j] * mu2);
    res->normalised_rms = st.last_rms;
    res->num_measurements = st.last_nmeas;
  }
  st.busy = 0;
}

__global__ void __launch_bounds__(128)
lsqnb_init_kernel(LsqBatchDev B, LsqCfgDev C, const OutfitIodResult *__restrict__ iod, OutfitLsqResult *__restrict__ out,
                  OutfitObsFit *__restrict__ fit, LsqNbState *__restrict__ state, unsigned *__restrict__ obs_traj,
                  unsigned long long *__restrict__ active) {
  const unsigned long long tr = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tr >= B.n_traj) return;
  unsigned num_free = 0;
  for (int j = 0; j < 6; ++j) num_free += C.free_el[j] ? 1u : 0u;
  const unsigned long long o0 = B.traj_offset[tr];
  const unsigned n_obs = (unsigned)(B.traj_offset[tr + 1] - o0);
  OutfitLsqResult *res = out + tr;
  OutfitObsFit *F = fit + o0;
  int ist = iod[tr].status;
  for (unsigned i = 0; i < n_obs; ++i) {  // ObsFitData::new (obs_fit_data.rs:105-116)
    F[i].residual_ra = 0.0; F[i].residual_dec = 0.0; F[i].chi = 0.0; F[i].selection = 0; F[i]._pad0 = 0;
    obs_traj[o0 + i] = (unsigned)tr;
    if (B.obs_status[o0 + i] != 0) ist = OUTFIT_ST_EPHEM_OUT_OF_RANGE;  // the reference panics (horizon_data.rs:722)
  }
  {
    double *z = reinterpret_cast<double *>(res);
    for (unsigned i = 0; i < sizeof(OutfitLsqResult) / 8; ++i) z[i] = 0.0;
  }
  LsqNbState st;
  {
    double *z = reinterpret_cast<double *>(&st);
    for (unsigned i = 0; i < sizeof(LsqNbState) / 8; ++i) z[i] = 0.0;
  }
  if (ist != OUTFIT_ST_OK) {
    res->status = ist; res->kind = OUTFIT_LSQ_NONE;
  } else {
    Orbit orb;
    orb.kind = iod[tr].element_kind; orb.corrected = iod[tr].corrected; orb.epoch = iod[tr].epoch;
    for (int j = 0; j < 6; ++j) orb.e[j] = iod[tr].elem[j];
    Equinoctial q;
    const int rq = to_equinoctial(orb, q);
    if (rq != 0) {
      res->status = rq; res->kind = OUTFIT_LSQ_NONE;
    } else {
      st.el[0] = q.epoch; st.el[1] = q.a; st.el[2] = q.h; st.el[3] = q.k; st.el[4] = q.p; st.el[5] = q.q; st.el[6] = q.lambda;
      st.last_rms = 1.7976931348623157e308; st.prev_rms = 1.7976931348623157e308;
      st.busy = 1;
      lsqnb_decide(st, C, num_free, iod[tr], res, F, n_obs);
      if (st.busy) atomicAdd(active, 1ull);
    }
  }
  state[tr] = st;
}

// gm [n_pert]; pert_pos [n_pert][3][n_traj] = the perturbers at each trajectory's IOD epoch (= the epoch of its elements)
__global__ void __launch_bounds__(kNbThreads, OUTFIT_NB_BPS)
lsqnb_partials_kernel(LsqBatchDev B, NbCfgDev cfg, const double *__restrict__ gm, const double *__restrict__ pert_pos,
                      const LsqNbState *__restrict__ state, const OutfitObsFit *__restrict__ fit,
                      const unsigned *__restrict__ obs_traj, LsqNbRec *__restrict__ rec) {
  extern __shared__ __align__(16) double nb_sm[];
  const unsigned lane = threadIdx.x & 31u;
  const int role = (int)(lane & 7u);
  const size_t gI = ((size_t)blockIdx.x * kNbThreads + threadIdx.x) >> 3;
  const bool live = gI < B.n_obs;
  const size_t tr = live ? obs_traj[gI] : 0;
  const LsqNbState *st = state + tr;
  const int phase = live && st->busy ? st->phase : 0;
  const bool need = phase != 0 && fit[gI].selection == 0;
  if (!__any_sync(0xffffffffu, need)) return;  // a warp whose four observations all belong to finished trajectories
  double el[7];
  for (int j = 0; j < 7; ++j) el[j] = phase == 2 ? st->el_lin[j] : st->el[j];
  NbPert P;
  P.n = (int)cfg.n_pert;
  for (int p = 0; p < kNbMaxPert; ++p) {
    if (p < P.n) {
      P.gm[p] = gm[p];
      P.pos[p] = V3{pert_pos[((size_t)p * 3 + 0) * B.n_traj + tr], pert_pos[((size_t)p * 3 + 1) * B.n_traj + tr],
                    pert_pos[((size_t)p * 3 + 2) * B.n_traj + tr]};
    } else {
      P.gm[p] = 0.0;
      P.pos[p] = V3{0.0, 0.0, 0.0};
    }
  }
  nb_prepare(P);
  bool ok = need;
  double y[6] = {0, 0, 0, 0, 0, 0};
  if (role >= 1 && role < 7) y[role - 1] = 1.0;
  if (need) {  // propagate_twobody(0, 0, ..): the state at the reference epoch
    V3 p0, v0, col[6], colv[6];
    ok = lsq_state_and_columns<false>(el, 0.0, p0, v0, col, colv);
    if (role == 0) { y[0] = p0.x; y[1] = p0.y; y[2] = p0.z; y[3] = v0.x; y[4] = v0.y; y[5] = v0.z; }
  }
  double span = ok ? B.mjd_tt[gI] - el[0] : 0.0;
  if (fabs(span) < 1e-14) span = 0.0;  // equinoctial_element.rs:923-931
  if (!(span == span)) { span = 0.0; ok = false; }
  unsigned nst = 0;
  const int rc = nb_dop853(P, role, y, span, cfg.atol, cfg.rtol, cfg.max_steps, nb_sm + threadIdx.x, lane, &nst);
  // rows 0..2 of Phi(t_obs): column k lives on lane k + 1 of the group
  double phi[3][6];
#pragma unroll
  for (int k = 0; k < 6; ++k)
#pragma unroll
    for (int c = 0; c < 3; ++c) phi[c][k] = __shfl_sync(0xffffffffu, y[c], (int)(lane & ~7u) + k + 1);
  if (role != 0 || !need) return;
  LsqNbRec r;
  r.ok = 0.0; r.ra = 0.0; r.dec = 0.0; r._pad = 0.0;
  for (int j = 0; j < 6; ++j) { r.pr[j] = 0.0; r.pd[j] = 0.0; }
  if (ok && rc == 0) {
    V3 p0, v0, col[6], colv[6];
    lsq_state_and_columns<true>(el, 0.0, p0, v0, col, colv);  // J0 (recomputed: 36 doubles do not ride through the integration)
    V3 dcol[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const double j0[6] = {col[j].x, col[j].y, col[j].z, colv[j].x, colv[j].y, colv[j].z};
      double acc[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        acc[c] = phi[c][0] * j0[0];
#pragma unroll
        for (int k = 1; k < 6; ++k) acc[c] = phi[c][k] * j0[k] + acc[c];
      }
      dcol[j] = V3{acc[0], acc[1], acc[2]};
    }
    const V3 obs{__ldg(B.scorer + gI), __ldg(B.scorer + B.n_obs + gI), __ldg(B.scorer + 2 * B.n_obs + gI)};
    lsq_topocentric(V3{y[0], y[1], y[2]}, V3{y[3], y[4], y[5]}, dcol, obs, r.ra, r.dec, r.pr, r.pd);
    r.ok = 1.0;
  }
  rec[gI] = r;
}

__global__ void __launch_bounds__(64)
lsqnb_step_kernel(LsqBatchDev B, LsqCfgDev C, const OutfitIodResult *__restrict__ iod, OutfitLsqResult *__restrict__ out,
                  OutfitObsFit *__restrict__ fit, double *__restrict__ tmp, LsqNbState *__restrict__ state,
                  const LsqNbRec *__restrict__ rec, unsigned long long *__restrict__ active) {
  const unsigned long long tr = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tr >= B.n_traj) return;
  if (!state[tr].busy) return;
  const double kMax = 1.7976931348623157e308;
  unsigned num_free = 0;
  for (int j = 0; j < 6; ++j) num_free += C.free_el[j] ? 1u : 0u;
  LsqNbState st = state[tr];
  const unsigned long long o0 = B.traj_offset[tr];
  const unsigned n_obs = (unsigned)(B.traj_offset[tr + 1] - o0);
  OutfitLsqResult *res = out + tr;
  OutfitObsFit *F = fit + o0;
  double *t_rra = tmp + o0, *t_rdec = tmp + B.n_obs + o0, *t_chi = tmp + 2 * B.n_obs + o0;
  if (st.phase == 1) {
    // single_iteration (single_iteration.rs:140-317) + solve_weighted_least_squares (least_square.rs:225-327)
    ++st.inner;
    ++st.total_it;
    double nm[36], cov[36], work[36];
    for (int i = 0; i < 36; ++i) nm[i] = 0.0;
    double rhs[6] = {0, 0, 0, 0, 0, 0};
    double qsum = 0.0;
    unsigned long long n_active = 0;
    for (unsigned i = 0; i < n_obs; ++i) {
      const unsigned long long gI = o0 + i;
      t_rra[i] = F[i].residual_ra; t_rdec[i] = F[i].residual_dec; t_chi[i] = F[i].chi;
      if (F[i].selection != 0) continue;
      const LsqNbRec *r = rec + gI;
      if (r->ok == 0.0) continue;
      double pr[6], pd[6];
      for (int j = 0; j < 6; ++j) { pr[j] = r->pr[j]; pd[j] = r->pd[j]; }
      const double sra = __ldg(B.sigma_ra + gI), sdec = __ldg(B.sigma_dec + gI);
      const double xr = lsq_angular_diff(__ldg(B.ra + gI) - 0.0, r->ra);
      const double xd = (__ldg(B.dec + gI) - 0.0) - r->dec;
      const double ca = xr / sra, cd = xd / sdec;
      t_rra[i] = xr; t_rdec[i] = xd; t_chi[i] = sqrt(ca * ca + cd * cd);
      const double wr = 1.0 / (sra * sra), wd = 1.0 / (sdec * sdec), wc = 0.0;
      ++n_active;
      for (int j = 0; j < 6; ++j) {
        for (int k = 0; k < 6; ++k)
          OFB_M6(nm, j, k) += pr[j] * wr * pr[k] + pd[j] * wd * pd[k] + wc * (pd[j] * pr[k] + pr[j] * pd[k]);
        rhs[j] += (pr[j] * wr + pd[j] * wc) * xr + (pr[j] * wc + pd[j] * wd) * xd;
      }
      qsum += wr * xr * xr + wd * xd * xd + 2.0 * wc * xr * xd;
    }
    const unsigned long long nmeas = 2 * n_active;
    for (int j = 0; j < 6; ++j)
      if (!C.free_el[j]) {
        for (int k = 0; k < 6; ++k) { OFB_M6(nm, j, k) = 0.0; OFB_M6(nm, k, j) = 0.0; }
        OFB_M6(nm, j, j) = 1.0;
        rhs[j] = 0.0;
      }
    const bool inv_ok = lsq_invert_normal_matrix(nm, cov, work);
    double dx[6] = {0, 0, 0, 0, 0, 0};
    if (inv_ok) lsq_gemv6(cov, rhs, dx);
    for (int j = 0; j < 6; ++j)
      if (!C.free_el[j]) dx[j] = 0.0;
    const double new_rms = nmeas > 0 ? sqrt(qsum / (double)nmeas) : 0.0;
    double cdx[6];
    lsq_gemv6(nm, dx, cdx);
    const double cnorm = sqrt(lsq_dot6(dx, cdx));
    double corrected[6];
    for (int j = 0; j < 6; ++j) corrected[j] = C.free_el[j] ? st.el[1 + j] + dx[j] : st.el[1 + j];
    if (!inv_ok) { st.fail_code = OUTFIT_ST_LSQ_INVERSION; st.post = 1; }
    else if (lsq_is_bizarre(corrected, C)) { st.fail_code = OUTFIT_ST_LSQ_BIZARRE; st.post = 1; }
    else if (st.prev_rms < kMax && new_rms / st.prev_rms >= C.rms_divergence_ratio) { st.fail_code = OUTFIT_ST_LSQ_DIVERGED; st.post = 1; }
    else {
      const bool stagnated = st.prev_rms < kMax && new_rms / st.prev_rms >= C.rms_stagnation_ratio;
      bool stop = false;
      if (stagnated) {
        if (++st.stagnation >= C.max_stagnation_iterations) stop = true;
      } else {
        st.stagnation = 0;
      }
      if (stop) {
        st.post = 1;
      } else {  // advance the state
        for (int j = 0; j < 7; ++j) st.el_lin[j] = st.el[j];
        st.have_lin = 1;
        for (int i = 0; i < 36; ++i) { res->normal_matrix[i] = nm[i]; st.last_cov[i] = cov[i]; }
        st.last_rms = new_rms;
        st.last_nmeas = nmeas;
        for (int j = 0; j < 6; ++j) st.el[1 + j] = corrected[j];
        for (unsigned i = 0; i < n_obs; ++i) { F[i].residual_ra = t_rra[i]; F[i].residual_dec = t_rdec[i]; F[i].chi = t_chi[i]; }
        st.prev_rms = new_rms;
        if (cnorm < C.convergence_threshold) { st.converged = 1; st.post = 1; }
      }
    }
  } else if (st.phase == 2) {
    // update_observation_selection (outlier_rejection.rs:118-235) with the equations of the last accepted step
    unsigned long long changes = 0;
    for (unsigned i = 0; i < n_obs; ++i) {
      const unsigned long long gI = o0 + i;
      const int sel = F[i].selection;
      if (sel == 2) continue;
      double pr[6] = {0, 0, 0, 0, 0, 0}, pd[6] = {0, 0, 0, 0, 0, 0};
      double wr = 1.0, wd = 1.0;
      const double sra = __ldg(B.sigma_ra + gI), sdec = __ldg(B.sigma_dec + gI);
      if (sel == 0) {
        const LsqNbRec *r = rec + gI;
        if (r->ok != 0.0) {
          for (int j = 0; j < 6; ++j) { pr[j] = r->pr[j]; pd[j] = r->pd[j]; }
          wr = 1.0 / (sra * sra); wd = 1.0 / (sdec * sdec);
        }
      }
      const double var_ra = sra * sra, var_dec = sdec * sdec;
      const double cov_cross = -sra * sdec * 0.0 / (wr * wd);
      double gga[6], ggd[6];
      lsq_gemv6(st.last_cov, pr, gga);
      lsq_gemv6(st.last_cov, pd, ggd);
      const double paa = lsq_dot6(pr, gga), pdd = lsq_dot6(pd, ggd), pad = lsq_dot6(pr, ggd);
      const double v00 = var_ra - paa, v01 = cov_cross - pad, v11 = var_dec - pdd;
      const double det = v00 * v11 - v01 * v01;
      const double scale = fmax(fabs(v00), fabs(v11));
      if (fabs(det) < kEps * scale * scale || scale == 0.0) continue;
      const double i00 = v11 / det, i01 = -v01 / det, i10 = -v01 / det, i11 = v00 / det;
      const double rr = F[i].residual_ra, rd = F[i].residual_dec;
      double y0 = i00 * rr, y1 = i10 * rr;
      y0 = i01 * rd + y0;
      y1 = i11 * rd + y1;
      const double chi2 = rr * y0 + rd * y1;
      if (sel == 0 && chi2 > C.chi2_reject) { F[i].selection = 1; ++changes; }
      else if (sel == 1 && chi2 <= C.chi2_recover) { F[i].selection = 0; ++changes; }
    }
    bool finish = false;
    if (changes == 0) finish = true;
    else if (++st.outer > C.max_outlier_rejection_passes) finish = true;
    else { st.inner = 0; st.prev_rms = kMax; st.stagnation = 0; st.converged = 0; st.have_lin = 0; st.post = 0; }
    if (finish) {
      // the end-of-loop block with `finish` already decided: the result record
      st.converged = 0;  // makes lsqnb_decide's own test finish (the flags are not read again)
    }
  }
  lsqnb_decide(st, C, num_free, iod[tr], res, F, n_obs);
  if (st.busy) atomicAdd(active, 1ull);
  state[tr] = st;
}

}  // namespace ofb
