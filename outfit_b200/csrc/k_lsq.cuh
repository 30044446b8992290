// k_lsq.cuh -- differential orbit correction (FitLSQ) with FOUR lanes per trajectory (algorithm, reference files
// and arithmetic: dev_lsq.cuh).
//
//   * lane 0 of a quad (the leader) owns the control flow of the trajectory: the Newton / rejection state
//     machine, the convergence tests and the scalars of the result record;
//   * the observation loops -- predicted position + 12 partials per observation (an equinoctial Kepler solve
//     and ~400 flops each) in the Newton step, the same partials + two 6x6 products in the rejection pass -- run
//     four observations at a time, observation i on lane i mod 4;
//   * the normal matrix is accumulated by all four lanes, each owning 9 of its 36 entries (entry e on lane
//     e mod 4), every entry summed over the observations IN OBSERVATION ORDER, so each entry sees exactly the
//     additions of a serial loop over the observations;
//   * the Cholesky factor is built row-parallel (row i on lane i mod 4), the six columns of the inverse
//     column-parallel, the covariance / normal-matrix rescaling entry-parallel -- each number by the serial
//     algorithm's operations in the serial order;
//   * the matrices (normal matrix, covariance, factorisation work space, last accepted covariance), the two
//     element vectors and the per-round partials live in shared memory, one contiguous block of kQsSlots (odd)
//     doubles per quad, so the eight leaders of a warp fall in eight different banks.
//
// A persistent grid: quads fetch trajectories from a work counter and advance them one Newton step (or one
// rejection pass) per loop trip, so a quad whose trajectory is done (2 steps for a diverging start, 3-8 for a
// converging one) takes the next one instead of idling.
//
// Round 1 ran one lane per trajectory with the five 6x6 matrices in local memory (255 registers, 1.5 KB of
// spills, 3.3-3.4 ms on 100k x 12).  On the way here, with the records bit-identical at every step
// (profiles/r2h_lsq_ab*.log): four lanes alone 4.2 ms (slower: instruction fetch, `no_instruction` at 10 stalled
// warps per issue -- one trip of the state machine swept 187 KB of code); 6x6 loops rolled 3.3 ms; contiguous
// blocks + parallel inverse columns 2.95 ms; lanes leaving the Kepler iteration of the partials together (they
// ran its 400-flop tail at 12 of 32 lanes) + transcribed sincos / atan2 1.99 ms; parallel Cholesky rows and
// result scaling 1.76 ms.
#pragma once
#include "dev_lsq.cuh"

struct LsqBatchDev {
  unsigned long long n_traj, n_obs;
  const unsigned long long *traj_offset;
  const double *mjd_tt, *ra, *dec, *sigma_ra, *sigma_dec;
  const double *scorer;  // [3][n_obs] observer position, equatorial J2000 (scorer_observer_kernel)
  const int *obs_status; // [n_obs] 0 | OUTFIT_ST_EPHEM_OUT_OF_RANGE
};

namespace ofb {

constexpr int kLsqQThreads = 64;
constexpr int kLsqQuads = kLsqQThreads / 4;
#ifndef OUTFIT_LSQQ_BPS
#define OUTFIT_LSQQ_BPS 6  // 168 registers, 12 warps per SM: 1.76 ms per 100 k trajectories against 2.00 at 5 blocks and 1.97 at 7
                           // (128 registers, more spills), same bytes
#endif
// shared slots of one quad
constexpr int kQsNm = 0, kQsCov = 36, kQsWork = 72, kQsLastCov = 108, kQsEl = 144, kQsElLin = 151, kQsRhs = 158,
              kQsQsum = 164, kQsBv = 165, kQsSlots = kQsBv + 4 * 16;

// lsq_cholesky6 (dev_lsq.cuh) on the four lanes of a quad: row i of the factor belongs to lane i mod 4.  Column j
// needs rows j..5 of the columns before it, final since their own step; every element receives the updates of
// the serial loop in the same order (k ascending), so the factor is the same to the bit.
__device__ __forceinline__ bool lsq_cholesky6_quad(double *m, unsigned q, unsigned qmask) {
#pragma unroll 1
  for (int j = 0; j < 6; ++j) {
    for (int i = (int)q; i < 6; i += 4)
      if (i >= j)
        for (int k = 0; k < j; ++k) {
          const double factor = -OFB_M6(m, j, k);
          OFB_M6(m, i, j) = factor * OFB_M6(m, i, k) + OFB_M6(m, i, j);
        }
    __syncwarp(qmask);
    const double diag = OFB_M6(m, j, j);
    if (diag == 0.0 || !(diag >= 0.0)) return false;  // the same value on the four lanes
    const double denom = sqrt(diag);
    __syncwarp(qmask);
    for (int i = (int)q; i < 6; i += 4) {
      if (i == j) OFB_M6(m, j, j) = denom;
      else if (i > j) OFB_M6(m, i, j) = OFB_M6(m, i, j) / denom;
    }
    __syncwarp(qmask);
  }
  return true;
}

__global__ void __launch_bounds__(kLsqQThreads, OUTFIT_LSQQ_BPS)
lsq_quad_kernel(LsqBatchDev B, LsqCfgDev C, const OutfitIodResult *__restrict__ iod, OutfitLsqResult *__restrict__ out,
                OutfitObsFit *__restrict__ fit, double *__restrict__ tmp, unsigned long long *__restrict__ next) {
  constexpr int S = 1;
  static_assert(kQsSlots % 2 == 1, "an odd block length spreads the quads of a warp over the banks");
  __shared__ double qsm[kQsSlots * kLsqQuads];
  const double kMax = 1.7976931348623157e308;
  const unsigned lane = threadIdx.x & 31u, q = lane & 3u;
  const unsigned qmask = 0xFu << (lane & ~3u);
  const int leader = (int)(lane & ~3u);
  const bool lead = q == 0;
  double *qs = qsm + (threadIdx.x >> 2) * kQsSlots;
#define QS(slot) qs[(slot)]
  unsigned num_free = 0;
  for (int j = 0; j < 6; ++j) num_free += C.free_el[j] ? 1u : 0u;
  // trajectory of the quad (every lane holds the addresses; the state machine is the leader's)
  unsigned long long tr = 0, o0 = 0;
  unsigned n_obs = 0;
  OutfitLsqResult *res = nullptr;
  OutfitObsFit *F = nullptr;
  double *t_rra = nullptr, *t_rdec = nullptr, *t_chi = nullptr;
  double last_rms = kMax, prev_rms = kMax;
  unsigned long long last_nmeas = 0, total_it = 0, outer = 0, inner = 0, stagnation = 0;
  bool have_lin = false, converged = false, busy = false, exhausted = false, post = false;
  int fail_code = 0;
  for (;;) {
    if (!busy && !exhausted) {  // fetch and set up the next trajectory (busy / exhausted are quad-uniform)
      if (lead) tr = atomicAdd(next, 1ull);
      tr = __shfl_sync(qmask, tr, leader);
      if (tr >= B.n_traj) {
        exhausted = true;
      } else {
        o0 = B.traj_offset[tr];
        n_obs = (unsigned)(B.traj_offset[tr + 1] - o0);
        res = out + tr;
        F = fit + o0;
        t_rra = tmp + o0; t_rdec = tmp + B.n_obs + o0; t_chi = tmp + 2 * B.n_obs + o0;
        int bad = 0;
        for (unsigned i = q; i < n_obs; i += 4) {  // ObsFitData::new (obs_fit_data.rs:105-116)
          F[i].residual_ra = 0.0; F[i].residual_dec = 0.0; F[i].chi = 0.0; F[i].selection = 0; F[i]._pad0 = 0;
          if (B.obs_status[o0 + i] != 0) bad = 1;  // the reference panics (horizon_data.rs:722)
        }
        bad |= __shfl_xor_sync(qmask, bad, 1);
        bad |= __shfl_xor_sync(qmask, bad, 2);
        {
          double *z = reinterpret_cast<double *>(res);
          for (unsigned i = q; i < sizeof(OutfitLsqResult) / 8; i += 4) z[i] = 0.0;
        }
        for (int i = (int)q; i < 36; i += 4) QS(kQsLastCov + i) = 0.0;
        __syncwarp(qmask);
        int start = 0;
        if (lead) {
          int ist = iod[tr].status;
          if (bad) ist = OUTFIT_ST_EPHEM_OUT_OF_RANGE;
          if (ist != OUTFIT_ST_OK) {
            res->status = ist; res->kind = OUTFIT_LSQ_NONE;
          } else {
            Orbit orb;
            orb.kind = iod[tr].element_kind; orb.corrected = iod[tr].corrected; orb.epoch = iod[tr].epoch;
            for (int j = 0; j < 6; ++j) orb.e[j] = iod[tr].elem[j];
            Equinoctial eq;
            const int rq = to_equinoctial(orb, eq);
            if (rq != 0) {
              res->status = rq; res->kind = OUTFIT_LSQ_NONE;
            } else {
              QS(kQsEl + 0) = eq.epoch; QS(kQsEl + 1) = eq.a; QS(kQsEl + 2) = eq.h; QS(kQsEl + 3) = eq.k;
              QS(kQsEl + 4) = eq.p; QS(kQsEl + 5) = eq.q; QS(kQsEl + 6) = eq.lambda;
              last_rms = kMax; last_nmeas = 0; total_it = 0; fail_code = 0;
              outer = 0; inner = 0; prev_rms = kMax; stagnation = 0; converged = false; have_lin = false;
              post = false;
              start = 1;
            }
          }
        }
        busy = __shfl_sync(qmask, start, leader) != 0;
      }
    }
    if (__all_sync(0xffffffffu, exhausted && !busy)) break;
    // ---- one Newton step -------------------------------------------------------------------------------
    int do_step = 0;
    if (lead && busy && !post) {
      if (inner >= C.max_newton_iterations) post = true;
      else do_step = 1;
    }
    do_step = __shfl_sync(qmask, do_step, leader);
    if (do_step) {
      // single_iteration (single_iteration.rs:140-317) + solve_weighted_least_squares (least_square.rs:225-327)
      __syncwarp(qmask);
      double el[7];
      for (int j = 0; j < 7; ++j) el[j] = QS(kQsEl + j);
      double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, racc0 = 0.0, racc1 = 0.0, qsum = 0.0;
      unsigned long long active = 0;
      double *bv = &QS(kQsBv + 16 * q);
      for (unsigned base = 0; base < n_obs; base += 4) {
        const unsigned i = base + q;
        int use = 0;
        if (i < n_obs) {
          const unsigned long long gI = o0 + i;
          t_rra[i] = F[i].residual_ra; t_rdec[i] = F[i].residual_dec; t_chi[i] = F[i].chi;
          if (F[i].selection == 0) {
            double ra, dec, pr[6], pd[6];
            const V3 obs{__ldg(B.scorer + gI), __ldg(B.scorer + B.n_obs + gI), __ldg(B.scorer + 2 * B.n_obs + gI)};
            if (lsq_obs_and_partials(el, __ldg(B.mjd_tt + gI), obs, ra, dec, pr, pd)) {
              const double sra = __ldg(B.sigma_ra + gI), sdec = __ldg(B.sigma_dec + gI);
              const double xr = lsq_angular_diff(__ldg(B.ra + gI) - 0.0, ra);
              const double xd = (__ldg(B.dec + gI) - 0.0) - dec;
              const double ca = xr / sra, cd = xd / sdec;
              t_rra[i] = xr; t_rdec[i] = xd; t_chi[i] = sqrt(ca * ca + cd * cd);
              for (int j = 0; j < 6; ++j) { bv[j * S] = pr[j]; bv[(6 + j) * S] = pd[j]; }
              bv[12 * S] = xr; bv[13 * S] = xd; bv[14 * S] = 1.0 / (sra * sra); bv[15 * S] = 1.0 / (sdec * sdec);
              use = 1;
            }
          }
        }
        __syncwarp(qmask);
        for (int o = 0; o < 4; ++o) {  // the four observations of the round, in observation order
          if (!__shfl_sync(qmask, use, leader + o)) continue;
          const double *w = &QS(kQsBv + 16 * o);
          const double wr = w[14 * S], wd = w[15 * S], wc = 0.0, xr = w[12 * S], xd = w[13 * S];
#pragma unroll
          for (int m = 0; m < 9; ++m) {
            const int e = (int)q + 4 * m, j = e % 6, k = e / 6;  // OFB_M6(nm, j, k) = nm[6 k + j]
            const double prj = w[j * S], prk = w[k * S], pdj = w[(6 + j) * S], pdk = w[(6 + k) * S];
            acc[m] += prj * wr * prk + pdj * wd * pdk + wc * (pdj * prk + prj * pdk);
          }
          {
            const double prj = w[q * S], pdj = w[(6 + q) * S];
            racc0 += (prj * wr + pdj * wc) * xr + (prj * wc + pdj * wd) * xd;
          }
          if (q < 2) {
            const double prj = w[(q + 4) * S], pdj = w[(10 + q) * S];
            racc1 += (prj * wr + pdj * wc) * xr + (prj * wc + pdj * wd) * xd;
          }
          if (lead) { ++active; qsum += wr * xr * xr + wd * xd * xd + 2.0 * wc * xr * xd; }
        }
        __syncwarp(qmask);
      }
#pragma unroll
      for (int m = 0; m < 9; ++m) QS(kQsNm + (int)q + 4 * m) = acc[m];
      QS(kQsRhs + q) = racc0;
      if (q < 2) QS(kQsRhs + q + 4) = racc1;
      __syncwarp(qmask);
      int advance = 0;
      double *nm = &QS(kQsNm), *cov = &QS(kQsCov), *work = &QS(kQsWork);
      // invert_normal_matrix (least_square.rs:329-342): Cholesky on the leader, then the six columns of the
      // inverse on the four lanes (column c on lane c mod 4); Householder QR on the leader if not positive definite
      if (lead)
        for (int j = 0; j < 6; ++j)
          if (!C.free_el[j]) {
            for (int k = 0; k < 6; ++k) { OFB_MS(nm, j, k) = 0.0; OFB_MS(nm, k, j) = 0.0; }
            OFB_MS(nm, j, j) = 1.0;
          }
      __syncwarp(qmask);
      for (int i = (int)q; i < 36; i += 4) work[i] = nm[i];
      __syncwarp(qmask);
      const int chol = lsq_cholesky6_quad(work, q, qmask) ? 1 : 0;
      if (chol) {
        lsq_cholesky6_inverse_column<S>(work, cov, (int)q);
        if (q < 2) lsq_cholesky6_inverse_column<S>(work, cov, (int)q + 4);
      }
      __syncwarp(qmask);
      if (lead) {
        ++inner;
        ++total_it;
        double rhs[6];
        for (int j = 0; j < 6; ++j) rhs[j] = C.free_el[j] ? QS(kQsRhs + j) : 0.0;
        const unsigned long long nmeas = 2 * active;
        bool inv_ok = chol != 0;
        if (!inv_ok) {
          for (int i = 0; i < 36; ++i) work[i] = nm[i];
          inv_ok = lsq_qr6_inverse<S>(work, cov);
          if (!inv_ok)
            for (int i = 0; i < 36; ++i) cov[i] = 0.0;
        }
        double dx[6] = {0, 0, 0, 0, 0, 0};
        if (inv_ok) lsq_gemv6<S>(cov, rhs, dx);
        for (int j = 0; j < 6; ++j)
          if (!C.free_el[j]) dx[j] = 0.0;
        const double new_rms = nmeas > 0 ? sqrt(qsum / (double)nmeas) : 0.0;
        double cdx[6];
        lsq_gemv6<S>(nm, dx, cdx);
        const double cnorm = sqrt(lsq_dot6(dx, cdx));
        double corrected[6];
        for (int j = 0; j < 6; ++j) corrected[j] = C.free_el[j] ? el[1 + j] + dx[j] : el[1 + j];
        if (!inv_ok) { fail_code = OUTFIT_ST_LSQ_INVERSION; post = true; }
        else if (lsq_is_bizarre(corrected, C)) { fail_code = OUTFIT_ST_LSQ_BIZARRE; post = true; }
        else if (prev_rms < kMax && new_rms / prev_rms >= C.rms_divergence_ratio) { fail_code = OUTFIT_ST_LSQ_DIVERGED; post = true; }
        else {
          const bool stagnated = prev_rms < kMax && new_rms / prev_rms >= C.rms_stagnation_ratio;
          bool stop = false;
          if (stagnated) {
            if (++stagnation >= C.max_stagnation_iterations) stop = true;
          } else {
            stagnation = 0;
          }
          if (stop) {
            post = true;
          } else {  // advance the state
            for (int j = 0; j < 7; ++j) QS(kQsElLin + j) = el[j];
            have_lin = true;
            last_rms = new_rms;
            last_nmeas = nmeas;
            for (int j = 0; j < 6; ++j) QS(kQsEl + 1 + j) = corrected[j];
            advance = 1;
            prev_rms = new_rms;
            if (cnorm < C.convergence_threshold) { converged = true; post = true; }
          }
        }
      }
      if (__shfl_sync(qmask, advance, leader)) {
        for (unsigned i = q; i < n_obs; i += 4) { F[i].residual_ra = t_rra[i]; F[i].residual_dec = t_rdec[i]; F[i].chi = t_chi[i]; }
        for (int i = (int)q; i < 36; i += 4) {
          res->normal_matrix[i] = QS(kQsNm + i);  // unscaled until the trajectory finishes
          QS(kQsLastCov + i) = QS(kQsCov + i);
        }
      }
      __syncwarp(qmask);
    }
    // ---- the inner loop has ended (diff_cor.rs:400-428) --------------------------------------------------
    int do_reject = 0;
    bool finish = false;
    if (lead && busy && post) {
      finish = fail_code != 0 || !C.enable_outlier_rejection ||
               (outer == 0 && last_rms < C.convergence_before_rejection_threshold) || !converged || !have_lin;
      if (!finish) do_reject = 1;
    }
    do_reject = __shfl_sync(qmask, do_reject, leader);
    if (do_reject) {
      // update_observation_selection (outlier_rejection.rs:118-235) at el_lin
      __syncwarp(qmask);
      double el_lin[7];
      for (int j = 0; j < 7; ++j) el_lin[j] = QS(kQsElLin + j);
      const double *last_cov = &QS(kQsLastCov);
      unsigned changes = 0;
      for (unsigned i = q; i < n_obs; i += 4) {
        const unsigned long long gI = o0 + i;
        const int sel = F[i].selection;
        if (sel == 2) continue;
        double pr[6] = {0, 0, 0, 0, 0, 0}, pd[6] = {0, 0, 0, 0, 0, 0};
        double wr = 1.0, wd = 1.0;
        const double sra = __ldg(B.sigma_ra + gI), sdec = __ldg(B.sigma_dec + gI);
        if (sel == 0) {
          double ra, dec;
          const V3 obs{__ldg(B.scorer + gI), __ldg(B.scorer + B.n_obs + gI), __ldg(B.scorer + 2 * B.n_obs + gI)};
          if (lsq_obs_and_partials(el_lin, __ldg(B.mjd_tt + gI), obs, ra, dec, pr, pd)) {
            wr = 1.0 / (sra * sra); wd = 1.0 / (sdec * sdec);
          } else {
            for (int j = 0; j < 6; ++j) { pr[j] = 0.0; pd[j] = 0.0; }
          }
        }
        const double var_ra = sra * sra, var_dec = sdec * sdec;
        const double cov_cross = -sra * sdec * 0.0 / (wr * wd);
        double gga[6], ggd[6];
        lsq_gemv6<S>(last_cov, pr, gga);
        lsq_gemv6<S>(last_cov, pd, ggd);
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
      __syncwarp(qmask);
      changes += __shfl_xor_sync(qmask, changes, 1);
      changes += __shfl_xor_sync(qmask, changes, 2);
      if (lead) {
        if (changes == 0) finish = true;
        else if (++outer > C.max_outlier_rejection_passes) finish = true;
        else { inner = 0; prev_rms = kMax; stagnation = 0; converged = false; have_lin = false; post = false; }
      }
    }
    const int fin = __shfl_sync(qmask, (lead && busy && post && finish) ? 1 : 0, leader);
    if (fin) {
      const int failed = __shfl_sync(qmask, fail_code, leader);
      if (failed)  // Err(_) => Ok(initial_orbit) (mod.rs:113)
        for (unsigned i = q; i < n_obs; i += 4) { F[i].residual_ra = 0.0; F[i].residual_dec = 0.0; F[i].chi = 0.0; F[i].selection = 0; }
      double mu2 = 1.0;
      if (lead) {
        res->status = OUTFIT_ST_OK;
        res->total_newton_iterations = total_it;
        if (fail_code) {
          res->kind = OUTFIT_LSQ_IOD_FALLBACK;
          res->fallback_cause = fail_code;
          res->epoch = iod[tr].epoch;
          for (int j = 0; j < 6; ++j) res->elem[j] = iod[tr].elem[j];
          res->normalised_rms = iod[tr].rms;
        } else {  // rescale_covariance (least_square.rs:371-394)
          double mu = 1.0;
          if (num_free < last_nmeas) {
            const double factor = sqrt((double)last_nmeas / (double)(last_nmeas - num_free));
            mu = last_rms > 1.0 ? last_rms * factor : factor;
          }
          mu2 = mu * mu;
          res->kind = OUTFIT_LSQ_CORRECTED;
          res->epoch = QS(kQsEl);
          for (int j = 0; j < 6; ++j) res->elem[j] = QS(kQsEl + 1 + j);
          res->normalised_rms = last_rms;
          res->num_measurements = last_nmeas;
        }
      }
      mu2 = __shfl_sync(qmask, mu2, leader);
      if (failed) {
        for (int i = (int)q; i < 36; i += 4) res->normal_matrix[i] = 0.0;
      } else {
#pragma unroll 1
        for (int i = (int)q; i < 36; i += 4) {
          res->covariance[i] = QS(kQsLastCov + i) * mu2;
          res->normal_matrix[i] = res->normal_matrix[i] / mu2;
        }
#pragma unroll 1
        for (int j = (int)q; j < 6; j += 4) res->sigma[j] = sqrt(QS(kQsLastCov + 7 * j) * mu2);
      }
      busy = false;
      __syncwarp(qmask);
    }
  }
#undef QS
}

}  // namespace ofb
