// k_iod.cuh -- the five kernels of the full-IOD phase pipeline (FitIOD::fit_full_iod, obs_dataset_api.rs:145-296;
// estimate_best_orbit, trajectory.rs:429-545).  Launched by outfit_b200.cu: launch_iod.
#pragma once
#include "../../include/outfit_b200.h"
#include "dev_iod.cuh"
#include "dev_correct.cuh"

using namespace ofb;

// =================================================================================================
// full-IOD pipeline.  Mapping: one warp per trajectory for the two trajectory-level steps (triplet
// selection, best-orbit fold with warp-shuffle argmin) and one lane per candidate = (triplet,
// noise realization) for the three numeric phases.  The phases are separate launches over flat
// candidate arrays because the fused single kernel was instruction-cache bound (ncu r01a/r01b:
// 57-73 % stall_no_inst with ~100 KB of hot SASS); per phase the hot loop is a few KB, every warp
// of an SM runs the same loop, and register use / occupancy is set per phase.
//   P0 triplets_kernel   warp / trajectory   best-K triplets            -> trip[T][K], ktraj[T]
//   P1 roots_kernel      lane / candidate    geometry, degree-8 poly, Aberth -> roots, code
//   P2 correct_kernel    lane / candidate    accept root, f-g correction -> state (r, v, epoch)
//   P3 score_kernel      lane / candidate    elements, equinoctial, arc RMS sum -> kind, sum, n_arc
//   P4 select_kernel     warp / trajectory   order-preserving fold + result record
// Candidate id = (t * K + r) * M + m  (K = max_triplets, M = 1 + n_noise_realizations).
// =================================================================================================
struct IodBatchDev {
  unsigned long long n_traj;
  unsigned long long n_obs;
  const unsigned long long *traj_offset;
  const double *mjd_tt, *ra, *dec, *sigma_ra, *sigma_dec;
  const double *helio;   // [3][n_obs]
  const double *scorer;  // [3][n_obs]
  const double *noise_z; // [n_traj][max_triplets][n_noise][6] or null
  const int *obs_status; // [n_obs] 0 | OUTFIT_ST_EPHEM_OUT_OF_RANGE (observer kernels)
};

struct IodScratch {
  unsigned *trip;        // [T][K] packed (i<<20 | j<<10 | k), ascending weight
  unsigned *ktraj;       // [T] number of triplets found
  int *code;             // [C] P1: 0 ok | OUTFIT_ST_* ; P3 overwrites with the score kind
  unsigned char *nroots; // [C]
  double *roots;         // [8][C] admissible roots in solver order
  int *state_kind;       // [C] 0 none, 1 PrelimOrbit, 2 CorrectedOrbit
  double *state;         // [7][C] r(t2) xyz, v(t2) xyz, epoch
  int *score_kind;       // [C] 0 gauss error (code in score_code), 1 abort, 2 break, 3 sum
  int *score_code;       // [C]
  double *score_sum;     // [C]
  unsigned *score_narc;  // [C]
  unsigned long long n_cand;
};

constexpr int kWarpsPerBlock = 4;
constexpr int kCandThreads = 128;

__device__ __forceinline__ void flush_work(const Work &w, unsigned long long *__restrict__ work_counters) {
  const unsigned *wp = reinterpret_cast<const unsigned *>(&w);
  const unsigned lane = threadIdx.x & 31u;
#pragma unroll
  for (int q = 0; q < (int)(sizeof(Work) / sizeof(unsigned)); ++q) {
    unsigned long long v = wp[q];
    if (__any_sync(0xffffffffu, v != 0)) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0 && work_counters) atomicAdd(&work_counters[q], v);
    }
  }
}

// ---- P0: best-K triplets, one warp per trajectory --------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
triplets_kernel(IodBatchDev B, IodDevParams P, IodScratch S, unsigned n_obs_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const size_t per_warp = ((size_t)n_obs_cap * sizeof(double) + (size_t)P.max_triplets * (8 + 4) + 15) & ~(size_t)15;
  TrajSmem sm;
  sm.t = reinterpret_cast<double *>(smem_raw + warp * per_warp);
  sm.heap_w = sm.t + n_obs_cap;
  sm.heap_x = reinterpret_cast<unsigned *>(sm.heap_w + P.max_triplets);
  const unsigned long long tr = (unsigned long long)blockIdx.x * kWarpsPerBlock + warp;
  if (tr >= B.n_traj) return;
  const unsigned long long o0 = B.traj_offset[tr];
  const unsigned n_obs = (unsigned)(B.traj_offset[tr + 1] - o0);
  for (unsigned i = lane; i < n_obs; i += 32) sm.t[i] = B.mjd_tt[o0 + i];
  __syncwarp();
  const unsigned K = select_triplets(sm, n_obs, P, lane);
  for (unsigned a = lane; a < K; a += 32) S.trip[tr * P.max_triplets + a] = sm.heap_x[a];
  if (lane == 0) S.ktraj[tr] = K;
}

// P0, one thread per trajectory (dev_iod.cuh: select_triplets_thread); dynamic shared memory =
// blockDim.x * max_triplets * 12 bytes
__global__ void __launch_bounds__(128)
triplets_thread_kernel(IodBatchDev B, IodDevParams P, IodScratch S) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *hw = reinterpret_cast<double *>(smem_raw);
  unsigned *hx = reinterpret_cast<unsigned *>(hw + (size_t)P.max_triplets * blockDim.x);
  const unsigned long long tr = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tr >= B.n_traj) return;
  const unsigned long long o0 = B.traj_offset[tr];
  const unsigned n_obs = (unsigned)(B.traj_offset[tr + 1] - o0);
  const HeapCol h{hw + threadIdx.x, hx + threadIdx.x, blockDim.x};
  const unsigned K = select_triplets_thread(h, B.mjd_tt + o0, n_obs, P);
  for (unsigned a = 0; a < K; ++a) S.trip[tr * P.max_triplets + a] = h.X(a);
  S.ktraj[tr] = K;
}

// candidate id -> (trajectory, triplet rank, realization); false when the slot is unused
__device__ __forceinline__ bool decode_candidate(unsigned long long cid, const IodDevParams &P, const IodScratch &S,
                                                 unsigned long long &tr, unsigned &r, unsigned &m) {
  const unsigned M = P.n_noise + 1;
  const unsigned long long tk = cid / M;
  m = (unsigned)(cid - tk * M);
  tr = tk / P.max_triplets;
  r = (unsigned)(tk - tr * P.max_triplets);
  return r < S.ktraj[tr];
}

// observations of the triplet (+ the host-drawn noise of this realization, gauss.rs:323-387)
__device__ __forceinline__ void load_triplet(const IodBatchDev &B, const IodDevParams &P, const IodScratch &S,
                                             unsigned long long tr, unsigned r, unsigned m, Triplet &g, unsigned (&idx)[3]) {
  const unsigned packed = S.trip[tr * P.max_triplets + r];
  idx[0] = packed >> 20; idx[1] = (packed >> 10) & 1023u; idx[2] = packed & 1023u;
  const unsigned long long o0 = B.traj_offset[tr];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const unsigned long long gI = o0 + idx[c];
    g.t[c] = __ldg(B.mjd_tt + gI);
    g.ra[c] = __ldg(B.ra + gI);
    g.dec[c] = __ldg(B.dec + gI);
    g.R[c] = V3{__ldg(B.helio + gI), __ldg(B.helio + B.n_obs + gI), __ldg(B.helio + 2 * B.n_obs + gI)};
  }
  if (m > 0) {
    const double *z = B.noise_z + (((size_t)tr * P.max_triplets + r) * P.n_noise + (m - 1)) * 6;
    const double2 z01 = __ldg(reinterpret_cast<const double2 *>(z));
    const double2 z23 = __ldg(reinterpret_cast<const double2 *>(z) + 1);
    const double2 z45 = __ldg(reinterpret_cast<const double2 *>(z) + 2);
    const unsigned long long g0 = o0 + idx[0], g1 = o0 + idx[1], g2 = o0 + idx[2];
    g.ra[0] = g.ra[0] + z01.x * (__ldg(B.sigma_ra + g0) * P.noise_scale);
    g.ra[1] = g.ra[1] + z01.y * (__ldg(B.sigma_ra + g1) * P.noise_scale);
    g.ra[2] = g.ra[2] + z23.x * (__ldg(B.sigma_ra + g2) * P.noise_scale);
    g.dec[0] = g.dec[0] + z23.y * (__ldg(B.sigma_dec + g0) * P.noise_scale);
    g.dec[1] = g.dec[1] + z45.x * (__ldg(B.sigma_dec + g1) * P.noise_scale);
    g.dec[2] = g.dec[2] + z45.y * (__ldg(B.sigma_dec + g2) * P.noise_scale);
  }
}

// ---- P1: geometry + polynomial + Aberth ---------------------------------------------------------
#ifndef OUTFIT_ROOTS_BPS
#define OUTFIT_ROOTS_BPS 4
#endif
__global__ void __launch_bounds__(kCandThreads, OUTFIT_ROOTS_BPS)
roots_kernel(IodBatchDev B, IodDevParams P, IodScratch S, unsigned long long *__restrict__ work_counters) {
  extern __shared__ __align__(16) double roots_sm[];  // [32][kCandThreads]: iterates + Aberth sums
  volatile double *zsm = roots_sm + threadIdx.x;
  const unsigned long long cid = (unsigned long long)blockIdx.x * kCandThreads + threadIdx.x;
  Work w;
  memset(&w, 0, sizeof w);
  unsigned long long tr;
  unsigned r, m;
  if (cid < S.n_cand && decode_candidate(cid, P, S, tr, r, m)) {
    ++w.candidates;
    ++w.gauss_solves;
    Triplet g;
    unsigned idx[3];
    load_triplet(B, P, S, tr, r, m, g, idx);
    GaussGeom gm;
    int code = 0;
    unsigned n = 0;
    double c0, c3, c6;
    if (!gauss_geometry(g, gm)) code = OUTFIT_ST_SINGULAR_DIRECTION_MATRIX;
    else if (!gauss_polynomial(g, gm, c0, c3, c6)) code = OUTFIT_ST_GAUSS_NO_ROOTS;
    else if (aberth8(c0, c3, c6, P.aberth_max_iter, P.aberth_eps, zsm, kCandThreads, w) == 2) code = OUTFIT_ST_POLY_ROOT_FAILED;
    else {
      // visit_real_positive_roots + plausibility window (gauss.rs:975-981, 1148), solver order kept
#pragma unroll 1
      for (int k = 0; k < 8; ++k) {
        const double re = zsm[k * kCandThreads];
        if (re > 0.0 && fabs(zsm[(8 + k) * kCandThreads]) < P.root_imag_eps && re >= P.r2_min_au && re <= P.r2_max_au) {
          S.roots[(size_t)n * S.n_cand + cid] = re;
          ++n;
        }
      }
      if (n == 0) code = OUTFIT_ST_GAUSS_NO_ROOTS;
    }
    S.code[cid] = code;
    S.nroots[cid] = (unsigned char)n;
  }
  flush_work(w, work_counters);
}

// ---- P2: roots -> accepted state, f-g correction ---------------------------------------------------
// One lane per candidate, register-resident f-g loop (dev_correct.cuh).  The triplet geometry is
// rebuilt here (6 sincos + the cofactor inverse: ~3 % of this phase) instead of being carried from P1
// through HBM (144 B per candidate).
#ifndef OUTFIT_CORRECT_BPS
#define OUTFIT_CORRECT_BPS 5  // 96 registers, 20 warps per SM: 42.3 ms against 44.5 at 4 blocks / 126 registers (round 2, r2a)
#endif
#ifdef OUTFIT_DEBUG_FGHIST
__device__ unsigned long long g_fghist[3][128];
#endif
template <bool COUNT>
__global__ void __launch_bounds__(kCorrectThreads, OUTFIT_CORRECT_BPS)
correct_kernel(IodBatchDev B, IodDevParams P, IodScratch S, unsigned long long *__restrict__ work_counters) {
  extern __shared__ __align__(16) double geo_sm[];
  double *my = geo_sm + threadIdx.x;
  const unsigned long long cid = (unsigned long long)blockIdx.x * kCorrectThreads + threadIdx.x;
  WorkC w;
  w.roots_accepted = 0; w.fg_iterations = 0; w.kepler_solves = 0; w.newton_steps = 0; w.sfunct_terms = 0; w.fg_skipped = 0;
#ifdef OUTFIT_DEBUG_STRAGGLERS
  const long long dbg_t0 = clock64();
#endif
  unsigned long long tr;
  unsigned r, m;
  if (cid < S.n_cand && decode_candidate(cid, P, S, tr, r, m)) {
    int kind = 0;
    const unsigned n = S.code[cid] == 0 ? S.nroots[cid] : 0u;
    if (n > 0) {
      // ---- triplet + noise -> unit vectors, heliocentric observer positions (shared memory) ----
      const unsigned packed = S.trip[tr * P.max_triplets + r];
      const unsigned long long o0 = B.traj_offset[tr];
      const double *z = m > 0 ? B.noise_z + (((size_t)tr * P.max_triplets + r) * P.n_noise + (m - 1)) * 6 : nullptr;
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        const unsigned long long gI = o0 + ((packed >> (20 - 10 * c)) & 1023u);
        double ra = __ldg(B.ra + gI), dec = __ldg(B.dec + gI);
        if (m > 0) {
          ra = ra + __ldg(z + c) * (__ldg(B.sigma_ra + gI) * P.noise_scale);
          dec = dec + __ldg(z + 3 + c) * (__ldg(B.sigma_dec + gI) * P.noise_scale);
        }
        double sr, cr, sd, cd;
        sincos(ra, &sr, &cr);
        sincos(dec, &sd, &cd);
        my[(SL_S0 + 3 * c + 0) * kCorrectThreads] = cr * cd;
        my[(SL_S0 + 3 * c + 1) * kCorrectThreads] = sr * cd;
        my[(SL_S0 + 3 * c + 2) * kCorrectThreads] = sd;
        my[(SL_R0 + 3 * c + 0) * kCorrectThreads] = __ldg(B.helio + gI);
        my[(SL_R0 + 3 * c + 1) * kCorrectThreads] = __ldg(B.helio + B.n_obs + gI);
        my[(SL_R0 + 3 * c + 2) * kCorrectThreads] = __ldg(B.helio + 2 * B.n_obs + gI);
        my[(SL_T0 + c) * kCorrectThreads] = __ldg(B.mjd_tt + gI);
      }
      const GeoSm G{my};
      {
        // gauss_prelim (gauss.rs:464-549): tau, a, b, cofactor inverse (rows of S^-1)
        const V3 S0 = G.v3(SL_S0), S1 = G.v3(SL_S1), S2 = G.v3(SL_S2);
        const double m11 = S0.x, m12 = S1.x, m13 = S2.x, m21 = S0.y, m22 = S1.y, m23 = S2.y, m31 = S0.z, m32 = S1.z, m33 = S2.z;
        const double mi1 = m22 * m33 - m32 * m23, mi2 = m21 * m33 - m31 * m23, mi3 = m21 * m32 - m31 * m22;
        const double det = m11 * mi1 - m12 * mi2 + m13 * mi3;  // != 0: P1 accepted this candidate
        const double num[9] = {mi1, m13 * m32 - m33 * m12, m12 * m23 - m22 * m13, -mi2, m11 * m33 - m31 * m13,
                               m13 * m21 - m23 * m11, mi3, m12 * m31 - m32 * m11, m11 * m22 - m21 * m12};
        // det != 0 and P1 found admissible roots with this very matrix: a |det| small enough to break the
        // reciprocal form (< 1e-290) would have sent the roots out of the plausibility window
        const double yd = 1.0 / det;
#pragma unroll
        for (int q = 0; q < 9; ++q) my[(SL_I0 + q) * kCorrectThreads] = div_mk(num[q], det, yd);
      }
      unsigned n_solutions = 0;
#pragma unroll 1
      for (unsigned k = 0; k < n; ++k) {
        V3 p1, vel;
        double ep;
        MidC mid;
        if (!accept_root_fast<COUNT>(G, P, S.roots[(size_t)k * S.n_cand + cid], p1, vel, ep, mid, w)) continue;
        ++n_solutions;
        // prelim_orbit (gauss.rs:1238-1247): first CorrectedOrbit in discovery order, else first pushed
        const bool first = kind == 0;
        if (first) {
          kind = 1;
          S.state[0 * S.n_cand + cid] = p1.x; S.state[1 * S.n_cand + cid] = p1.y; S.state[2 * S.n_cand + cid] = p1.z;
          S.state[3 * S.n_cand + cid] = vel.x; S.state[4 * S.n_cand + cid] = vel.y; S.state[5 * S.n_cand + cid] = vel.z;
          S.state[6 * S.n_cand + cid] = ep;
        }
        if (fg_correction_fast<COUNT>(G, P, p1, vel, mid, ep, w)) {
          kind = 2;
          S.state[0 * S.n_cand + cid] = p1.x; S.state[1 * S.n_cand + cid] = p1.y; S.state[2 * S.n_cand + cid] = p1.z;
          S.state[3 * S.n_cand + cid] = vel.x; S.state[4 * S.n_cand + cid] = vel.y; S.state[5 * S.n_cand + cid] = vel.z;
          S.state[6 * S.n_cand + cid] = ep;
          break;
        }
        if (n_solutions >= P.max_tested_solutions) break;
      }
    }
    S.state_kind[cid] = kind;
#ifdef OUTFIT_DEBUG_STRAGGLERS
    // slowest thread of the launch ((cycles >> 8) << 26 | cid) and the total thread time (cycles >> 8)
    const unsigned long long dc = (unsigned long long)(clock64() - dbg_t0);
    atomicMax(work_counters + 20, ((dc >> 8) << 26) | (cid & 0x3ffffffull));
    atomicAdd(work_counters + 24, dc >> 8);
#endif
  }
#ifdef OUTFIT_DEBUG_FGHIST
  if (COUNT) {  // distribution of the executed Kepler Newton steps per candidate, and of the maximum over each warp
    const unsigned mine = w.newton_steps;
    const unsigned mx = __reduce_max_sync(0xffffffffu, mine), sm = __reduce_add_sync(0xffffffffu, mine);
    atomicAdd(&g_fghist[0][min(mine / 8u, 127u)], 1ull);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&g_fghist[1][min(mx / 8u, 127u)], 1ull);
      atomicAdd(&g_fghist[2][0], (unsigned long long)sm); atomicAdd(&g_fghist[2][1], (unsigned long long)mx); atomicAdd(&g_fghist[2][2], 1ull);
    }
  }
#endif
  if (COUNT) {
    Work wk;
    memset(&wk, 0, sizeof wk);
    wk.roots_accepted = w.roots_accepted; wk.fg_iterations = w.fg_iterations; wk.kepler_solves = w.kepler_solves;
    wk.newton_steps = w.newton_steps; wk.sfunct_terms = w.sfunct_terms; wk.fg_skipped = w.fg_skipped;
    flush_work(wk, work_counters);
  }
}

__device__ __forceinline__ void state_to_orbit(const IodScratch &S, unsigned long long cid, int state_kind, Orbit &orb) {
  const V3 rr = V3{S.state[0 * S.n_cand + cid], S.state[1 * S.n_cand + cid], S.state[2 * S.n_cand + cid]};
  const V3 vv = V3{S.state[3 * S.n_cand + cid], S.state[4 * S.n_cand + cid], S.state[5 * S.n_cand + cid]};
  // build_result (gauss.rs:1063): rotate to ecliptic J2000, state -> elements
  ccek1(equ_to_ecl(rr), equ_to_ecl(vv), S.state[6 * S.n_cand + cid], orb);
  orb.corrected = state_kind == 2 ? 1 : 0;
}

// ---- P3: elements -> equinoctial -> arc RMS sum --------------------------------------------------
#ifndef OUTFIT_SCORE_BPS
#define OUTFIT_SCORE_BPS 7  // 72 registers: 16.75 ms per 100 k trajectories against 17.04 at 6 blocks and 16.94 at 8 (r2e)
#endif
template <bool COUNT>
__global__ void __launch_bounds__(kCandThreads, OUTFIT_SCORE_BPS)
score_kernel(IodBatchDev B, IodDevParams P, IodScratch S, unsigned long long *__restrict__ work_counters) {
  const unsigned long long cid = (unsigned long long)blockIdx.x * kCandThreads + threadIdx.x;
  Work w;
  memset(&w, 0, sizeof w);
  unsigned long long tr;
  unsigned r, m;
  if (cid < S.n_cand && decode_candidate(cid, P, S, tr, r, m)) {
    int kind, code = 0;
    double sum = 0.0;
    unsigned n_arc = 0;
    const int gcode = S.code[cid];
    const int sk = S.state_kind[cid];
    if (gcode != 0) { kind = 0; code = gcode; }
    else if (sk == 0) { kind = 0; code = OUTFIT_ST_GAUSS_NO_ROOTS; }
    else {
      Orbit orb;
      state_to_orbit(S, cid, sk, orb);
      Equinoctial eq;
      const int rq = to_equinoctial(orb, eq);
      if (rq != 0) { kind = 1; code = rq; }
      else {
        // select_rms_interval (trajectory.rs:294-350)
        const unsigned packed = S.trip[tr * P.max_triplets + r];
        const unsigned i0 = packed >> 20, i2 = packed & 1023u;
        const unsigned long long o0 = B.traj_offset[tr];
        const unsigned n_obs = (unsigned)(B.traj_offset[tr + 1] - o0);
        const double *T = B.mjd_tt + o0;
        const double t1 = __ldg(T + i0), t3 = __ldg(T + i2);
        double dtw = P.extf >= 0.0 ? (t3 - t1) * P.extf : 10.0 * (__ldg(T + n_obs - 1) - __ldg(T));
        if (P.dtmax >= 0.0) dtw = fmax(dtw, P.dtmax);
        unsigned is = 0, ie = n_obs - 1;
        for (int ii = (int)i0; ii >= 0; --ii) {
          if (t1 - __ldg(T + ii) > dtw) break;
          is = (unsigned)ii;
        }
        for (unsigned ii = i2; ii < n_obs; ++ii) {
          if (__ldg(T + ii) - t3 > dtw) break;
          ie = ii;
        }
        n_arc = ie - is + 1;
        const ScoreOrbit so = make_score_orbit(eq);
        kind = 3;
        if (!so.elliptic) {
          kind = 2;
        } else {
#pragma unroll 1
          for (unsigned ii = is; ii <= ie; ++ii) {
            const unsigned long long gI = o0 + ii;
            const double dec_o = __ldg(B.dec + gI);
            double v;
            if (!ephemeris_error<COUNT>(so, __ldg(T + ii), __ldg(B.ra + gI), dec_o, cos_angle(dec_o), __ldg(B.sigma_ra + gI),
                                 __ldg(B.sigma_dec + gI),
                                 V3{__ldg(B.scorer + gI), __ldg(B.scorer + B.n_obs + gI), __ldg(B.scorer + 2 * B.n_obs + gI)},
                                 v, w)) {
              kind = 2;
              break;
            }
            const double ns = sum + v;
            if (ns >= INFINITY) { kind = 2; break; }
            sum = ns;
          }
        }
      }
    }
    S.score_kind[cid] = kind;
    S.score_code[cid] = code;
    S.score_sum[cid] = sum;
    S.score_narc[cid] = n_arc;
  }
  if (COUNT) flush_work(w, work_counters);
}

// ---- P4: per-trajectory fold, one warp per trajectory (trajectory.rs:429-545) ----------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
select_kernel(IodBatchDev B, IodDevParams P, IodScratch S, OutfitIodResult *__restrict__ out) {
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned long long tr = (unsigned long long)blockIdx.x * kWarpsPerBlock + warp;
  if (tr >= B.n_traj) return;
  const unsigned M = P.n_noise + 1;
  const unsigned K = S.ktraj[tr];
  const unsigned long long o0 = B.traj_offset[tr];
  const unsigned n_obs = (unsigned)(B.traj_offset[tr + 1] - o0);
  OutfitIodResult res;
  memset(&res, 0, sizeof res);
  res.rms = NAN;
  {
    // An observation epoch outside the loaded ephemeris: the reference panics ("Time outside ephemeris
    // range", horizon_data.rs:722); here the trajectory carries the error as a value and the others go on.
    int bad = 0;
    for (unsigned i = lane; i < n_obs; i += 32) bad |= B.obs_status[o0 + i];
    if (__any_sync(0xffffffffu, bad != 0)) {
      if (lane == 0) {
        res.status = OUTFIT_ST_EPHEM_OUT_OF_RANGE;
        out[tr] = res;
      }
      return;
    }
  }
  if (K == 0) {
    if (lane == 0) {
      res.status = OUTFIT_ST_NO_FEASIBLE_TRIPLETS;
      res.span = n_obs == 0 ? 0.0 : B.mjd_tt[o0 + n_obs - 1] - B.mjd_tt[o0];
      out[tr] = res;
    }
    return;
  }
  const unsigned n_cand = K * M;
  const unsigned long long cbase0 = tr * (unsigned long long)P.max_triplets * M;
  double best_rms = INFINITY;
  unsigned best_c = 0xffffffffu;
  int abort_code = 0;
  unsigned abort_c = 0xffffffffu;
  int last_code = 0;
  double last_val = 0.0;
  for (unsigned cbase = 0; cbase < n_cand; cbase += 32) {
    const unsigned c = cbase + lane;
    int kind = -1, code = 0;
    double sum = 0.0;
    unsigned n_arc = 0;
    if (c < n_cand) {
      kind = S.score_kind[cbase0 + c];
      code = S.score_code[cbase0 + c];
      sum = S.score_sum[cbase0 + c];
      n_arc = S.score_narc[cbase0 + c];
    }
    // (a) the first candidate whose conversion to equinoctial fails aborts the trajectory (`?`)
    {
      const unsigned ab = __ballot_sync(0xffffffffu, kind == 1);
      if (ab != 0 && abort_c == 0xffffffffu) {
        const int src = __ffs(ab) - 1;
        abort_c = cbase + src;
        abort_code = __shfl_sync(0xffffffffu, code, src);
      }
    }
    // (b) running best with the reference's pruning rule: a candidate replaces the best iff its
    //     full sum stays below best^2 * 2N (never pruned) and sqrt(sum / 2N) < best (strict)
    {
      const double denom = 2.0 * (double)n_arc;
      const double rms_c = sqrt(sum / denom);
      unsigned from = 0;
      for (;;) {
        const double cutoff = isfinite(best_rms) ? best_rms * best_rms * denom : INFINITY;
        const bool acc = kind == 3 && lane >= from && !(sum >= cutoff) && isfinite(rms_c) && rms_c < best_rms;
        const unsigned bal = __ballot_sync(0xffffffffu, acc);
        if (bal == 0) break;
        const int src = __ffs(bal) - 1;
        best_rms = __shfl_sync(0xffffffffu, rms_c, src);
        best_c = cbase + src;
        from = src + 1;
        if (from >= 32) break;
      }
    }
    // (c) error of the LAST candidate in evaluation order (used only when nothing succeeded, in
    //     which case the running best stayed +inf for every candidate)
    if (c == n_cand - 1) {
      if (kind == 0) { last_code = code; last_val = 0.0; }
      else if (kind == 2) { last_code = OUTFIT_ST_NON_FINITE_SCORE; last_val = INFINITY; }
      else if (kind == 3) { last_code = OUTFIT_ST_NON_FINITE_SCORE; last_val = sqrt(sum / (2.0 * (double)n_arc)); }
    }
  }
  const unsigned last_lane = (n_cand - 1) & 31u;
  const int l_code = __shfl_sync(0xffffffffu, last_code, last_lane);
  const double l_val = __shfl_sync(0xffffffffu, last_val, last_lane);
  if (lane != 0) return;
  if (abort_c != 0xffffffffu) {
    res.status = abort_code;
    res.attempts = abort_c + 1;
  } else if (best_c != 0xffffffffu) {
    const unsigned r = best_c / M;
    const unsigned long long cid = cbase0 + best_c;
    Orbit orb;
    state_to_orbit(S, cid, S.state_kind[cid], orb);
    res.status = OUTFIT_ST_OK;
    res.attempts = n_cand;
    res.corrected = orb.corrected;
    res.element_kind = orb.kind;
    res.epoch = orb.epoch;
#pragma unroll
    for (int q = 0; q < 6; ++q) res.elem[q] = orb.e[q];
    res.rms = best_rms;
    const unsigned packed = S.trip[tr * P.max_triplets + r];
    res.triplet_idx[0] = packed >> 20;
    res.triplet_idx[1] = (packed >> 10) & 1023u;
    res.triplet_idx[2] = packed & 1023u;
    res.triplet_rank = r;
    res.realization = best_c - r * M;
  } else {
    res.status = OUTFIT_ST_NO_VIABLE_ORBIT;
    res.cause = l_code;
    res.cause_value = l_val;
    res.attempts = n_cand;
  }
  out[tr] = res;
}

