// dev_ephemeris.cuh -- two-body `Combined` ephemeris (apparent RA/Dec, distances, phase angle, solar
// elongation, radial velocity, angular rates) for many orbits x many epochs of one observer.
//
// Reference behaviour (/root/reference/src):
//   OrbitalElements::compute::<Combined>   ephemeris/mod.rs:189-292, request.rs:181-205
//   propagate / observer_pv                ephemeris/apparent_position.rs:135-160, 264-296
//   assemble_apparent_position             ephemeris/apparent_position.rs:315-357
//   compute_geometry                       ephemeris/geometry.rs:204-345
//   PropagatorKind::TwoBody                propagator/mod.rs:84-91, 128-135
//   propagate_twobody (equinoctial)        orbit_type/equinoctial_element.rs:326-348, 639-867
//   HorizonRecord::interpolate (velocity)  jpl_ephem/horizon/horizon_records.rs:204-298
//
// Mapping.  The observer state depends on the epoch only, yet the reference recomputes it (pvobs +
// six Chebyshev bodies) for every (orbit, epoch) entry: here `ephemeris_observer_kernel` evaluates it
// once per epoch into a [9][E] table, and `ephemeris_twobody_kernel` runs one thread per orbit, hoists
// everything of the propagation that does not depend on the epoch, and walks the epochs with the
// table staged in shared memory by a bulk asynchronous copy (cp.async.bulk + mbarrier: the TMA engine
// moves the tile while the threads convert their elements).  Outputs are plane-major
// [quantity][epoch][orbit], so the 9 stores of a warp are 9 contiguous 256-byte segments.
#pragma once
#include "dev_geometry.cuh"

namespace ofb {

// Chebyshev position AND velocity of one body (km, km/day), HorizonRecord::interpolate
__device__ __forceinline__ void cheb_posvel(const double *__restrict__ blk, unsigned off, unsigned nc, unsigned nsub,
                                            double tau, double block_days, V3 &pos, V3 &vel) {
  const double fs = floor(tau * (double)nsub);
  const double mx = (double)nsub - 1.0;
  const unsigned sub = (unsigned)(fs < mx ? fs : mx);
  const double *cf = blk + off + (size_t)sub * nc * 3;
  const double temp = (double)nsub * tau;
  const double tc = 2.0 * (rem_euclid(temp, 1.0) + (double)(long long)tau) - 1.0;
  const double twot = tc + tc;
  const double vfac = (2.0 * (double)nsub) / block_days;
  // T_0 = 1, T_1 = tc ; T'_0 = 0, T'_1 = 1, T'_2 = 4 tc, T'_i = 2tc T'_{i-1} + 2 T_{i-1} - T'_{i-2}
  double tm2 = 1.0, tm1 = tc;
  double dm2 = 1.0, dm1 = twot + twot;  // T'_1, T'_2
  double x = __ldg(cf) * 1.0, y = __ldg(cf + nc) * 1.0, z = __ldg(cf + 2 * nc) * 1.0;
  x += __ldg(cf + 1) * tc; y += __ldg(cf + nc + 1) * tc; z += __ldg(cf + 2 * nc + 1) * tc;
  double vx = 0.0 + __ldg(cf) * 0.0, vy = 0.0 + __ldg(cf + nc) * 0.0, vz = 0.0 + __ldg(cf + 2 * nc) * 0.0;
  vx += __ldg(cf + 1) * 1.0; vy += __ldg(cf + nc + 1) * 1.0; vz += __ldg(cf + 2 * nc + 1) * 1.0;
  for (unsigned i = 2; i < nc; ++i) {
    const double ti = twot * tm1 - tm2;
    double di;
    if (i == 2) di = dm1;
    else di = twot * dm1 + 2.0 * tm1 - dm2;
    const double cx = __ldg(cf + i), cy = __ldg(cf + nc + i), cz = __ldg(cf + 2 * nc + i);
    x += cx * ti; y += cy * ti; z += cz * ti;
    vx += cx * di; vy += cy * di; vz += cz * di;
    tm2 = tm1; tm1 = ti;
    if (i > 2) dm2 = dm1;
    dm1 = di;
  }
  pos = V3{x, y, z};
  vel = V3{vfac * vx, vfac * vy, vfac * vz};
}

// heliocentric Earth position + velocity (equatorial J2000, AU, AU/day): JPLEphem::earth_ephemeris(.., true)
__device__ __forceinline__ bool earth_posvel(const EphemDev &E, double et, V3 &pos, V3 &vel) {
  const double et_jd = 2400000.5 + trunc(et);
  if (et_jd < E.jd_start || et_jd > E.jd_end) return false;
  long long nr = (long long)floor((et_jd - E.jd_start) / E.block_days);
  if (fabs(et_jd - E.jd_end) < 1e-10) nr -= 1;
  if (nr < 0 || (size_t)nr >= E.n_blocks) return false;
  const double interval_start = (double)nr * E.block_days + E.jd_start;
  const double tau = ((et_jd - interval_start) + (et - trunc(et))) / E.block_days;
  const double *blk = E.cheb + (size_t)nr * E.block_stride;
  V3 pe, ve, pm, vm, ps, vs;
  cheb_posvel(blk, E.ipt[0][0], E.ipt[0][1], E.ipt[0][2], tau, E.block_days, pe, ve);
  cheb_posvel(blk, E.ipt[1][0], E.ipt[1][1], E.ipt[1][2], tau, E.block_days, pm, vm);
  cheb_posvel(blk, E.ipt[2][0], E.ipt[2][1], E.ipt[2][2], tau, E.block_days, ps, vs);
  const double dem = 1.0 + E.emrat;
  pos = V3{((pe.x - pm.x / dem) - ps.x) / kAuKm, ((pe.y - pm.y / dem) - ps.y) / kAuKm, ((pe.z - pm.z / dem) - ps.z) / kAuKm};
  vel = V3{((ve.x - vm.x / dem) - vs.x) / kAuKm, ((ve.y - vm.y / dem) - vs.y) / kAuKm, ((ve.z - vm.z / dem) - vs.z) / kAuKm};
  return true;
}

// observer_pv (apparent_position.rs:264-296), one thread per epoch.
// table: [9][e_stride] = obs_pos_equ xyz, obs_vel_equ xyz (= Earth velocity), earth_pos_equ xyz
__global__ void __launch_bounds__(128)
ephemeris_observer_kernel(EphemDev E, size_t n_epochs, size_t e_stride, const double *__restrict__ mjd_tt,
                          const double *__restrict__ mjd_ut1, const double *__restrict__ bf_epoch, double bfx, double bfy,
                          double bfz, double *__restrict__ table, int *__restrict__ status) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= e_stride) return;
  double o[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) o[q] = NAN;
  int st = 0;
  if (e < n_epochs) {
    // bf_epoch: [3][n_epochs] per-epoch body-fixed position (a request with several observers) or null
    const V3 bfv = bf_epoch ? V3{bf_epoch[e], bf_epoch[n_epochs + e], bf_epoch[2 * n_epochs + e]} : V3{bfx, bfy, bfz};
    const V3 geo = pvobs_position(mjd_tt[e], mjd_ut1[e], bfv);
    V3 ep, ev;
    if (earth_posvel(E, mjd_tt[e], ep, ev)) {
      const V3 op = ep + ecl_to_equ(geo);
      o[0] = op.x; o[1] = op.y; o[2] = op.z; o[3] = ev.x; o[4] = ev.y; o[5] = ev.z; o[6] = ep.x; o[7] = ep.y; o[8] = ep.z;
    } else {
      st = 17;
    }
    status[e] = st;
  }
#pragma unroll
  for (int q = 0; q < 9; ++q) table[(size_t)q * e_stride + e] = o[q];
}

// ---- bulk asynchronous copy global -> shared, completion on an mbarrier (TMA engine, sm_90+) ----
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// retropropagate (ephemeris/aberration.rs:223-234) with the hoisted per-orbit constants: heliocentric position,
// equatorial J2000, at the epoch `t` (= t_obs - separation / c).  Same statement of propagate_twobody as the
// main evaluation of the kernel below (equinoctial_element.rs:326-348, 639-759).  false <=> Kepler solve failed.
struct EphOrbitC {
  double a, h, k, lambda, t_ref, n_mot, lon_peri, ch, ck, bhk, sx0, cx0;
  V3 fv, gv;
};
__device__ __forceinline__ bool eph_retarded_position(const EphOrbitC &o, double t, bool active, V3 &pos_equ) {
  // called by EVERY lane of the warp (inactive ones with active = false): the Newton trip count is warp-uniform
  const double dt = t - o.t_ref;
  double lam1 = rem_euclid(o.lambda + o.n_mot * (dt - 0.0), kTwoPi);
  if (lam1 < o.lon_peri) lam1 += kTwoPi;
  const double eps = kEps * 1e2;
  double x = kPi + o.lon_peri, sF = o.sx0, cF = o.cx0;
  int iter = 0;
  bool run = active, last = false, ok = true, have = true;
  while (__any_sync(0xffffffffu, run)) {
    if (run) {
      if (!have) sincos_angle(x, &sF, &cF);
      have = false;
      if (last) {
        run = false;
      } else {
        const double f = x - o.k * sF + o.h * cF - lam1;
        const double d = 1.0 - o.k * cF - o.h * sF;
        if (fabs(f) < eps) {
          run = false;
        } else if (fabs(d) < eps) {
          if (iter == 0) { x = x + 1.0; iter = 1; }
          else { ok = false; run = false; }
        } else {
          const double x1 = x - bf_div(f, d);
          const bool conv = fabs(x - x1) < eps;
          x = x1;
          if (conv) last = true;
          else if (++iter >= 25) { ok = false; run = false; }
        }
      }
    }
  }
  const double xe = o.a * (o.ch * cF + o.bhk * sF - o.k);
  const double ye = o.a * (o.ck * sF + o.bhk * cF - o.h);
  pos_equ = ecl_to_equ(xe * o.fv + ye * o.gv);
  return ok;
}

constexpr int kEphThreads = 128;
constexpr int kEphTile = 128;  // epochs per shared-memory tile: 9 rows x 128 x 8 B = 9 KB

// One thread per orbit.  kind: 0 Keplerian, 1 Equinoctial, 2 Cometary; elem [6][n_orbits].
// out [9][n_epochs][n_orbits] = ra, dec, geocentric_dist, heliocentric_dist, phase_angle,
// solar_elongation, radial_velocity, d_ra_dt, d_dec_dt; status [n_epochs][n_orbits].
#ifndef OUTFIT_EPH_BPS
#define OUTFIT_EPH_BPS 6  // 80 registers, 24 warps per SM: 5.67 ms per 1e8 entries against 5.86 at 5 blocks and 6.16 at 4 (r2d)
#endif
// SECOND = AberrationOrder::Second (aberration.rs:195-209): the line of sight comes from two back-propagations by the
// light time instead of the linear shift; a separate instantiation, the first-order kernel is untouched by it.
// NBODY = PropagatorKind::NBody (propagator/mod.rs:93-101): the heliocentric state of every (epoch, orbit) entry comes
// from the DOP853 integration of k_nbody.cuh (state_in [6][n_epochs][n_orbits], equatorial J2000; state_status
// [n_epochs][n_orbits]) instead of the Kepler solve; everything after the state is the same code.
template <bool SECOND, bool NBODY = false>
__global__ void __launch_bounds__(kEphThreads, OUTFIT_EPH_BPS)
ephemeris_twobody_kernel(size_t n_orbits, const int *__restrict__ kind, const double *__restrict__ epoch,
                         const double *__restrict__ elem, size_t n_epochs, size_t e_stride,
                         const double *__restrict__ mjd_tt, const double *__restrict__ table,
                         const int *__restrict__ obs_status, double *__restrict__ out, int *__restrict__ status,
                         const double *__restrict__ state_in = nullptr, const int *__restrict__ state_status = nullptr) {
  __shared__ __align__(16) double tile[9 * kEphTile];
  __shared__ __align__(8) unsigned long long bar;
  const size_t i = (size_t)blockIdx.x * kEphThreads + threadIdx.x;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  // first tile in flight while the threads convert their elements
  const unsigned first = (unsigned)(n_epochs < (size_t)kEphTile ? ((n_epochs + 1) & ~(size_t)1) : kEphTile);
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, 9u * first * 8u);
    for (int q = 0; q < 9; ++q) bulk_g2s(tile + q * kEphTile, table + (size_t)q * e_stride, first * 8u, &bar);
  }
  // ---- per-orbit constants (OrbitalElements::compute preamble + the epoch-independent part of
  //      propagate_twobody) ----
  int orbit_status = i < n_orbits ? 0 : -1;  // -1: a thread beyond the batch (walks the epochs, stores nothing)
  double a = 0, h = 0, k = 0, lambda = 0, t_ref = 0, n_mot = 0, lon_peri = 0, ch = 0, ck = 0, bhk = 0;
  V3 fv = V3{0, 0, 0}, gv = V3{0, 0, 0};
  double sx0 = 0, cx0 = 0;  // sincos of the Newton start pi + lon_peri: the same for every epoch of the orbit
  if (i < n_orbits) {
    Equinoctial eq;
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
      if (to_equinoctial(orb, eq) != 0) orbit_status = 9;  // every entry: InvalidConversion (mod.rs:196-213)
    }
    if (orbit_status == 0) {
      const double e2 = eq.h * eq.h + eq.k * eq.k;
      if (sqrt(e2) >= 1.0) {
        orbit_status = 9;  // check_elliptical_orbit, wrapped as InvalidConversion (mod.rs:219-240)
      } else {
        a = eq.a; h = eq.h; k = eq.k; lambda = eq.lambda; t_ref = eq.epoch;
        n_mot = sqrt(kMu / ((a * a) * a));
        lon_peri = (e2 > kEps * 1e2) ? rem_euclid(atan2(h, k), kTwoPi) : 0.0;
        const double beta = 1.0 / (1.0 + sqrt(1.0 - e2));
        bhk = beta * h * k;
        ch = 1.0 - beta * (h * h);
        ck = 1.0 - beta * (k * k);
        const double u = 1.0 + eq.p * eq.p + eq.q * eq.q;
        const double inv_u = 1.0 / u;
        const double common = 2.0 * eq.p * eq.q * inv_u;
        fv = V3{(1.0 - eq.p * eq.p + eq.q * eq.q) * inv_u, common, -2.0 * eq.p * inv_u};
        gv = V3{common, (1.0 + eq.p * eq.p - eq.q * eq.q) * inv_u, 2.0 * eq.q * inv_u};
        sincos_angle(kPi + lon_peri, &sx0, &cx0);
      }
    }
  }
  unsigned parity = 0;
  for (size_t e0 = 0; e0 < n_epochs; e0 += kEphTile) {
    const unsigned te = (unsigned)(n_epochs - e0 < (size_t)kEphTile ? n_epochs - e0 : kEphTile);
    mbar_wait(&bar, parity);
    parity ^= 1u;
    // Every lane of the warp walks the epochs (threads beyond n_orbits carry orbit_status = -1 and store nothing), and
    // the Newton loop below runs a WARP-UNIFORM number of trips (vote): with per-lane `break`s the lanes that converged
    // first ran ahead into the ~700-instruction tail on their own, and the tail was issued twice per epoch at 16 of 32
    // lanes (ncu r2b: 2.01 executions of the tail per warp and epoch).  Same arithmetic per lane, same bits.
#pragma unroll 1
    for (unsigned j = 0; j < te; ++j) {
      const size_t e = e0 + j;
      double o[9];
      int st = orbit_status;
      if (st == 0) st = __ldg(obs_status + e);
      const double t_obs = __ldg(mjd_tt + e);
      const double dt = t_obs - t_ref;
      double lam1 = rem_euclid(lambda + n_mot * (dt - 0.0), kTwoPi);
      if (lam1 < lon_peri) lam1 += kTwoPi;
      // generalised Kepler equation, roots 0.0.8 Newton (equinoctial_element.rs:326-348)
      const double eps = kEps * 1e2;
      double x = kPi + lon_peri, sF = sx0, cF = cx0;
      int iter = 0;
      bool run = st == 0 && !NBODY, last = false, ok = true, have = true;  // the first evaluation point is per orbit (hoisted above)
      while (!NBODY && __any_sync(0xffffffffu, run)) {
        if (run) {
          if (!have) sincos_angle(x, &sF, &cF);
          have = false;
          if (last) {
            run = false;
          } else {
            const double f = x - k * sF + h * cF - lam1;
            const double d = 1.0 - k * cF - h * sF;
            if (fabs(f) < eps) {
              run = false;
            } else if (fabs(d) < eps) {
              if (iter == 0) { x = x + 1.0; iter = 1; }
              else { ok = false; run = false; }
            } else {
              const double x1 = x - bf_div(f, d);
              const bool conv = fabs(x - x1) < eps;
              x = x1;
              if (conv) last = true;
              else if (++iter >= 25) { ok = false; run = false; }
            }
          }
        }
      }
      if (st == 0 && !ok) st = 11;  // RootFindingError
      const double xe = a * (ch * cF + bhk * sF - k);
      const double ye = a * (ck * sF + bhk * cF - h);
      const double vc = bf_div(n_mot * (a * a), bf_sqrt(xe * xe + ye * ye));
      const double vxe = vc * (bhk * cF - ch * sF);
      const double vye = vc * (ck * cF - bhk * sF);
      V3 ap = ecl_to_equ(xe * fv + ye * gv);   // ROT_ECLMJ2000_TO_EQUMJ2000 * pos_ecl
      V3 av = ecl_to_equ(vxe * fv + vye * gv);
      if (NBODY) {
        const size_t ii = i < n_orbits ? i : n_orbits - 1;
        const size_t at = e * n_orbits + ii, pl = n_epochs * n_orbits;
        ap = V3{state_in[at], state_in[pl + at], state_in[2 * pl + at]};
        av = V3{state_in[3 * pl + at], state_in[4 * pl + at], state_in[5 * pl + at]};
        if (st == 0) st = state_status[at];
      }
      const V3 op = V3{tile[j], tile[kEphTile + j], tile[2 * kEphTile + j]};
      const V3 ov = V3{tile[3 * kEphTile + j], tile[4 * kEphTile + j], tile[5 * kEphTile + j]};
      const V3 ep = V3{tile[6 * kEphTile + j], tile[7 * kEphTile + j], tile[8 * kEphTile + j]};
      const double helio = bf_sqrt(dot(ap, ap));
      const V3 dgeo = ap - ep;
      const double geo = bf_sqrt(dot(dgeo, dgeo));
      const V3 raw = ap - op;
      V3 topo;
      if (SECOND) {
        const EphOrbitC oc{a, h, k, lambda, t_ref, n_mot, lon_peri, ch, ck, bhk, sx0, cx0, fv, gv};
        V3 r1, r2;
        bool ab_ok = eph_retarded_position(oc, t_obs - div_by_const(bf_sqrt(dot(raw, raw)), kVlightAu, 1.0 / kVlightAu), st == 0, r1);
        const V3 d1 = r1 - op;
        ab_ok = eph_retarded_position(oc, t_obs - div_by_const(bf_sqrt(dot(d1, d1)), kVlightAu, 1.0 / kVlightAu), st == 0 && ab_ok, r2) && ab_ok;
        topo = r2 - op;
        if (st == 0 && !ab_ok) st = 11;  // the back-propagation's Kepler solve failed: RootFindingError
      } else {
        const double ltt = div_by_const(bf_sqrt(dot(raw, raw)), kVlightAu, 1.0 / kVlightAu);  // RN(x / c), Markstein
        topo = raw - ltt * av;
      }
      o[0] = rem_euclid(atan2_finite(topo.y, topo.x), kTwoPi);
      o[1] = atan2_finite(topo.z, bf_sqrt(topo.x * topo.x + topo.y * topo.y));
      o[2] = geo;
      o[3] = helio;
      const double rho = bf_sqrt(dot(topo, topo));
      const double r_obs = bf_sqrt(dot(op, op));
      o[4] = acos_unit(clampd(bf_div(dot(ap, topo), helio * rho), -1.0, 1.0));
      o[5] = acos_unit(clampd(bf_div(-dot(op, topo), r_obs * rho), -1.0, 1.0));
      const V3 vt = av - ov;
      o[6] = bf_div(dot(topo, vt), rho);
      const double dxy2 = topo.x * topo.x + topo.y * topo.y;
      const double dxy = bf_sqrt(dxy2);
      if (dxy < kEps * rho) {
        o[7] = 0.0; o[8] = 0.0;
      } else {
        o[7] = bf_div(-topo.y * vt.x + topo.x * vt.y, dxy2);
        o[8] = bf_div(-topo.z * topo.x * vt.x - topo.z * topo.y * vt.y + dxy2 * vt.z, rho * rho * dxy);
      }
      if (st != 0) {
#pragma unroll
        for (int q = 0; q < 9; ++q) o[q] = NAN;
      }
      if (i < n_orbits) {
#pragma unroll
        for (int q = 0; q < 9; ++q) out[((size_t)q * n_epochs + e) * n_orbits + i] = o[q];
        status[e * n_orbits + i] = st;
      }
    }
    // next tile: every thread is done reading this one
    __syncthreads();
    const size_t en = e0 + kEphTile;
    if (en < n_epochs && threadIdx.x == 0) {
      const size_t left = n_epochs - en;
      const unsigned nt = (unsigned)(left < (size_t)kEphTile ? ((left + 1) & ~(size_t)1) : kEphTile);
      mbar_expect_tx(&bar, 9u * nt * 8u);
      for (int q = 0; q < 9; ++q) bulk_g2s(tile + q * kEphTile, table + (size_t)q * e_stride + en, nt * 8u, &bar);
    }
  }
}

}  // namespace ofb
