// dev_geometry.cuh -- per-observation observer geometry on the device.
//
// Reference behaviour:
//   JPLEphem::earth_ephemeris (Horizons/DE backend)   jpl_ephem/mod.rs:145-174,
//       horizon/horizon_data.rs:711-849, horizon/horizon_records.rs:204-298
//   Observer::pvobs / helio_position                  observer_extension.rs:180-237
//   equequ, rnut80, nutn80, prec, obleq               earth_orientation.rs:119-593
//   rotpn(Equt(date) -> Eclm(J2000)), rotmt           ref_system.rs:379-462
//   gmst                                              time.rs:326-361
//   scorer's observer position                        observation_ephemeris.rs:303-318
//
// Design differences from the reference: the IAU-1980 nutation series is evaluated from a
// 106-row coefficient table in constant memory (the reference uses a hand-factorised
// angle-addition form and evaluates it twice per observation; here once), the rotation chain is
// composed analytically from 3 elementary rotations' sines/cosines, and the Chebyshev recurrence
// runs once per body on registers (no basis vectors are materialised).
#pragma once
#include "dev_kepler.cuh"

namespace ofb {

constexpr unsigned kMaxCheb = 18;

struct EphemDev {
  const double *cheb;
  size_t n_blocks, block_stride;
  double jd_start, jd_end, block_days;
  unsigned ipt[3][3];  // EMB, Moon, Sun: 0-based offset in block, n_coeff, n_sub
  double emrat;
};

// Chebyshev position of one body at normalised block time tau in [0,1] (km)
__device__ __forceinline__ V3 cheb_position(const double *__restrict__ blk, unsigned off, unsigned nc,
                                            unsigned nsub, double tau) {
  const double fs = floor(tau * (double)nsub);
  const double mx = (double)nsub - 1.0;
  const unsigned sub = (unsigned)(fs < mx ? fs : mx);
  const double *cf = blk + off + (size_t)sub * nc * 3;
  const double temp = (double)nsub * tau;
  const double tc = 2.0 * (rem_euclid(temp, 1.0) + (double)(long long)tau) - 1.0;
  const double twot = tc + tc;
  double tm2 = 1.0, tm1 = tc;
  double x = __ldg(cf) * 1.0, y = __ldg(cf + nc) * 1.0, z = __ldg(cf + 2 * nc) * 1.0;
  x += __ldg(cf + 1) * tc; y += __ldg(cf + nc + 1) * tc; z += __ldg(cf + 2 * nc + 1) * tc;
  for (unsigned i = 2; i < nc; ++i) {
    const double ti = twot * tm1 - tm2;
    x += __ldg(cf + i) * ti;
    y += __ldg(cf + nc + i) * ti;
    z += __ldg(cf + 2 * nc + i) * ti;
    tm2 = tm1;
    tm1 = ti;
  }
  return V3{x, y, z};
}

// heliocentric Earth (equatorial J2000, AU); false <=> time outside the table (reference panics)
__device__ __forceinline__ bool earth_position(const EphemDev &E, double et, V3 &earth) {
  const double et_jd = 2400000.5 + trunc(et);
  if (et_jd < E.jd_start || et_jd > E.jd_end) return false;
  long long nr = (long long)floor((et_jd - E.jd_start) / E.block_days);
  if (fabs(et_jd - E.jd_end) < 1e-10) nr -= 1;
  if (nr < 0 || (size_t)nr >= E.n_blocks) return false;
  const double interval_start = (double)nr * E.block_days + E.jd_start;
  const double tau = ((et_jd - interval_start) + (et - trunc(et))) / E.block_days;
  const double *blk = E.cheb + (size_t)nr * E.block_stride;
  const V3 emb = cheb_position(blk, E.ipt[0][0], E.ipt[0][1], E.ipt[0][2], tau);
  const V3 moon = cheb_position(blk, E.ipt[1][0], E.ipt[1][1], E.ipt[1][2], tau);
  const V3 sun = cheb_position(blk, E.ipt[2][0], E.ipt[2][1], E.ipt[2][2], tau);
  const double dem = 1.0 + E.emrat;
  earth = V3{((emb.x - moon.x / dem) - sun.x) / kAuKm, ((emb.y - moon.y / dem) - sun.y) / kAuKm,
             ((emb.z - moon.z / dem) - sun.z) / kAuKm};
  return true;
}

// ---- IAU 1980 nutation from the coefficient table ---------------------------------------------
struct NutTerm {
  signed char m[5];
  float pad_;
  double a0, a1, b0, b1;
};
__constant__ NutTerm c_nut[106] = {
#define NUT_ROW(l, lp, f, d, o, a0, a1, b0, b1) {{l, lp, f, d, o}, 0.f, a0, a1, b0, b1},
#include "nutation_rows.inc"
#undef NUT_ROW
};

constexpr double kRadSec = kPi / 648000.0;
constexpr double kRaDeg = kPi / 180.0;

__device__ __forceinline__ double obleq(double tjm) {
  const double ob0 = ((23.0 * 3600.0 + 26.0 * 60.0) + 21.448) * kRadSec;
  const double t = (tjm - 51544.5) / 36525.0;
  return ((0.00181 * kRadSec * t + -0.0006 * kRadSec) * t + -46.815 * kRadSec) * t + ob0;
}

// dpsi, deps in arcsec
__device__ __noinline__ void nutation_iau1980(double tjm, double &dpsi, double &deps) {
  const double t = (tjm - 51544.5) / 36525.0;
  const double t2 = t * t, t3 = t2 * t;
  double arg[5];
  arg[0] = fma(0.064, t3, fma(31.310, t2, fma(1717915922.633, t, 485866.733))) * kRadSec;    // l
  arg[1] = fma(-0.012, t3, fma(-0.577, t2, fma(129596581.224, t, 1287099.804))) * kRadSec;   // l'
  arg[2] = fma(0.011, t3, fma(-13.257, t2, fma(1739527263.137, t, 335778.877))) * kRadSec;   // F
  arg[3] = fma(0.019, t3, fma(-6.891, t2, fma(1602961601.328, t, 1072261.307))) * kRadSec;   // D
  arg[4] = fma(0.008, t3, fma(7.455, t2, fma(-6962890.539, t, 450160.280))) * kRadSec;       // Omega
  double sp = 0.0, se = 0.0;
  // smallest terms first: the sum is dominated by a handful of large amplitudes
  for (int i = 105; i >= 0; --i) {
    const NutTerm &T = c_nut[i];
    const double a = (double)T.m[0] * arg[0] + (double)T.m[1] * arg[1] + (double)T.m[2] * arg[2] +
                     (double)T.m[3] * arg[3] + (double)T.m[4] * arg[4];
    double s, c;
    sincos(a, &s, &c);
    sp += (T.a0 + T.a1 * t) * s;
    se += (T.b0 + T.b1 * t) * c;
  }
  dpsi = sp * 1e-4;
  deps = se * 1e-4;
}

struct M3 {
  double m[3][3];
};
__device__ __forceinline__ V3 mul(const M3 &A, V3 v) {
  return V3{(A.m[0][0] * v.x + A.m[0][1] * v.y) + A.m[0][2] * v.z, (A.m[1][0] * v.x + A.m[1][1] * v.y) + A.m[1][2] * v.z,
            (A.m[2][0] * v.x + A.m[2][1] * v.y) + A.m[2][2] * v.z};
}
__device__ __forceinline__ V3 mul_t(const M3 &A, V3 v) {  // A^T v
  return V3{(A.m[0][0] * v.x + A.m[1][0] * v.y) + A.m[2][0] * v.z, (A.m[0][1] * v.x + A.m[1][1] * v.y) + A.m[2][1] * v.z,
            (A.m[0][2] * v.x + A.m[1][2] * v.y) + A.m[2][2] * v.z};
}
__device__ __forceinline__ M3 matmul(const M3 &A, const M3 &B) {
  M3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C.m[i][j] = (A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j]) + A.m[i][2] * B.m[2][j];
  return C;
}
// active rotation about a principal axis (nalgebra Rotation3::from_axis_angle; ref_system.rs:453)
__device__ __forceinline__ M3 rot_axis(double ang, int axis) {
  double s, c;
  sincos(ang, &s, &c);
  M3 R = {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}};
  if (axis == 0) { R.m[1][1] = c; R.m[1][2] = -s; R.m[2][1] = s; R.m[2][2] = c; }
  else if (axis == 1) { R.m[0][0] = c; R.m[0][2] = s; R.m[2][0] = -s; R.m[2][2] = c; }
  else { R.m[0][0] = c; R.m[0][1] = -s; R.m[1][0] = s; R.m[1][1] = c; }
  return R;
}

__device__ __forceinline__ double gmst(double tjm) {
  const double itjm = floor(tjm);
  const double t = (itjm - 51544.5) / 36525.0;
  double g0 = ((-6.2e-6 * t + 9.3104e-2) * t + 8640184.812866) * t + 24110.54841;
  g0 *= kTwoPi / 86400.0;
  double g = g0 + ((tjm - trunc(tjm)) * kTwoPi) * 1.00273790934;
  long long i = (long long)floor(g / kTwoPi);
  if (g < 0.0) i -= 1;
  return g - (double)i * kTwoPi;
}

// geocentric observer position in ecliptic mean J2000 from body-fixed coordinates (pvobs)
__device__ __forceinline__ V3 pvobs_position(double mjd_tt, double mjd_ut1, V3 r_bf) {
  double dpsi, deps;
  nutation_iau1980(mjd_tt, dpsi, deps);
  const double epsm = obleq(mjd_tt);
  const double gast = gmst(mjd_ut1) + kRadSec * dpsi * cos(epsm);
  // rnut80 = R_x(eps_m) R_z(-dpsi) R_x(-(eps_m + deps))
  const M3 rn = matmul(matmul(rot_axis(epsm, 0), rot_axis(-(dpsi * kRadSec), 2)), rot_axis(-(epsm + deps * kRadSec), 0));
  // prec = R_z(-zeta) R_y(theta) R_z(-z)
  const double t = (mjd_tt - 51544.5) / 36525.0;
  const double zeta = ((0.0000050 * kRaDeg * t + 0.0000839 * kRaDeg) * t + 0.6406161 * kRaDeg) * t;
  const double z = ((0.0000051 * kRaDeg * t + 0.0003041 * kRaDeg) * t + 0.6406161 * kRaDeg) * t;
  const double theta = ((-0.0000116 * kRaDeg * t + -0.0001185 * kRaDeg) * t + 0.5567530 * kRaDeg) * t;
  const M3 pr = matmul(matmul(rot_axis(-zeta, 2), rot_axis(theta, 1)), rot_axis(-z, 2));
  // rot1 = rnut^T prec^T R_x(obl(J2000));  dx = rot1^T rot(-gast)^T r_bf
  const V3 a = mul_t(rot_axis(-gast, 2), r_bf);  // undo Earth rotation
  const V3 b = mul(rn, a);                        // (rnut^T)^T
  const V3 c = mul(pr, b);                        // (prec^T)^T
  return mul_t(rot_axis(obleq(51544.5), 0), c);
}

// ---- kernels ------------------------------------------------------------------------------------
// scorer observer position: ROT_ecl->equ (geo_ecl + ROT_equ->ecl earth)   [3][n] plane-major
__global__ void __launch_bounds__(128)
scorer_observer_kernel(EphemDev E, size_t n, const double *__restrict__ mjd_tt, const double *__restrict__ geo_ecl,
                       double *__restrict__ scorer, int *__restrict__ status) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  V3 earth;
  V3 o = V3{NAN, NAN, NAN};
  int st = 0;
  if (earth_position(E, mjd_tt[i], earth)) {
    const V3 g = V3{geo_ecl[i], geo_ecl[n + i], geo_ecl[2 * n + i]};
    o = ecl_to_equ(g + equ_to_ecl(earth));
  } else {
    st = 17;
  }
  scorer[i] = o.x; scorer[n + i] = o.y; scorer[2 * n + i] = o.z;
  if (status) status[i] = st;
}

// OutfitCache build: geocentric (ecliptic) + heliocentric (equatorial) observer positions
__global__ void __launch_bounds__(128)
observer_cache_kernel(EphemDev E, size_t n, const double *__restrict__ mjd_tt, const double *__restrict__ mjd_ut1,
                      const double *__restrict__ bf, double *__restrict__ geo_ecl, double *__restrict__ helio_equ,
                      int *__restrict__ status) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V3 r_bf = V3{bf[i], bf[n + i], bf[2 * n + i]};
  const V3 dx = pvobs_position(mjd_tt[i], mjd_ut1[i], r_bf);
  geo_ecl[i] = dx.x; geo_ecl[n + i] = dx.y; geo_ecl[2 * n + i] = dx.z;
  V3 earth;
  V3 h = V3{NAN, NAN, NAN};
  int st = 0;
  if (earth_position(E, mjd_tt[i], earth)) h = earth + ecl_to_equ(dx);
  else st = 17;
  helio_equ[i] = h.x; helio_equ[n + i] = h.y; helio_equ[2 * n + i] = h.z;
  if (status) status[i] = st;
}

}  // namespace ofb
