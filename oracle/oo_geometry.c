/* oo_geometry.c -- ORACLE (test infrastructure only): Earth orientation, frame rotations, GMST,
 * observer geometry and the DE-style Chebyshev Earth ephemeris.  Restates
 * src/earth_orientation.rs, src/ref_system.rs, src/time.rs::gmst, src/observer_extension.rs,
 * src/jpl_ephem/mod.rs::earth_ephemeris, src/jpl_ephem/horizon/{horizon_data,horizon_records}.rs.
 *
 * The nutation series below keeps the reference's exact factorisation and summation order
 * (earth_orientation.rs:170-423) because its KAT (earth_orientation.rs:606-611) is an exact f64
 * equality; the product's CUDA kernel evaluates the same IAU-1980 series from a coefficient table
 * instead and is compared against this oracle with a tolerance. */
#include <math.h>
#include <string.h>
#include "oo.h"
#include "oo_linalg.h"

#define RADSEC (OO_PI / 648000.0)
#define RADEG (OO_PI / 180.0)

/* earth_orientation.rs:119-129 */
double oo_obleq(double tjm) {
  double ob0 = ((23.0 * 3600.0 + 26.0 * 60.0) + 21.448) * RADSEC;
  double ob1 = -46.815 * RADSEC;
  double ob2 = -0.0006 * RADSEC;
  double ob3 = 0.00181 * RADSEC;
  double t = (tjm - OO_T2000) / 36525.0;
  return ((ob3 * t + ob2) * t + ob1) * t + ob0;
}

/* earth_orientation.rs:178-181 : explicit mul_add chain (the only FMAs on the path) */
static double as_rad(double a0, double a1, double a2, double a3, double t, double t2, double t3) {
  return fma(a3, t3, fma(a2, t2, fma(a1, t, a0))) * RADSEC;
}

/* earth_orientation.rs:170-423 */
void oo_nutn80(double tjm, double *dpsi_out, double *deps_out) {
  double t = (tjm - OO_T2000) / 36525.0;
  double t2 = t * t;
  double t3 = t2 * t;
  double l = as_rad(485866.733, 1717915922.633, 31.310, 0.064, t, t2, t3);
  double p = as_rad(1287099.804, 129596581.224, -0.577, -0.012, t, t2, t3);
  double f = as_rad(335778.877, 1739527263.137, -13.257, 0.011, t, t2, t3);
  double d = as_rad(1072261.307, 1602961601.328, -6.891, 0.019, t, t2, t3);
  double n = as_rad(450160.280, -6962890.539, 7.455, 0.008, t, t2, t3);
  double x = f + f;
  double sl = sin(l), cl = cos(l);
  double sp = sin(p), cp = cos(p);
  double sx = sin(x), cx = cos(x);
  double sd = sin(d), cd = cos(d);
  double sn = sin(n), cn = cos(n);
  double cp2 = 2.0 * cp * cp - 1.0;
  double sp2 = 2.0 * sp * cp;
  double cd2 = 2.0 * cd * cd - 1.0;
  double sd2 = 2.0 * sd * cd;
  double cn2 = 2.0 * cn * cn - 1.0;
  double sn2 = 2.0 * sn * cn;
  double cl2 = 2.0 * cl * cl - 1.0;
  double sl2 = 2.0 * sl * cl;
  double ca = cx * cd2 + sx * sd2;
  double sa = sx * cd2 - cx * sd2;
  double cb = ca * cn - sa * sn;
  double sb = sa * cn + ca * sn;
  double cc = cb * cn - sb * sn;
  double sc_ = sb * cn + cb * sn;
  double cv = cx * cd2 - sx * sd2;
  double sv = sx * cd2 + cx * sd2;
  double ce = cv * cn - sv * sn;
  double se = sv * cn + cv * sn;
  double cf = ce * cn - se * sn;
  double sf = se * cn + ce * sn;
  double cg = cl * cd2 + sl * sd2;
  double sg = sl * cd2 - cl * sd2;
  double ch = cx * cn2 - sx * sn2;
  double sh = sx * cn2 + cx * sn2;
  double cj = ch * cl - sh * sl;
  double sj = sh * cl + ch * sl;
  double ck = cj * cl - sj * sl;
  double sk = sj * cl + cj * sl;
  double cm = cx * cl2 + sx * sl2;
  double sm = sx * cl2 - cx * sl2;
  double cq = cl * cd + sl * sd;
  double sq = sl * cd - cl * sd;
  double cr = 2.0 * cq * cq - 1.0;
  double sr = 2.0 * sq * cq;
  double cs = cx * cn - sx * sn;
  double ss = sx * cn + cx * sn;
  double ct = cs * cl - ss * sl;
  double st = ss * cl + cs * sl;
  double cu = cf * cl + sf * sl;
  double su = sf * cl - cf * sl;
  double cw = cp * cg - sp * sg;
  double sw = sp * cg + cp * sg;
  double dpsi =
  -(171996.0 + 174.2 * t) * sn + (2062.0 + 0.2 * t) * sn2 + 46.0 * (sm * cn + cm * sn)
  - 11.0 * sm
  - 3.0 * (sm * cn2 + cm * sn2)
  - 3.0 * (sq * cp - cq * sp)
  - 2.0 * (sb * cp2 - cb * sp2)
  + (sn * cm - cn * sm)
  - (13187.0 + 1.6 * t) * sc_
  + (1426.0 - 3.4 * t) * sp
  - (517.0 - 1.2 * t) * (sc_ * cp + cc * sp)
  + (217.0 - 0.5 * t) * (sc_ * cp - cc * sp)
  + (129.0 + 0.1 * t) * sb
  + 48.0 * sr
  - 22.0 * sa
  + (17.0 - 0.1 * t) * sp2
  - 15.0 * (sp * cn + cp * sn)
  - (16.0 - 0.1 * t) * (sc_ * cp2 + cc * sp2)
  - 12.0 * (sn * cp - cn * sp);
  dpsi += -6.0 * (sn * cr - cn * sr) - 5.0 * (sb * cp - cb * sp)
  + 4.0 * (sr * cn + cr * sn)
  + 4.0 * (sb * cp + cb * sp)
  - 4.0 * sq
  + (sr * cp + cr * sp)
  + (sn * ca - cn * sa)
  - (sp * ca - cp * sa)
  + (sp * cn2 + cp * sn2)
  + (sn * cq - cn * sq)
  - (sp * ca + cp * sa)
  - (2274.0 + 0.2 * t) * sh
  + (712.0 + 0.1 * t) * sl
  - (386.0 + 0.4 * t) * ss
  - 301.0 * sj
  - 158.0 * sg
  + 123.0 * (sh * cl - ch * sl)
  + 63.0 * sd2
  + (63.0 + 0.1 * t) * (sl * cn + cl * sn)
  - (58.0 + 0.1 * t) * (sn * cl - cn * sl)
  - 59.0 * su
  - 51.0 * st
  - 38.0 * sf
  + 29.0 * sl2;
  dpsi += 29.0 * (sc_ * cl + cc * sl) - 31.0 * sk
  + 26.0 * sx
  + 21.0 * (ss * cl - cs * sl)
  + 16.0 * (sn * cg - cn * sg)
  - 13.0 * (sn * cg + cn * sg)
  - 10.0 * (se * cl - ce * sl)
  - 7.0 * (sg * cp + cg * sp)
  + 7.0 * (sh * cp + ch * sp)
  - 7.0 * (sh * cp - ch * sp)
  - 8.0 * (sf * cl + cf * sl)
  + 6.0 * (sl * cd2 + cl * sd2)
  + 6.0 * (sc_ * cl2 + cc * sl2)
  - 6.0 * (sn * cd2 + cn * sd2)
  - 7.0 * se
  + 6.0 * (sb * cl + cb * sl)
  - 5.0 * (sn * cd2 - cn * sd2)
  + 5.0 * (sl * cp - cl * sp)
  - 5.0 * (ss * cl2 + cs * sl2)
  - 4.0 * (sp * cd2 - cp * sd2);
  dpsi += 4.0 * (sl * cx - cl * sx) - 4.0 * sd - 3.0 * (sl * cp + cl * sp)
  + 3.0 * (sl * cx + cl * sx)
  - 3.0 * (sj * cp - cj * sp)
  - 3.0 * (su * cp - cu * sp)
  - 2.0 * (sn * cl2 - cn * sl2)
  - 3.0 * (sk * cl + ck * sl)
  - 3.0 * (sf * cp - cf * sp)
  + 2.0 * (sj * cp + cj * sp)
  - 2.0 * (sb * cl - cb * sl);
  dpsi += 2.0 * (sn * cl2 + cn * sl2) - 2.0 * (sl * cn2 + cl * sn2)
  + 2.0 * (sl * cl2 + cl * sl2)
  + 2.0 * (sh * cd + ch * sd)
  + (sn2 * cl - cn2 * sl)
  - (sg * cd2 - cg * sd2)
  + (sf * cl2 - cf * sl2)
  - 2.0 * (su * cd2 + cu * sd2)
  - (sr * cd2 - cr * sd2)
  + (sw * ch + cw * sh)
  - (sl * ce + cl * se)
  - (sf * cr - cf * sr)
  + (su * ca + cu * sa)
  + (sg * cp - cg * sp)
  + (sb * cl2 + cb * sl2)
  - (sf * cl2 + cf * sl2)
  - (st * ca - ct * sa)
  + (sc_ * cx + cc * sx)
  + (sj * cr + cj * sr)
  - (sg * cx + cg * sx);
  dpsi += (sp * cs + cp * ss) + (sn * cw - cn * sw)
  - (sn * cx - cn * sx)
  - (sh * cd - ch * sd)
  - (sp * cd2 + cp * sd2)
  - (sl * cv - cl * sv)
  - (ss * cp - cs * sp)
  - (sw * cn + cw * sn)
  - (sl * ca - cl * sa)
  + (sl2 * cd2 + cl2 * sd2)
  - (sf * cd2 + cf * sd2)
  + (sp * cd + cp * sd);
  double deps = (92025.0 + 8.9 * t) * cn - (895.0 - 0.5 * t) * cn2 - 24.0 * (cm * cn - sm * sn)
  + (cm * cn2 - sm * sn2)
  + (cb * cp2 + sb * sp2)
  + (5736.0 - 3.1 * t) * cc
  + (54.0 - 0.1 * t) * cp
  + (224.0 - 0.6 * t) * (cc * cp - sc_ * sp)
  - (95.0 - 0.3 * t) * (cc * cp + sc_ * sp)
  - 70.0 * cb
  + cr
  + 9.0 * (cp * cn - sp * sn)
  + 7.0 * (cc * cp2 - sc_ * sp2)
  + 6.0 * (cn * cp + sn * sp)
  + 3.0 * (cn * cr + sn * sr)
  + 3.0 * (cb * cp + sb * sp)
  - 2.0 * (cr * cn - sr * sn)
  - 2.0 * (cb * cp - sb * sp);
  deps += (977.0 - 0.5 * t) * ch - 7.0 * cl + 200.0 * cs + (129.0 - 0.1 * t) * cj
  - cg
  - 53.0 * (ch * cl + sh * sl)
  - 2.0 * cd2
  - 33.0 * (cl * cn - sl * sn)
  + 32.0 * (cn * cl + sn * sl)
  + 26.0 * cu
  + 27.0 * ct
  + 16.0 * cf
  - cl2
  - 12.0 * (cc * cl - sc_ * sl)
  + 13.0 * ck
  - cx
  - 10.0 * (cs * cl + ss * sl)
  - 8.0 * (cn * cg + sn * sg)
  + 7.0 * (cn * cg - sn * sg)
  + 5.0 * (ce * cl + se * sl)
  - 3.0 * (ch * cp - sh * sp)
  + 3.0 * (ch * cp + sh * sp)
  + 3.0 * (cf * cl - sf * sl)
  - 3.0 * (cc * cl2 - sc_ * sl2)
  + 3.0 * (cn * cd2 - sn * sd2)
  + 3.0 * ce
  - 3.0 * (cb * cl - sb * sl)
  + 3.0 * (cn * cd2 + sn * sd2)
  + 3.0 * (cs * cl2 - ss * sl2)
  + (cj * cp + sj * sp)
  + (cu * cp + su * sp)
  + (cn * cl2 + sn * sl2)
  + (ck * cl - sk * sl)
  + (cf * cp + sf * sp)
  - (cj * cp - sj * sp)
  + (cb * cl + sb * sl)
  - (cn * cl2 - sn * sl2)
  + (cl * cn2 - sl * sn2)
  - (ch * cd - sh * sd)
  - (cn2 * cl + sn2 * sl)
  - (cf * cl2 + sf * sl2)
  + (cu * cd2 - su * sd2)
  - (cw * ch - sw * sh)
  + (cl * ce - sl * se)
  + (cf * cr + sf * sr)
  - (cb * cl2 - sb * sl2);
  *dpsi_out = dpsi * 1e-4;
  *deps_out = deps * 1e-4;
}

/* earth_orientation.rs:459-479 */
void oo_rnut80(double tjm, double m[9]) {
  double epsm = oo_obleq(tjm);
  double dpsi, deps;
  oo_nutn80(tjm, &dpsi, &deps);
  dpsi *= RADSEC;
  double epst = epsm + deps * RADSEC;
  double r1[9], r2[9], r3[9], r12[9];
  oo_rotmt(epsm, 0, r1);
  oo_rotmt(-dpsi, 2, r2);
  oo_rotmt(-epst, 0, r3);
  oo_matmul(r1, r2, r12);
  oo_matmul(r12, r3, m);
}

/* earth_orientation.rs:508-518 */
double oo_equequ(double tjm) {
  double oblm = oo_obleq(tjm);
  double dpsi, deps;
  oo_nutn80(tjm, &dpsi, &deps);
  return RADSEC * dpsi * cos(oblm);
}

/* earth_orientation.rs:561-593 */
void oo_prec(double tjm, double m[9]) {
  double zed = 0.6406161 * RADEG, zd = 0.6406161 * RADEG, thd = 0.5567530 * RADEG;
  double zedd = 0.0000839 * RADEG, zdd = 0.0003041 * RADEG, thdd = -0.0001185 * RADEG;
  double zeddd = 0.0000050 * RADEG, zddd = 0.0000051 * RADEG, thddd = -0.0000116 * RADEG;
  double t = (tjm - OO_T2000) / 36525.0;
  double zeta = ((zeddd * t + zedd) * t + zed) * t;
  double z = ((zddd * t + zdd) * t + zd) * t;
  double theta = ((thddd * t + thdd) * t + thd) * t;
  double r1[9], r2[9], r3[9], r12[9];
  oo_rotmt(-zeta, 2, r1);
  oo_rotmt(theta, 1, r2);
  oo_rotmt(-z, 2, r3);
  oo_matmul(r1, r2, r12);
  oo_matmul(r12, r3, m);
}

/* time.rs:326-361 */
double oo_gmst(double tjm) {
  const double C0 = 24110.54841, C1 = 8640184.812866, C2 = 9.3104e-2, C3 = -6.2e-6;
  const double RAP = 1.00273790934;
  double itjm = floor(tjm);
  double t = (itjm - OO_T2000) / 36525.0;
  double gmst0 = ((C3 * t + C2) * t + C1) * t + C0;
  gmst0 *= OO_DPI / 86400.0;
  double fract = tjm - trunc(tjm); /* f64::fract */
  double h = fract * OO_DPI;
  double g = gmst0 + h * RAP;
  long long i = (long long)floor(g / OO_DPI);
  if (g < 0.0) i -= 1;
  g -= (double)i * OO_DPI;
  return g;
}

/* ref_system.rs:118-411 : general rotpn over {Equm, Equt, Eclm} x {J2000, Epoch(d)} */
enum { SYS_EQUM = 0, SYS_EQUT = 1, SYS_ECLM = 2 };
typedef struct { int sys; int is_j2000; double date; } refsys;
static double rs_date(const refsys *r) { return r->is_j2000 ? OO_T2000 : r->date; }
int oo_rotpn(int src_sys, int src_j2000, double src_date, int dst_sys, int dst_j2000,
             double dst_date, double rot[9]) {
  refsys cur = {src_sys, src_j2000, src_date}, dst = {dst_sys, dst_j2000, dst_date};
  double step[9], tmp[9];
  for (int i = 0; i < 9; i++) rot[i] = 0.0;
  OO_M(rot, 0, 0) = OO_M(rot, 1, 1) = OO_M(rot, 2, 2) = 1.0;
  for (int it = 0; it < 20; it++) {
    int epochs_equal = (cur.is_j2000 && dst.is_j2000) ? 1
                                                      : (fabs(rs_date(&cur) - rs_date(&dst)) <= 1e-6);
    if (!epochs_equal) {
      /* transform_to_equm_date :246-272 */
      if (cur.is_j2000) {
        if (cur.sys == SYS_ECLM) { oo_rotmt(-oo_obleq(OO_T2000), 1, step); cur.sys = SYS_EQUM; }
        else if (cur.sys == SYS_EQUT) { oo_rnut80(OO_T2000, tmp); oo_transpose(tmp, step); cur.sys = SYS_EQUM; }
        else {
          if (dst.is_j2000) return -1;
          oo_prec(dst.date, step);
          cur.is_j2000 = 0; cur.date = dst.date;
        }
      } else {
        if (cur.sys == SYS_ECLM) { oo_rotmt(-oo_obleq(cur.date), 1, step); cur.sys = SYS_EQUM; }
        else if (cur.sys == SYS_EQUT) { oo_rnut80(cur.date, tmp); oo_transpose(tmp, step); cur.sys = SYS_EQUM; }
        else { oo_prec(cur.date, tmp); oo_transpose(tmp, step); cur.is_j2000 = 1; }
      }
      oo_matmul(rot, step, rot);
      continue;
    }
    if (cur.sys == dst.sys) return 0;
    /* transform_to_target_system :297-312 */
    double d = rs_date(&cur);
    if (cur.sys == SYS_EQUT) { oo_rnut80(d, tmp); oo_transpose(tmp, step); cur.sys = SYS_EQUM; }
    else if (cur.sys == SYS_ECLM) { oo_rotmt(-oo_obleq(d), 0, step); cur.sys = SYS_EQUM; }
    else if (dst.sys == SYS_EQUT) { oo_rnut80(d, step); cur.sys = SYS_EQUT; }
    else if (dst.sys == SYS_ECLM) { oo_rotmt(oo_obleq(d), 0, step); cur.sys = SYS_ECLM; }
    else return -1;
    oo_matmul(rot, step, rot);
  }
  return -1;
}
void oo_rotpn_equt_date_to_eclm_j2000(double tjm, double m[9]) {
  oo_rotpn(SYS_EQUT, 0, tjm, SYS_ECLM, 1, 0.0, m);
}

/* observer_extension.rs:159-178 ; constants.rs ERAU, EARTH_ROTATION */
void oo_earth_fixed_position(double lon_rad, double rho_cos_phi, double rho_sin_phi, double r[3],
                             double v[3]) {
  const double ERAU = (6378137.0 / 1000.0) / OO_AU;
  double sl = sin(lon_rad), cl = cos(lon_rad);
  r[0] = ERAU * rho_cos_phi * cl;
  r[1] = ERAU * rho_cos_phi * sl;
  r[2] = ERAU * rho_sin_phi;
  double w[3] = {0.0, 0.0, OO_DPI * 1.00273790934};
  oo_cross3(w, r, v);
}

/* observer_extension.rs:180-221.  mjd_tt is the value Epoch::to_mjd_tt_days() returns and
 * mjd_ut1 the value epoch.to_ut1(provider).to_mjd_tai_days() returns: hifitime's epoch
 * arithmetic is not restated (parity unpinned: its only KATs need the UT1/DE440 downloads). */
void oo_pvobs(double mjd_tt, double mjd_ut1, const double r_bf[3], const double v_bf[3],
              double dx[3], double dv[3]) {
  oo_tls_cnt.pvobs_evals++;
  double gast = oo_gmst(mjd_ut1) + oo_equequ(mjd_tt);
  double rot[9], rot1[9], rot1t[9], rott[9], rotmat[9];
  oo_rotmt(-gast, 2, rot);
  oo_rotpn_equt_date_to_eclm_j2000(mjd_tt, rot1);
  oo_transpose(rot1, rot1t);
  oo_transpose(rot, rott);
  oo_matmul(rot1t, rott, rotmat);
  oo_matvec(rotmat, r_bf, dx);
  if (v_bf && dv) oo_matvec(rotmat, v_bf, dv);
}

/* horizon_records.rs:204-298 for one body; pos (and vel) in km (km/day) */
static void cheb_body(const double *blk, const uint32_t ipt[3], double tau, double block_days,
                      int with_vel, double pos[3], double vel[3]) {
  uint32_t off = ipt[0], nc = ipt[1], nsub = ipt[2];
  /* horizon_data.rs:774 */
  double fs = floor(tau * (double)nsub);
  double mx = (double)nsub - 1.0;
  size_t sub = (size_t)(fs < mx ? fs : mx);
  const double *cf = blk + off + (size_t)sub * nc * 3;
  long long dt1 = (long long)tau;
  double temp = (double)nsub * tau;
  double tc = 2.0 * (oo_rem_euclid(temp, 1.0) + (double)dt1) - 1.0;
  double twot = 0.0;
  double tch[32] = {0.0}, tder[32] = {0.0};
  tch[0] = 1.0;
  if (tc != tch[1]) { tch[1] = tc; twot = tc + tc; }
  for (uint32_t i = 2; i < nc; i++) tch[i] = twot * tch[i - 1] - tch[i - 2];
  double vfac = with_vel ? (2.0 * (double)nsub) / block_days : 0.0;
  tder[1] = 1.0;
  tder[2] = twot + twot;
  if (with_vel)
    for (uint32_t i = 3; i < nc; i++) tder[i] = twot * tder[i - 1] + 2.0 * tch[i - 1] - tder[i - 2];
  for (int ax = 0; ax < 3; ax++) {
    double s = 0.0, sv = 0.0;
    for (uint32_t i = 0; i < nc; i++) s += cf[ax * nc + i] * tch[i];
    pos[ax] = s;
    if (with_vel) {
      for (uint32_t i = 0; i < nc; i++) sv += cf[ax * nc + i] * tder[i];
      vel[ax] = vfac * sv;
    }
  }
}

/* HorizonRecord::interpolate (horizon_records.rs:204-298) on ONE record: coeffs[3][n_coeff] (x | y | z), tau in
 * [0, 1] over the record, n_sub as the reference passes it; pos / vel in the coefficients' units (per day over
 * `span_days`).  Exposed for the restatement of the reference's record tests (horizon_records.rs:356-520). */
void oo_cheb_record(const double *coeffs, uint32_t n_coeff, double tau, uint32_t n_sub, double span_days,
                    int with_vel, double pos[3], double vel[3]) {
  /* cheb_body picks the sub-interval's coefficient set itself (horizon_data.rs:774); a single record is the
   * same set for every sub-interval */
  double blk[8 * 3 * 32];
  const uint32_t ns = n_sub ? n_sub : 1;
  for (uint32_t sb = 0; sb < ns && sb < 8; sb++) memcpy(blk + (size_t)sb * n_coeff * 3, coeffs, sizeof(double) * 3 * n_coeff);
  const uint32_t ipt[3] = {0, n_coeff, ns};
  cheb_body(blk, ipt, tau, span_days, with_vel, pos, vel);
}

/* jpl_ephem/mod.rs:145-174 ; horizon_data.rs:711-735, 810-849 ; interpolation_result.rs:82 */
int oo_earth_ephemeris(const oo_ephem_table *tab, double et, int with_vel, double pos[3],
                       double vel[3]) {
  oo_tls_cnt.earth_cheb_evals++;
  double et_jd = 2400000.5 + trunc(et);
  if (et_jd < tab->jd_start || et_jd > tab->jd_end) return OO_ERR_EPHEM_OUT_OF_RANGE;
  size_t nr = (size_t)floor((et_jd - tab->jd_start) / tab->block_days);
  if (fabs(et_jd - tab->jd_end) < 1e-10) nr -= 1;
  if (nr >= tab->n_blocks) return OO_ERR_EPHEM_OUT_OF_RANGE;
  double interval_start = (double)nr * tab->block_days + tab->jd_start;
  double fract = et - trunc(et);
  double tau = ((et_jd - interval_start) + fract) / tab->block_days;
  const double *blk = tab->cheb + nr * tab->block_stride;
  double pe[3], pm[3], ps[3], ve[3], vm[3], vs[3];
  cheb_body(blk, tab->ipt[0], tau, tab->block_days, with_vel, pe, ve);
  cheb_body(blk, tab->ipt[1], tau, tab->block_days, with_vel, pm, vm);
  cheb_body(blk, tab->ipt[2], tau, tab->block_days, with_vel, ps, vs);
  double dem = 1.0 + tab->emrat;
  for (int i = 0; i < 3; i++) {
    pos[i] = ((pe[i] - pm[i] / dem) - ps[i]) / OO_AU;
    if (with_vel) vel[i] = ((ve[i] - vm[i] / dem) - vs[i]) / OO_AU;
  }
  return OO_OK;
}

/* constants.rs:107-121 ROT_ECLMJ2000_TO_EQUMJ2000 / :93-105 inverse, column-major */
static const double ROT_ECL2EQU[9] = {1.0, 0.0, 0.0,
                                      0.0, 9.174820620691818e-1, 3.977771559319137e-1,
                                      0.0, -3.977771559319137e-1, 9.174820620691818e-1};
static const double ROT_EQU2ECL[9] = {1.0, 0.0, 0.0,
                                      0.0, 9.174820620691818e-1, -3.977771559319137e-1,
                                      0.0, 3.977771559319137e-1, 9.174820620691818e-1};

/* observer_extension.rs:223-237 */
int oo_helio_position(const oo_ephem_table *tab, double mjd_tt, const double geo_ecl[3],
                      double helio_equ[3]) {
  double e[3], r[3];
  int rc = oo_earth_ephemeris(tab, mjd_tt, 0, e, NULL);
  if (rc != OO_OK) return rc;
  oo_matvec(ROT_ECL2EQU, geo_ecl, r);
  for (int i = 0; i < 3; i++) helio_equ[i] = e[i] + r[i];
  return OO_OK;
}

/* observation_ephemeris.rs:303-318 */
int oo_scorer_observer_position(const oo_ephem_table *tab, double mjd_tt, const double geo_ecl[3],
                                double obs_equ[3]) {
  double e[3], ee[3], s[3];
  int rc = oo_earth_ephemeris(tab, mjd_tt, 0, e, NULL);
  if (rc != OO_OK) return rc;
  oo_matvec(ROT_EQU2ECL, e, ee);
  for (int i = 0; i < 3; i++) s[i] = geo_ecl[i] + ee[i];
  oo_matvec(ROT_ECL2EQU, s, obs_equ);
  return OO_OK;
}
