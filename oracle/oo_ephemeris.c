/*
 * oo_ephemeris.c -- CPU ORACLE (test infrastructure only) for the two-body `Combined` ephemeris
 * (BASELINE configs[4]).
 *
 * Follows (paths relative to /root/reference/src):
 *   OrbitalElements::compute            ephemeris/mod.rs:189-292
 *   Combined::compute_one               ephemeris/request.rs:181-205
 *   propagate / observer_pv             ephemeris/apparent_position.rs:135-160, 264-296
 *   assemble_apparent_position          ephemeris/apparent_position.rs:315-340
 *   compute_geometry + helpers          ephemeris/geometry.rs:204-345
 *   PropagatorKind::TwoBody             propagator/mod.rs:84-91, 128-135
 *   correct_aberration_first_order      ephemeris/aberration.rs:139-145
 *   correct_aberration_second_order     ephemeris/aberration.rs:195-234 (AberrationOrder::Second, :60-75)
 *   check_elliptical_orbit              ephemeris/observation_ephemeris.rs:288-296
 *
 * Parity unpinned (the reference's KATs for this path need DE440 + UT1): hifitime's ns epoch
 * quantisation (the epochs enter as plain f64 MJD TT / UT1), photom's CartesianCoord -> EquCoord
 * (restated from the formula in the reference's doc comment, apparent_position.rs:343-357:
 * ra = atan2(y, x) mod 2 pi, dec = atan2(z, sqrt(x^2 + y^2))), and the Chebyshev velocity.
 */
#include <math.h>
#include <string.h>

#include "oo.h"
#include "oo_linalg.h"

static const double ROT_ECL2EQU[9] = {1.0, 0.0, 0.0,
                                      0.0, 9.174820620691818e-1, 3.977771559319137e-1,
                                      0.0, -3.977771559319137e-1, 9.174820620691818e-1};
static const double VLIGHT_AU = 2.99792458e5 / 149597870.7 * 86400.0; /* constants.rs */

/* apparent_position.rs:264-296 */
int oo_ephemeris_observer_pv(const oo_ephem_table *tab, double mjd_tt, double mjd_ut1,
                             const double r_bf[3], double obs_pos_equ[3], double obs_vel_equ[3],
                             double earth_pos_equ[3]) {
  double geo_ecl[3], rot_geo[3];
  oo_pvobs(mjd_tt, mjd_ut1, r_bf, NULL, geo_ecl, NULL);
  int rc = oo_earth_ephemeris(tab, mjd_tt, 1, earth_pos_equ, obs_vel_equ);
  if (rc != OO_OK) return rc;
  oo_matvec(ROT_ECL2EQU, geo_ecl, rot_geo);
  for (int i = 0; i < 3; i++) obs_pos_equ[i] = earth_pos_equ[i] + rot_geo[i];
  return OO_OK;
}

/* EphemerisConfig::aberration (ephemeris/mod.rs:129-142): 1 = AberrationOrder::First (default), 2 = Second.
 * A process-wide switch of the oracle (test infrastructure), set before a batch call. */
static int g_aberration_order = 1;
void oo_set_aberration_order(int order) { g_aberration_order = order == 2 ? 2 : 1; }

/* retropropagate (aberration.rs:223-234): heliocentric position, equatorial J2000, at t_obs - separation / c */
static int retropropagate(const oo_elements *equi, double obs_time_mjd, double separation, double r[3]) {
  double dt_light = separation / VLIGHT_AU;
  double t_retarded = obs_time_mjd - dt_light;
  double dt_orbit = t_retarded - equi->epoch;
  double pe[3], ve[3];
  int rc = oo_propagate_twobody(equi, 0.0, dt_orbit, pe, ve);
  if (rc != OO_OK) return rc;
  oo_matvec(ROT_ECL2EQU, pe, r);
  return OO_OK;
}

/* assemble_apparent_position + compute_geometry from the propagated heliocentric state (equatorial J2000) */
static int entry_from_state(const oo_elements *equi, double obs_time_mjd, const double ap[3], const double av[3],
                            const double obs_pos[3], const double obs_vel[3], const double earth_pos[3], double out[9]) {
  int rc;
  /* assemble_apparent_position */
  double helio = oo_norm3(ap);
  double dgeo[3], topo_raw[3];
  for (int i = 0; i < 3; i++) { dgeo[i] = ap[i] - earth_pos[i]; topo_raw[i] = ap[i] - obs_pos[i]; }
  double geo = oo_norm3(dgeo);
  double topo[3];
  if (g_aberration_order == 2) {
    /* correct_aberration_second_order (aberration.rs:195-209): two back-propagations by the light time; the two-body
     * propagator is used for both passes whatever the main propagator is */
    double r1[3], d1[3], r2[3];
    rc = retropropagate(equi, obs_time_mjd, oo_norm3(topo_raw), r1);
    if (rc != OO_OK) return rc;
    for (int i = 0; i < 3; i++) d1[i] = r1[i] - obs_pos[i];
    rc = retropropagate(equi, obs_time_mjd, oo_norm3(d1), r2);
    if (rc != OO_OK) return rc;
    for (int i = 0; i < 3; i++) topo[i] = r2[i] - obs_pos[i];
  } else {
    double ltt = oo_norm3(topo_raw) / VLIGHT_AU;
    for (int i = 0; i < 3; i++) topo[i] = topo_raw[i] - ltt * av[i];
  }
  out[0] = oo_rem_euclid(atan2(topo[1], topo[0]), OO_DPI);
  out[1] = atan2(topo[2], sqrt(topo[0] * topo[0] + topo[1] * topo[1]));
  out[2] = geo;
  out[3] = helio;
  /* compute_geometry */
  double rho = oo_norm3(topo);
  double r_obs = oo_norm3(obs_pos);
  out[4] = acos(oo_clamp(oo_dot3(ap, topo) / (helio * rho), -1.0, 1.0));
  out[5] = acos(oo_clamp(-oo_dot3(obs_pos, topo) / (r_obs * rho), -1.0, 1.0));
  double vt[3];
  for (int i = 0; i < 3; i++) vt[i] = av[i] - obs_vel[i];
  out[6] = oo_dot3(topo, vt) / rho;
  double dx = topo[0], dy = topo[1], dz = topo[2];
  double dxy2 = dx * dx + dy * dy;
  double dxy = sqrt(dxy2);
  if (dxy < OO_EPS * rho) {
    out[7] = 0.0;
    out[8] = 0.0;
  } else {
    out[7] = (-dy * vt[0] + dx * vt[1]) / dxy2;
    out[8] = (-dz * dx * vt[0] - dz * dy * vt[1] + dxy2 * vt[2]) / (rho * rho * dxy);
  }
  return OO_OK;
}

/* One (orbit, epoch) entry given the observer state: out[9] = ra, dec, geocentric_dist,
 * heliocentric_dist, phase_angle, solar_elongation, radial_velocity, d_ra_dt, d_dec_dt. */
int oo_ephemeris_entry(const oo_elements *equi, double obs_time_mjd, const double obs_pos[3],
                       const double obs_vel[3], const double earth_pos[3], double out[9]) {
  /* propagator/mod.rs:84-91 : dt from the reference epoch, propagate_twobody(0.0, dt, false) */
  double dt = obs_time_mjd - equi->epoch;
  double pe[3], ve[3], ap[3], av[3];
  int rc = oo_propagate_twobody(equi, 0.0, dt, pe, ve);
  if (rc != OO_OK) return rc;
  oo_matvec(ROT_ECL2EQU, pe, ap);
  oo_matvec(ROT_ECL2EQU, ve, av);
  return entry_from_state(equi, obs_time_mjd, ap, av, obs_pos, obs_vel, earth_pos, out);
}

/* The same entry with PropagatorKind::NBody (propagator/mod.rs:93-101: propagate_nbody, then the same rotation) */
int oo_ephemeris_entry_nbody(const oo_elements *equi, double obs_time_mjd, const oo_perturber *pert, size_t n_pert,
                             double atol, double rtol, const double obs_pos[3], const double obs_vel[3],
                             const double earth_pos[3], double out[9]) {
  double pe[3], ve[3], ap[3], av[3];
  int rc = oo_propagate_nbody(equi, obs_time_mjd, pert, n_pert, atol, rtol, pe, ve, NULL, NULL);
  if (rc != OO_OK) return rc;
  oo_matvec(ROT_ECL2EQU, pe, ap);
  oo_matvec(ROT_ECL2EQU, ve, av);
  return entry_from_state(equi, obs_time_mjd, ap, av, obs_pos, obs_vel, earth_pos, out);
}

/* OrbitalElements::compute::<Combined> with PropagatorKind::NBody for one orbit, one observer, n_epochs epochs */
void oo_ephemeris_nbody(const oo_ephem_table *tab, const oo_elements *orbit, size_t n_epochs, const double *mjd_tt,
                        const double *mjd_ut1, const double r_bf[3], const oo_perturber *pert, size_t n_pert, double atol,
                        double rtol, double *out, int32_t *status) {
  oo_elements equi;
  int rc = oo_to_equinoctial(orbit, &equi);
  if (rc == OO_OK) {
    double h = equi.e[1], k = equi.e[2];
    if (sqrt(h * h + k * k) >= 1.0) rc = OO_ERR_INVALID_CONVERSION;
  } else {
    rc = OO_ERR_INVALID_CONVERSION;
  }
  for (size_t e = 0; e < n_epochs; e++) {
    double o[9];
    for (int q = 0; q < 9; q++) o[q] = NAN;
    int st = rc;
    if (st == OO_OK) {
      double op[3], ov[3], ep[3];
      st = oo_ephemeris_observer_pv(tab, mjd_tt[e], mjd_ut1[e], r_bf, op, ov, ep);
      if (st == OO_OK) st = oo_ephemeris_entry_nbody(&equi, mjd_tt[e], pert, n_pert, atol, rtol, op, ov, ep, o);
      if (st != OO_OK)
        for (int q = 0; q < 9; q++) o[q] = NAN;
    }
    for (int q = 0; q < 9; q++) out[(size_t)q * n_epochs + e] = o[q];
    status[e] = st;
  }
}

/* OrbitalElements::compute::<Combined> for one orbit, one observer, n_epochs epochs.
 * out: [9][n_epochs] plane-major; status[n_epochs].  Conversion failures and e >= 1 mark every
 * entry InvalidConversion (mod.rs:196-240). */
void oo_ephemeris_twobody(const oo_ephem_table *tab, const oo_elements *orbit, size_t n_epochs,
                          const double *mjd_tt, const double *mjd_ut1, const double r_bf[3],
                          double *out, int32_t *status) {
  oo_elements equi;
  int rc = oo_to_equinoctial(orbit, &equi);
  if (rc == OO_OK) {
    double h = equi.e[1], k = equi.e[2];
    if (sqrt(h * h + k * k) >= 1.0) rc = OO_ERR_INVALID_CONVERSION; /* wrapped InvalidOrbit */
  } else {
    rc = OO_ERR_INVALID_CONVERSION;
  }
  for (size_t e = 0; e < n_epochs; e++) {
    double o[9];
    for (int q = 0; q < 9; q++) o[q] = NAN;
    int st = rc;
    if (st == OO_OK) {
      double op[3], ov[3], ep[3];
      st = oo_ephemeris_observer_pv(tab, mjd_tt[e], mjd_ut1[e], r_bf, op, ov, ep);
      if (st == OO_OK) st = oo_ephemeris_entry(&equi, mjd_tt[e], op, ov, ep, o);
      if (st != OO_OK)
        for (int q = 0; q < 9; q++) o[q] = NAN;
    }
    for (int q = 0; q < 9; q++) out[(size_t)q * n_epochs + e] = o[q];
    status[e] = st;
  }
}
