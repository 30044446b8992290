/*
 * oo.h -- CPU ORACLE for the batched-IOD hot path of FusRoman/Outfit (v4.1.0).
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithm, operation for operation, used (a) by tests/ as the parity checker
 * for the CUDA path, (b) by __graft_entry__.smoke(), (c) by bench.py's
 * cpu_baseline / --impl reference leg.  The product library
 * (outfit_b200/csrc -> liboutfit_b200.so) never links, includes or calls
 * anything in this directory.
 *
 * The reference is pure Rust and cannot be compiled in this image (no
 * cargo/rustc, crates not vendored), so there is no oracle/_ref build; the
 * oracle is pinned instead against the reference's own known-answer tests
 * (tests/golden/reference_kats.json, extracted from the reference's #[test]
 * blocks; see tests/test_oracle_kats.py).  Pieces whose only reference KATs
 * need DE440 / UT1 downloads (hifitime epoch quantisation, rand's SmallRng +
 * ziggurat, photom's error model) are "parity unpinned" and are kept OUT of
 * the oracle: those values enter through the batch as inputs.
 *
 * Conventions: matrices are column-major double[9] (m[3*c + r]), the storage
 * nalgebra uses, so the reference's `as_slice()` KATs compare 1:1.
 * Compile with -ffp-contract=off: Rust never contracts a*b+c into an FMA; the
 * only FMAs on the path are the explicit mul_add calls that are restated here
 * with fma().
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src).
 */
#ifndef OO_H
#define OO_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OO_DPI 6.283185307179586476925286766559   /* constants.rs: DPI = 2*PI */
#define OO_PI 3.14159265358979323846
#define OO_EPS 2.220446049250313e-16              /* f64::EPSILON */
#define OO_GAUSS_GRAV 0.01720209895               /* constants.rs:70 */
#define OO_T2000 51544.5
#define OO_AU 149597870.7
#define OO_M(m, r, c) ((m)[3 * (c) + (r)])

/* ---- error / status codes shared with the C-ABI (include/outfit_b200.h) --- */
enum {
  OO_OK = 0,
  OO_ERR_SINGULAR_DIRECTION_MATRIX = 1, /* OutfitError::SingularDirectionMatrix */
  OO_ERR_GAUSS_NO_ROOTS = 2,            /* GaussNoRootsFound */
  OO_ERR_POLY_ROOT_FAILED = 3,          /* PolynomialRootFindingFailed */
  OO_ERR_SPURIOUS_ROOT = 4,             /* SpuriousRootDetected */
  OO_ERR_VELOCITY_CORRECTION = 5,       /* VelocityCorrectionError */
  OO_ERR_NEWTON_KEPLER = 6,             /* NewtonRaphsonKeplerConvergence */
  OO_ERR_BRENT_KEPLER = 7,              /* BrentDekkerKeplerConvergence */
  OO_ERR_DEGENERATE_STATE = 8,          /* DegenerateState */
  OO_ERR_INVALID_CONVERSION = 9,        /* InvalidConversion (parabolic -> a) */
  OO_ERR_INVALID_ORBIT = 10,            /* InvalidOrbit (e<=1 hyperbolic M, e>=1 scorer) */
  OO_ERR_ROOT_FINDING = 11,             /* RootFindingError (roots crate) */
  OO_ERR_NON_FINITE_SCORE = 12,         /* NonFiniteScore(v) */
  OO_ERR_NO_FEASIBLE_TRIPLETS = 13,     /* NoFeasibleTriplets{..} */
  OO_ERR_NO_VIABLE_ORBIT = 14,          /* NoViableOrbit{cause,attempts} */
  OO_ERR_OBSERVATION_NOT_FOUND = 15,    /* ObservationNotFound */
  OO_ERR_INVALID_IOD_PARAMETER = 16,    /* InvalidIODParameter */
  OO_ERR_EPHEM_OUT_OF_RANGE = 17        /* panic "Time outside ephemeris range" */
};

/* ---- op counters (algorithmic-flop accounting) ---------------- */
typedef struct {
  uint64_t sfunct_calls, sfunct_terms;     /* stumpff.rs series terms          */
  uint64_t newton_steps;                   /* newton_solver.rs run_newton its  */
  uint64_t prelim_calls, prelim_steps;     /* prelim_elliptic/hyperbolic its   */
  uint64_t kepler_universal_solves;        /* UniversalKeplerParams::solve     */
  uint64_t brent_evals;
  uint64_t aberth_solves, aberth_sweeps;
  uint64_t gauss_solves, roots_accepted;
  uint64_t fg_iterations;                  /* pos_and_vel_correction outer its */
  uint64_t ecc_controls;
  uint64_t orbits_built;                   /* ccek1 + to_equinoctial           */
  uint64_t scorer_evals, scorer_newton;    /* ephemeris_error, Kepler-F steps  */
  uint64_t earth_cheb_evals;               /* earth_ephemeris (position)       */
  uint64_t pvobs_evals;
  uint64_t propagate_universal_calls;
} oo_counters;
void oo_counters_reset(void);
void oo_counters_get(oo_counters *out);    /* sums over all threads since reset */

/* ---- IODParams (initial_orbit_determination/mod.rs:225-266) -------------- */
typedef struct {
  uint64_t n_noise_realizations;
  double noise_scale, extf, dtmax, dt_min, dt_max_triplet, optimal_interval_time;
  uint64_t max_obs_for_triplets;
  uint32_t max_triplets;
  double gap_max;
  double max_ecc, max_perihelion_au, min_rho2_au;
  uint32_t aberth_max_iter;
  double aberth_eps, kepler_eps;
  uint64_t max_tested_solutions;
  double r2_min_au, r2_max_au;
  double newton_eps;
  uint64_t newton_max_it;
  double root_imag_eps;
} oo_iod_params;
void oo_iod_params_default(oo_iod_params *p);          /* mod.rs:308-344 */
int oo_iod_params_validate(const oo_iod_params *p);    /* mod.rs:544-624 */

/* ---- Kepler (src/kepler) -------------------------------------------------- */
enum { OO_SOLVER_NEWTON = 0, OO_SOLVER_BRENT = 1, OO_SOLVER_AUTO = 2 };
enum { OO_PARABOLIC_CARDANO = 0, OO_PARABOLIC_NEWTON = 1 };
typedef struct {
  double dt, r0, sig0, mu, alpha, e0;
  int kind;                /* SolverKind                          params.rs:49 */
  double convergency;      /* SolverParams::convergency           params.rs:26 */
  int has_psi_guess;
  double psi_guess;
  uint64_t max_iter_prelim_kepuni;
  int parabolic_method;
} oo_kepler_params;
typedef struct { double psi, s0, s1, s2, s3; } oo_kepler_solution;

void oo_kepler_params_default_solver(oo_kepler_params *p); /* params.rs:35-73 */
void oo_s_funct(double psi, double alpha, double s[4]);    /* stumpff.rs:78   */
double oo_prelim_elliptic(const oo_kepler_params *p);      /* prelim_elliptic.rs:72 */
double oo_prelim_hyperbolic(const oo_kepler_params *p);    /* prelim_hyperbolic.rs:45 */
double oo_prelim_parabolic(const oo_kepler_params *p);     /* prelim_parabolic.rs:120 */
int oo_prelim_kepuni(const oo_kepler_params *p, double *psi); /* params.rs:185 */
int oo_solve_kepuni_newton(const oo_kepler_params *p, oo_kepler_solution *out); /* newton_solver.rs:151 */
int oo_solve_kepuni_brent(const oo_kepler_params *p, oo_kepler_solution *out);  /* brent_dekker_solver.rs:469 */
int oo_kepler_solve(const oo_kepler_params *p, oo_kepler_solution *out);        /* params.rs:130 */
/* velocity.rs:94 ; returns OO_OK or error code; out: v[3], f, g, chi */
int oo_velocity_correction_with_guess(const double x1[3], const double x2[3], const double v2[3],
                                      double dt, double peri_max, double ecc_max, int has_guess,
                                      double chi_guess, double eps, double v_out[3], double *f,
                                      double *g, double *chi);
/* propagation.rs:114 ; out[11] = r1[3], v1[3], f, g, fdot, gdot, psi */
int oo_propagate_universal(const double r[3], const double v[3], double t0, double t1, int kind,
                           double convergency, double out[11]);
void oo_propagate_universal_batch(size_t n, const double *r0v0_soa /*6*n*/, const double *t0,
                                  const double *t1, int kind, double convergency,
                                  double *out_soa /*11*n*/, int32_t *status, int n_threads);

/* ---- elements (orb_elem.rs, orbit_type/) -------------------------------- */
enum { OO_ELEM_KEPLERIAN = 0, OO_ELEM_EQUINOCTIAL = 1, OO_ELEM_COMETARY = 2 };
typedef struct {
  int kind;          /* OO_ELEM_* */
  double epoch;      /* reference_epoch (MJD TT) */
  double e[6];       /* Keplerian (a,e,i,Omega,omega,M) | Cometary (q,e,i,Omega,omega,nu)
                        | Equinoctial (a,h,k,p,q,lambda) */
} oo_elements;
void oo_rotmt(double alpha, int axis, double m[9]);            /* ref_system.rs:453 */
/* orb_elem.rs:257 ; returns 0 when |h| == 0 (None) else 1 */
int oo_eccentricity_control(const double r[3], const double v[3], double peri_max, double ecc_max,
                            int *accepted, double *ecc, double *peri, double *energy);
void oo_ccek1(const double r[3], const double v[3], double epoch, oo_elements *out); /* orb_elem.rs:58 */
int oo_to_equinoctial(const oo_elements *in, oo_elements *out); /* orbit_type/mod.rs:399 */
/* equinoctial_element.rs:326 */
int oo_equinoctial_solve_kepler(const oo_elements *eq, double mean_longitude_t1, double lon_peri,
                                double *F);
/* equinoctial_element.rs:809 (compute_derivatives = false) */
int oo_propagate_twobody(const oo_elements *eq, double t0, double t1, double pos[3], double vel[3]);

/* ---- Gauss (initial_orbit_determination/gauss.rs) ------------------------ */
typedef struct {
  uint64_t idx[3];
  double ra[3], dec[3], t[3];
  double obs_pos[9]; /* observer_helio_position, columns = epochs (equatorial J2000) */
} oo_gauss_obs;
typedef struct {
  int corrected;     /* GaussResult::CorrectedOrbit (1) | PrelimOrbit (0) */
  oo_elements orbit;
} oo_gauss_result;
int oo_gauss_prelim(const oo_gauss_obs *g, double *tau1, double *tau3, double unit[9],
                    double inv_unit[9], double a[3], double b[3]);          /* gauss.rs:532 */
void oo_coeff_eight_poly(const oo_gauss_obs *g, const double unit[9], const double inv_unit[9],
                         const double a[3], const double b[3], double c630[3]); /* gauss.rs:585 */
/* aberth 0.4.1 (un-vendored crate): returns 0 converged, 1 max-iter, 2 failed; roots re/im [8] */
int oo_aberth8(const double poly[9], uint32_t max_iter, double eps, double re[8], double im[8],
               uint32_t *sweeps);
int oo_solve_8poly(const double poly[9], uint32_t max_iter, double aberth_eps, double root_eps,
                   double roots[8], int *n_roots);                          /* gauss.rs:648 */
int oo_position_vector_and_reference_epoch(const oo_gauss_obs *g, const oo_iod_params *p,
                                           const double unit[9], const double inv_unit[9],
                                           const double c[3], double pos[9], double *epoch); /* gauss.rs:702 */
void oo_gibbs_correction(const double pos[9], double tau1, double tau3, double v[3]); /* gauss.rs:754 */
int oo_pos_and_vel_correction(const oo_gauss_obs *g, const oo_iod_params *p, const double pos_in[9],
                              const double vel_in[3], const double unit[9], const double inv_unit[9],
                              double peri_max, double ecc_max, double err_max, uint64_t itmax,
                              double pos[9], double vel[3], double *epoch); /* gauss.rs:1284 ; 1=Some */
int oo_prelim_orbit_all(const oo_gauss_obs *g, const oo_iod_params *p, oo_gauss_result out[],
                        int cap, int *n_out);                               /* gauss.rs:1119 */
int oo_prelim_orbit(const oo_gauss_obs *g, const oo_iod_params *p, oo_gauss_result *out); /* gauss.rs:1238 */

/* ---- triplets (triplet_generation/) -------------------------------------- */
typedef struct { double weight; uint64_t i, j, k; } oo_weighted_triplet;
size_t oo_downsample_uniform_with_edges(size_t n, size_t max_keep, size_t *keep); /* index_generator.rs:66 */
/* enumerates like TripletIndexGenerator (index_generator.rs:133-271); returns count written (<= cap) */
size_t oo_enumerate_triplets(const double *epochs, size_t n, double dt_min, double dt_max,
                             uint64_t *ijk /*3*cap*/, size_t cap);
double oo_triplet_weight_with_inv(double t1, double t2, double t3, double inv_dtw); /* mod.rs:229 */
/* generate_triplets phase 1 (mod.rs:328-408): best-K ascending; returns count */
size_t oo_best_k_triplets(const double *mjd_tt, size_t n_obs, const oo_iod_params *p,
                          oo_weighted_triplet *out /*max_triplets*/);

/* ---- ephemeris table + observer geometry --------------------------------- */
typedef struct {
  const double *cheb;    /* n_blocks * block_stride doubles: per 32-d block the EMB, Moon, Sun
                            coefficient sets laid out exactly as in the DE record
                            (sub-interval major, then x|y|z, then coefficient) */
  size_t n_blocks;
  size_t block_stride;   /* doubles per block */
  double jd_start, jd_end, block_days;
  uint32_t ipt[3][3];    /* EMB, Moon, Sun: {offset into block (0-based, doubles), n_coeff, n_sub} */
  double emrat;
} oo_ephem_table;
/* jpl_ephem/mod.rs:145 + horizon_data.rs:711-849 + horizon_records.rs:204 ; AU, AU/day equatorial */
int oo_earth_ephemeris(const oo_ephem_table *tab, double mjd_tt, int with_vel, double pos[3],
                       double vel[3]);
/* HorizonRecord::interpolate on one record (horizon_records.rs:204-298); n_coeff <= 32, n_sub <= 8 */
void oo_cheb_record(const double *coeffs, uint32_t n_coeff, double tau, uint32_t n_sub, double span_days,
                    int with_vel, double pos[3], double vel[3]);
double oo_obleq(double tjm);                          /* earth_orientation.rs:119 */
void oo_nutn80(double tjm, double *dpsi, double *deps); /* :170 */
void oo_rnut80(double tjm, double m[9]);              /* :459 */
double oo_equequ(double tjm);                         /* :508 */
void oo_prec(double tjm, double m[9]);                /* :561 */
double oo_gmst(double tjm_ut1);                       /* time.rs:326 */
/* ref_system.rs:379 ; sys: 0 Equm, 1 Equt, 2 Eclm ; *_j2000 != 0 selects RefEpoch::J2000 */
int oo_rotpn(int src_sys, int src_j2000, double src_date, int dst_sys, int dst_j2000,
             double dst_date, double rot[9]);
void oo_rotpn_equt_date_to_eclm_j2000(double tjm, double m[9]); /* the pair pvobs uses */
void oo_earth_fixed_position(double lon_rad, double rho_cos_phi, double rho_sin_phi, double r[3],
                             double v[3]);            /* observer_extension.rs:159-178 */
/* observer_extension.rs:180 ; dx, dv ecliptic mean J2000 (AU, AU/day) */
void oo_pvobs(double mjd_tt, double mjd_ut1, const double r_bf[3], const double v_bf[3],
              double dx[3], double dv[3]);
/* observer_extension.rs:223 */
int oo_helio_position(const oo_ephem_table *tab, double mjd_tt, const double geo_ecl[3],
                      double helio_equ[3]);
/* observation_ephemeris.rs:303 : the scorer's observer position (equatorial) */
int oo_scorer_observer_position(const oo_ephem_table *tab, double mjd_tt, const double geo_ecl[3],
                                double obs_equ[3]);

/* ---- scorer + per-trajectory driver -------------------------------------- */
typedef struct {
  size_t n;
  const double *mjd_tt, *ra, *dec, *sigma_ra, *sigma_dec; /* time-sorted, radians */
  const double *helio_equ;   /* 3*n, AoS xyz: OutfitCache helio position (equatorial J2000) */
  const double *geo_ecl;     /* 3*n, AoS xyz: OutfitCache geocentric position (ecliptic J2000) */
  const double *scorer_obs_equ; /* optional 3*n: precomputed observation_ephemeris.rs:303 result
                                   (dedup variant; NULL = evaluate Earth per call like the reference) */
} oo_traj_view;
/* observation_ephemeris.rs:388 */
int oo_ephemeris_error(const oo_traj_view *tv, size_t i, const oo_ephem_table *tab,
                       const oo_elements *equi, double *chi2);
/* observation_ephemeris.rs:369 */
int oo_compute_apparent_position(const oo_traj_view *tv, size_t i, const oo_ephem_table *tab,
                                 const oo_elements *equi, double *ra, double *dec);
int oo_select_rms_interval(const oo_traj_view *tv, const uint64_t idx[3], const oo_iod_params *p,
                           size_t *i_start, size_t *i_end);     /* trajectory.rs:294 */
int oo_rms_orbit_error(const oo_traj_view *tv, const oo_ephem_table *tab, const oo_gauss_obs *g,
                       const oo_elements *equi, const oo_iod_params *p, int has_prune, double prune,
                       double *rms);                            /* trajectory.rs:352 */
typedef struct {
  int32_t status;           /* OO_OK | NO_FEASIBLE_TRIPLETS | NO_VIABLE_ORBIT | INVALID_CONVERSION | INVALID_ORBIT */
  int32_t cause;            /* NoViableOrbit.cause code */
  double cause_value;       /* NonFiniteScore payload */
  uint64_t attempts;        /* NoViableOrbit.attempts */
  double span;              /* NoFeasibleTriplets.span */
  int32_t corrected;
  int32_t element_kind;
  double epoch, elem[6];
  double rms;
  uint32_t triplet_idx[3];
  uint32_t triplet_rank;    /* rank in the ascending-weight list */
  uint32_t realization;     /* 0 = unperturbed */
} oo_iod_result;
/* trajectory.rs:429 ; noise_z: [n_triplets_found][n_noise][6] host-drawn standard normals or NULL
   (NULL requires n_noise_realizations == 0) */
void oo_estimate_best_orbit(const oo_traj_view *tv, const oo_ephem_table *tab,
                            const oo_iod_params *p, const double *noise_z, oo_iod_result *out);
/* obs_dataset_api.rs:145-207 restated over a flat batch; one task per trajectory (dynamic
   scheduling) like par_iter_traj_id.  dedup_earth != 0 evaluates the scorer's Earth position once
   per observation instead of once per (candidate, observation) (not the reference's behaviour;
   reported separately). */
void oo_fit_full_iod(size_t n_traj, const uint64_t *traj_offset, const double *mjd_tt,
                     const double *ra, const double *dec, const double *sigma_ra,
                     const double *sigma_dec, const double *helio_equ, const double *geo_ecl,
                     const oo_ephem_table *tab, const oo_iod_params *p, const double *noise_z,
                     const uint64_t *noise_offset, oo_iod_result *out, int n_threads,
                     int dedup_earth);

/* ---- noise deviates (oo_rng.c; rand / rand_distr restated, parity unpinned) ------------------ */
uint64_t oo_splitmix64_next(uint64_t *x);
void oo_xoshiro_seed_from_u64(uint64_t seed, uint64_t s[4]);   /* SmallRng::seed_from_u64 */
uint64_t oo_xoshiro_next(uint64_t s[4]);                        /* Xoshiro256++ */
void oo_ziggurat_tables(double x[257], double f[257]);
double oo_standard_normal(uint64_t s[4]);                       /* rand_distr::StandardNormal */
void oo_draw_noise(uint64_t seed, size_t n, double *out);       /* n deviates in draw order */

/* ---- N-body propagator (propagator/nbody.rs; oo_nbody.c) --------------------------------------- */
enum { OO_ERR_NBODY = 21 };     /* NBodyPropagationFailed (nbody.rs:516-522) */
typedef struct { double gm; double pos[3]; } oo_perturber; /* PerturberSnapshot (nbody.rs:17-32): heliocentric, ecliptic J2000, AU */
void oo_nbody_rhs(const double *y42, const oo_perturber *pert, size_t n_pert, double *dy42);  /* nbody.rs:339-356 */
int oo_dop853_nbody(double *y42, double span, const oo_perturber *pert, size_t n_pert, double atol, double rtol,
                    uint32_t max_steps, uint32_t *n_steps, uint32_t *n_rejected);
int oo_propagate_nbody(const oo_elements *equi, double t1_mjd_tt, const oo_perturber *pert, size_t n_pert, double atol,
                       double rtol, double pos[3], double vel[3], double stm[36], uint32_t *n_steps);
double oo_planet_gm(int body);  /* planet_gm.rs */

/* ---- two-body `Combined` ephemeris (ephemeris/mod.rs:189-292; oo_ephemeris.c) -------------- */
/* apparent_position.rs:264-296 : observer position / velocity (= Earth velocity) / Earth position,
   equatorial mean J2000, AU and AU/day */
void oo_set_aberration_order(int order); /* 1 First (default) | 2 Second (aberration.rs:60-75,195-234) */
int oo_ephemeris_observer_pv(const oo_ephem_table *tab, double mjd_tt, double mjd_ut1,
                             const double r_bf[3], double obs_pos_equ[3], double obs_vel_equ[3],
                             double earth_pos_equ[3]);
/* one entry: out[9] = ra, dec, geocentric_dist, heliocentric_dist, phase_angle, solar_elongation,
   radial_velocity, d_ra_dt, d_dec_dt (apparent_position.rs:315-340, geometry.rs:204-345) */
int oo_ephemeris_entry(const oo_elements *equi, double obs_time_mjd, const double obs_pos[3],
                       const double obs_vel[3], const double earth_pos[3], double out[9]);
/* OrbitalElements::compute::<Combined>, one orbit x one observer x n_epochs; out [9][n_epochs] */
void oo_ephemeris_nbody(const oo_ephem_table *tab, const oo_elements *orbit, size_t n_epochs, const double *mjd_tt,
                        const double *mjd_ut1, const double r_bf[3], const oo_perturber *pert, size_t n_pert, double atol,
                        double rtol, double *out, int32_t *status);
void oo_ephemeris_twobody(const oo_ephem_table *tab, const oo_elements *orbit, size_t n_epochs,
                          const double *mjd_tt, const double *mjd_ut1, const double r_bf[3],
                          double *out, int32_t *status);
/* batch.rs:134-183 over a flat batch; out [9][n_epochs][n_orbits], status [n_epochs][n_orbits] */
void oo_ephemeris_twobody_batch(const oo_ephem_table *tab, size_t n_orbits, const int32_t *kind,
                                const double *epoch, const double *elem, size_t n_epochs,
                                const double *mjd_tt, const double *mjd_ut1, const double r_bf[3],
                                double *out, int32_t *status, int n_threads, int dedup_observer);

/* ---- differential orbit correction (differential_orbit_correction/; oo_lsq.c) --------------- */
enum { OO_ERR_LSQ_INVERSION = 18,   /* DifferentialCorrectionFailed (normal-equation inversion) */
       OO_ERR_LSQ_BIZARRE = 19,     /* BizarreOrbit */
       OO_ERR_LSQ_DIVERGED = 20 };  /* DifferentialCorrectionDiverged */
enum { OO_LSQ_KIND_NONE = 0,          /* no orbit: `status` holds the IOD / conversion error */
       OO_LSQ_KIND_CORRECTED = 1,     /* FitOrbitResult::DifferentialCorrection */
       OO_LSQ_KIND_IOD_FALLBACK = 2 };/* the loop failed: the IOD result is returned (mod.rs:113) */
typedef struct {                    /* DifferentialCorrectionConfig, diff_cor.rs:100-192 */
  uint64_t max_newton_iterations, max_outlier_rejection_passes;
  double convergence_threshold, convergence_before_rejection_threshold;
  double rms_stagnation_ratio, rms_divergence_ratio;
  uint64_t max_stagnation_iterations;
  int32_t enable_outlier_rejection;
  double chi2_rejection_threshold, chi2_recovery_threshold;     /* OutlierRejectionConfig */
  double eccentricity_limit, min_semi_major_axis, max_semi_major_axis, min_periapsis_distance,
      max_apoapsis_distance;                                     /* EquinoctialLimits */
  int32_t free_elements[6];
} oo_lsq_config;
typedef struct {                    /* ObservationEquation, least_square.rs:60-101 */
  double d_ra[6], d_dec[6], residual_ra, residual_dec, weight_ra, weight_dec, weight_cross;
  int32_t active;
} oo_obs_equation;
typedef struct {                    /* ObsFitData, obs_fit_data.rs:60-116 ; selection 0/1/2 */
  double sigma_ra, sigma_dec, bias_ra, bias_dec, residual_ra, residual_dec, chi;
  int32_t selection;
} oo_obs_fit_data;
typedef struct {                    /* DifferentialCorrectionResult; matrices column-major 6x6 */
  double correction[6], normal_matrix[36], covariance[36], normalised_rms;
  uint64_t num_measurements;
  int32_t inversion_succeeded;
} oo_lsq_solution;
typedef struct {
  int32_t status;                   /* OO_OK or the error that left no orbit */
  int32_t kind;                     /* OO_LSQ_KIND_* */
  int32_t fallback_cause;           /* OO_ERR_LSQ_* when kind == IOD_FALLBACK */
  double epoch, elem[6];            /* CORRECTED: equinoctial (a,h,k,p,q,lambda); FALLBACK: the IOD elements */
  double sigma[6];                  /* EquinoctialUncertainty::from_covariance */
  double normal_matrix[36], covariance[36];
  double normalised_rms;            /* FALLBACK: the IOD rms */
  uint64_t total_newton_iterations, num_measurements;
} oo_lsq_result;
void oo_lsq_config_default(oo_lsq_config *c);
int oo_is_bizarre(const double eq[6], const oo_lsq_config *c);
void oo_compute_derivative(const double eq[6], double t0, double t1, double n, double lam1, double F,
                           double inv_u, double beta, double sF, double cF, double xe, double ye,
                           double vxe, double vye, const double fv[3], const double gv[3],
                           const double pos[3], const double vel[3], double dpos[18], double dvel[18]);
int oo_propagate_twobody_partials(const oo_elements *eq, double t0, double t1, double pos[3],
                                  double vel[3], double dpos[18], double dvel[18]);
int oo_obs_and_partials(const oo_traj_view *tv, size_t i, const oo_ephem_table *tab,
                        const oo_elements *equi, double *ra, double *dec, double d_ra[6],
                        double d_dec[6]);
void oo_lsq_set_nbody(const oo_perturber *pert, size_t n_pert, double atol, double rtol);  /* thread-local; n_pert = 0: two-body */
int oo_obs_and_partials_nbody(const oo_traj_view *tv, size_t i, const oo_ephem_table *tab, const oo_elements *equi,
                              const oo_perturber *pert, size_t n_pert, double atol, double rtol, double *ra,
                              double *dec, double d_ra[6], double d_dec[6]);   /* observation_ephemeris.rs:452-486 */
double oo_angular_diff(double a, double b);
int oo_invert_normal_matrix(const double m[36], double inv[36]);
void oo_solve_weighted_least_squares(size_t n, const oo_obs_equation *eqs, const int32_t free_elements[6],
                                     oo_lsq_solution *out);
void oo_rescale_covariance(double normal_matrix[36], double covariance[36], size_t num_free,
                           size_t num_measurements, double normalised_rms);
size_t oo_update_observation_selection(size_t n, oo_obs_fit_data *fit, const oo_obs_equation *eqs,
                                       const double covariance[36], double chi2_reject,
                                       double chi2_recover);
int oo_run_differential_correction(const oo_traj_view *tv, const oo_ephem_table *tab,
                                   const oo_elements *initial, const oo_lsq_config *cfg,
                                   oo_obs_fit_data *fit, oo_lsq_result *out);
void oo_differential_correction(const oo_traj_view *tv, const oo_ephem_table *tab, const oo_iod_result *iod,
                                const oo_lsq_config *cfg, oo_lsq_result *out, oo_obs_fit_data *fit);
void oo_equinoctial_to_keplerian(const double eq[6], double kep[6]);   /* keplerian_element.rs:185-233 */
void oo_jacobian_to_keplerian(const double eq[6], double jac[36]);     /* equinoctial_element.rs:1049-1140 */
void oo_propagate_covariance(const double jac[36], const double cov[36], double out[36]); /* uncertainty.rs:412 */
/* FitLSQ::fit_lsq (obs_dataset_api.rs:113-190) with initial_orbits = Some(IOD results), over a flat
   batch, one task per trajectory.  fit: [sum n] per-observation fit data (final state). */
void oo_fit_lsq(size_t n_traj, const uint64_t *traj_offset, const double *mjd_tt, const double *ra,
                const double *dec, const double *sigma_ra, const double *sigma_dec,
                const double *geo_ecl, const oo_ephem_table *tab, const oo_lsq_config *cfg,
                const oo_iod_result *iod, oo_lsq_result *out, oo_obs_fit_data *fit, int n_threads);

/* The same with DifferentialCorrectionConfig::propagator = PropagatorKind::NBody: pert[n_traj][n_pert] = the perturbers
   frozen at each trajectory's IOD epoch. */
void oo_fit_lsq_nbody(size_t n_traj, const uint64_t *traj_offset, const double *mjd_tt, const double *ra,
                      const double *dec, const double *sigma_ra, const double *sigma_dec, const double *geo_ecl,
                      const oo_ephem_table *tab, const oo_lsq_config *cfg, const oo_iod_result *iod,
                      const oo_perturber *pert, size_t n_pert, double atol, double rtol, oo_lsq_result *out,
                      oo_obs_fit_data *fit, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
