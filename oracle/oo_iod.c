#define _POSIX_C_SOURCE 200809L
/* oo_iod.c -- ORACLE (test infrastructure only): RMS scorer, per-trajectory best-orbit selection
 * and the batch driver.  Restates src/ephemeris/observation_ephemeris.rs:222-416,
 * src/ephemeris/aberration.rs:139-145, src/trajectory.rs:277-545 and
 * src/initial_orbit_determination/obs_dataset_api.rs:145-296 (the photom / RNG / hash-map parts
 * stay on the caller's side of the boundary: inputs are time-sorted, sigma-corrected observations
 * and, when n_noise_realizations > 0, the host-drawn standard-normal deviates in draw order). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>
#include "oo.h"
#include "oo_linalg.h"

_Thread_local oo_counters oo_tls_cnt;
static oo_counters g_cnt;
static pthread_mutex_t g_cnt_mu = PTHREAD_MUTEX_INITIALIZER;
static void counters_flush(void) {
  pthread_mutex_lock(&g_cnt_mu);
  uint64_t *g = (uint64_t *)&g_cnt, *t = (uint64_t *)&oo_tls_cnt;
  for (size_t i = 0; i < sizeof(oo_counters) / sizeof(uint64_t); i++) { g[i] += t[i]; t[i] = 0; }
  pthread_mutex_unlock(&g_cnt_mu);
}
void oo_counters_reset(void) {
  counters_flush();
  pthread_mutex_lock(&g_cnt_mu);
  memset(&g_cnt, 0, sizeof g_cnt);
  pthread_mutex_unlock(&g_cnt_mu);
}
void oo_counters_get(oo_counters *out) {
  counters_flush();
  pthread_mutex_lock(&g_cnt_mu);
  *out = g_cnt;
  pthread_mutex_unlock(&g_cnt_mu);
}

/* initial_orbit_determination/mod.rs:308-344 */
void oo_iod_params_default(oo_iod_params *p) {
  p->n_noise_realizations = 20;
  p->noise_scale = 1.0;
  p->extf = -1.0;
  p->dtmax = 30.0;
  p->dt_min = 0.03;
  p->dt_max_triplet = 150.0;
  p->optimal_interval_time = 20.0;
  p->max_obs_for_triplets = 100;
  p->max_triplets = 10;
  p->gap_max = 8.0 / 24.0;
  p->max_ecc = 5.0;
  p->max_perihelion_au = 1.0e3;
  p->min_rho2_au = 0.01;
  p->aberth_max_iter = 50;
  p->aberth_eps = 1.0e-6;
  p->kepler_eps = 1e3 * OO_EPS;
  p->max_tested_solutions = 3;
  p->r2_min_au = 0.05;
  p->r2_max_au = 200.0;
  p->newton_eps = 1.0e-10;
  p->newton_max_it = 50;
  p->root_imag_eps = 1.0e-6;
}
/* mod.rs:544-624 (NaN fails every comparison, like partial_cmp) */
int oo_iod_params_validate(const oo_iod_params *p) {
#define GE0(x) ((x) >= 0.0)
#define GT0(x) ((x) > 0.0)
  if (!GE0(p->noise_scale)) return OO_ERR_INVALID_IOD_PARAMETER;
  if (!GE0(p->dt_min) || !GE0(p->dt_max_triplet) || !GE0(p->dtmax)) return OO_ERR_INVALID_IOD_PARAMETER;
  if (!GE0(p->max_ecc)) return OO_ERR_INVALID_IOD_PARAMETER;
  if (!GE0(p->root_imag_eps)) return OO_ERR_INVALID_IOD_PARAMETER;
  if (!GT0(p->max_perihelion_au) || !GT0(p->min_rho2_au) || !GT0(p->aberth_eps) ||
      !GT0(p->kepler_eps) || !GT0(p->newton_eps))
    return OO_ERR_INVALID_IOD_PARAMETER;
  if (p->newton_max_it == 0 || p->aberth_max_iter == 0 || p->max_tested_solutions < 1)
    return OO_ERR_INVALID_IOD_PARAMETER;
  if (!(GT0(p->r2_min_au) && GT0(p->r2_max_au) && p->r2_min_au <= p->r2_max_au))
    return OO_ERR_INVALID_IOD_PARAMETER;
  return OO_OK;
}

static const double ROT_ECL2EQU[9] = {1.0, 0.0, 0.0,
                                      0.0, 9.174820620691818e-1, 3.977771559319137e-1,
                                      0.0, -3.977771559319137e-1, 9.174820620691818e-1};

/* observation_ephemeris.rs:369-386 (+ :222-275, :288-339, aberration.rs:139-145) */
int oo_compute_apparent_position(const oo_traj_view *tv, size_t i, const oo_ephem_table *tab,
                                 const oo_elements *equi, double *ra, double *dec) {
  const double vlight_au = 2.99792458e5 / OO_AU * 86400.0;
  double h = equi->e[1], k = equi->e[2];
  if (sqrt(h * h + k * k) >= 1.0) return OO_ERR_INVALID_ORBIT; /* check_elliptical_orbit */
  double dt = tv->mjd_tt[i] - equi->epoch;
  double pe[3], ve[3];
  int rc = oo_propagate_twobody(equi, 0.0, dt, pe, ve);
  if (rc != OO_OK) return rc;
  double obs[3];
  if (tv->scorer_obs_equ) {
    memcpy(obs, &tv->scorer_obs_equ[3 * i], sizeof obs);
  } else {
    rc = oo_scorer_observer_position(tab, tv->mjd_tt[i], &tv->geo_ecl[3 * i], obs);
    if (rc != OO_OK) return rc;
  }
  double ap[3], av[3];
  oo_matvec(ROT_ECL2EQU, pe, ap);
  oo_matvec(ROT_ECL2EQU, ve, av);
  double rel[3], cor[3];
  for (int c = 0; c < 3; c++) rel[c] = ap[c] - obs[c];
  double ltt = oo_norm3(rel) / vlight_au;
  for (int c = 0; c < 3; c++) cor[c] = rel[c] - ltt * av[c];
  double rho_xy = hypot(cor[0], cor[1]);
  *dec = atan2(cor[2], rho_xy);
  *ra = oo_rem_euclid(atan2(cor[1], cor[0]), OO_DPI);
  return OO_OK;
}

/* observation_ephemeris.rs:388-416 */
int oo_ephemeris_error(const oo_traj_view *tv, size_t i, const oo_ephem_table *tab,
                       const oo_elements *equi, double *chi2) {
  double alpha, delta;
  oo_tls_cnt.scorer_evals++;
  int rc = oo_compute_apparent_position(tv, i, tab, equi, &alpha, &delta);
  if (rc != OO_OK) return rc;
  double da = fmod(tv->ra[i] - alpha, OO_DPI);
  if (da > OO_PI) da -= OO_DPI;
  double dd = tv->dec[i] - delta;
  double a = cos(tv->dec[i]) * (da / tv->sigma_ra[i]);
  double b = dd / tv->sigma_dec[i];
  *chi2 = a * a + b * b;
  return OO_OK;
}

/* trajectory.rs:294-350 */
int oo_select_rms_interval(const oo_traj_view *tv, const uint64_t idx[3], const oo_iod_params *p,
                           size_t *i_start, size_t *i_end) {
  size_t nobs = tv->n;
  if (idx[0] >= nobs || idx[2] >= nobs || nobs == 0) return OO_ERR_OBSERVATION_NOT_FOUND;
  double t1 = tv->mjd_tt[idx[0]], t3 = tv->mjd_tt[idx[2]];
  double dt;
  if (p->extf >= 0.0) dt = (t3 - t1) * p->extf;
  else dt = 10.0 * (tv->mjd_tt[nobs - 1] - tv->mjd_tt[0]);
  if (p->dtmax >= 0.0) dt = (dt != dt) ? p->dtmax : (dt > p->dtmax ? dt : p->dtmax); /* f64::max */
  size_t is = 0;
  for (size_t ii = idx[0] + 1; ii-- > 0;) {
    if (t1 - tv->mjd_tt[ii] > dt) break;
    is = ii;
  }
  size_t ie = nobs - 1;
  for (size_t ii = idx[2]; ii < nobs; ii++) {
    if (tv->mjd_tt[ii] - t3 > dt) break;
    ie = ii;
  }
  *i_start = is;
  *i_end = ie;
  return OO_OK;
}

/* trajectory.rs:352-427 */
int oo_rms_orbit_error(const oo_traj_view *tv, const oo_ephem_table *tab, const oo_gauss_obs *g,
                       const oo_elements *equi, const oo_iod_params *p, int has_prune, double prune,
                       double *rms) {
  size_t is, ie;
  int rc = oo_select_rms_interval(tv, g->idx, p, &is, &ie);
  if (rc != OO_OK) return rc;
  double n_obs = (double)(ie - is + 1);
  double denom = 2.0 * n_obs;
  if (!has_prune) {
    double sum = 0.0;
    for (size_t i = is; i <= ie; i++) {
      double v;
      rc = oo_ephemeris_error(tv, i, tab, equi, &v);
      if (rc != OO_OK) return rc;
      sum = sum + v;
    }
    *rms = sqrt(sum / denom);
    return OO_OK;
  }
  double cutoff = isfinite(prune) ? prune * prune * denom : INFINITY;
  double sum = 0.0;
  for (size_t i = is; i <= ie; i++) {
    double v;
    rc = oo_ephemeris_error(tv, i, tab, equi, &v);
    if (rc != OO_OK) { *rms = prune; return OO_OK; }
    double ns = sum + v;
    if (ns >= cutoff) { *rms = prune; return OO_OK; }
    sum = ns;
  }
  *rms = sqrt(sum / denom);
  return OO_OK;
}

/* trajectory.rs:429-545 */
void oo_estimate_best_orbit(const oo_traj_view *tv, const oo_ephem_table *tab,
                            const oo_iod_params *p, const double *noise_z, oo_iod_result *out) {
  memset(out, 0, sizeof *out);
  out->rms = NAN;
  oo_weighted_triplet *trip =
      (oo_weighted_triplet *)malloc(sizeof(oo_weighted_triplet) * ((size_t)p->max_triplets + 1));
  size_t nt = oo_best_k_triplets(tv->mjd_tt, tv->n, p, trip);
  if (nt == 0) {
    out->status = OO_ERR_NO_FEASIBLE_TRIPLETS;
    out->span = tv->n == 0 ? 0.0 : tv->mjd_tt[tv->n - 1] - tv->mjd_tt[0];
    free(trip);
    return;
  }
  double best_rms = INFINITY;
  int have_best = 0;
  oo_gauss_result best;
  uint32_t best_rank = 0, best_real = 0;
  uint64_t best_idx[3] = {0, 0, 0};
  int last_err = 0;
  double last_err_val = 0.0;
  uint64_t attempts = 0;
  size_t nn = (size_t)p->n_noise_realizations;
  for (size_t r = 0; r < nt; r++) {
    oo_gauss_obs base;
    /* build_gauss_obs, triplet_generation/mod.rs:414-440 (indices are NOT mapped back through
       the down-sampling `keep` list -- reference behaviour) */
    base.idx[0] = trip[r].i; base.idx[1] = trip[r].j; base.idx[2] = trip[r].k;
    double sra[3], sdec[3];
    for (int c = 0; c < 3; c++) {
      size_t o = (size_t)base.idx[c];
      base.ra[c] = tv->ra[o];
      base.dec[c] = tv->dec[o];
      base.t[c] = tv->mjd_tt[o];
      for (int ax = 0; ax < 3; ax++) OO_M(base.obs_pos, ax, c) = tv->helio_equ[3 * o + ax];
      /* extract_errors trajectory.rs:277 ; realizations_iter gauss.rs:332-333 */
      sra[c] = tv->sigma_ra[o] * p->noise_scale;
      sdec[c] = tv->sigma_dec[o] * p->noise_scale;
    }
    for (size_t m = 0; m <= nn; m++) {
      oo_gauss_obs real = base;
      if (m > 0) {
        const double *z = &noise_z[(r * nn + (m - 1)) * 6];
        for (int c = 0; c < 3; c++) {
          real.ra[c] = base.ra[c] + z[c] * sra[c];
          real.dec[c] = base.dec[c] + z[3 + c] * sdec[c];
        }
      }
      attempts++;
      oo_gauss_result gr;
      int rc = oo_prelim_orbit(&real, p, &gr);
      if (rc != OO_OK) { last_err = rc; last_err_val = 0.0; continue; }
      oo_elements equi;
      rc = oo_to_equinoctial(&gr.orbit, &equi);
      if (rc != OO_OK) { /* `?` aborts the whole trajectory */
        out->status = rc;
        out->attempts = attempts;
        free(trip);
        return;
      }
      double rms;
      rc = oo_rms_orbit_error(tv, tab, &real, &equi, p, 1, best_rms, &rms);
      if (rc != OO_OK) { last_err = rc; last_err_val = 0.0; continue; }
      if (!isfinite(rms)) { last_err = OO_ERR_NON_FINITE_SCORE; last_err_val = rms; continue; }
      if (rms < best_rms) {
        best_rms = rms;
        best = gr;
        have_best = 1;
        best_rank = (uint32_t)r;
        best_real = (uint32_t)m;
        memcpy(best_idx, base.idx, sizeof best_idx);
      }
    }
  }
  out->attempts = attempts;
  if (have_best) {
    out->status = OO_OK;
    out->corrected = best.corrected;
    out->element_kind = best.orbit.kind;
    out->epoch = best.orbit.epoch;
    memcpy(out->elem, best.orbit.e, sizeof out->elem);
    out->rms = best_rms;
    for (int c = 0; c < 3; c++) out->triplet_idx[c] = (uint32_t)best_idx[c];
    out->triplet_rank = best_rank;
    out->realization = best_real;
  } else {
    out->status = OO_ERR_NO_VIABLE_ORBIT;
    out->cause = last_err;
    out->cause_value = last_err_val;
  }
  free(trip);
}

/* ---- pthread parallel-for with dynamic scheduling (one task = `chunk` consecutive units),
 *      the same decomposition as rayon's par_iter over trajectory ids -------------------- */
typedef struct {
  atomic_llong next;
  long long n, chunk;
  void (*fn)(long long lo, long long hi, void *ctx);
  void *ctx;
} pfor_t;
static void *pfor_worker(void *arg) {
  pfor_t *pf = (pfor_t *)arg;
  for (;;) {
    long long lo = atomic_fetch_add(&pf->next, pf->chunk);
    if (lo >= pf->n) break;
    long long hi = lo + pf->chunk < pf->n ? lo + pf->chunk : pf->n;
    pf->fn(lo, hi, pf->ctx);
  }
  counters_flush();
  return NULL;
}
static void parallel_for(long long n, long long chunk, int n_threads,
                         void (*fn)(long long, long long, void *), void *ctx) {
  if (n_threads <= 0) {
    long nc = sysconf(_SC_NPROCESSORS_ONLN);
    n_threads = nc > 0 ? (int)nc : 1;
  }
  if (n_threads > 256) n_threads = 256;
  pfor_t pf;
  atomic_init(&pf.next, 0);
  pf.n = n; pf.chunk = chunk > 0 ? chunk : 1; pf.fn = fn; pf.ctx = ctx;
  if (n_threads == 1) { pfor_worker(&pf); return; }
  pthread_t th[256];
  int started = 0;
  for (int i = 0; i < n_threads; i++)
    if (pthread_create(&th[started], NULL, pfor_worker, &pf) == 0) started++;
  if (started == 0) pfor_worker(&pf);
  for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
}

/* obs_dataset_api.rs:145-207 over a flat SoA batch */
typedef struct {
  const uint64_t *traj_offset;
  const double *mjd_tt, *ra, *dec, *sigma_ra, *sigma_dec, *helio_equ, *geo_ecl;
  const oo_ephem_table *tab;
  const oo_iod_params *p;
  const double *noise_z;
  const uint64_t *noise_offset;
  oo_iod_result *out;
  int dedup_earth;
} iod_ctx;
static void iod_task(long long lo, long long hi, void *vctx) {
  iod_ctx *c = (iod_ctx *)vctx;
  for (long long t = lo; t < hi; t++) {
    size_t o = (size_t)c->traj_offset[t], n = (size_t)(c->traj_offset[t + 1] - c->traj_offset[t]);
    oo_traj_view tv;
    tv.n = n;
    tv.mjd_tt = c->mjd_tt + o; tv.ra = c->ra + o; tv.dec = c->dec + o;
    tv.sigma_ra = c->sigma_ra + o; tv.sigma_dec = c->sigma_dec + o;
    tv.helio_equ = c->helio_equ + 3 * o; tv.geo_ecl = c->geo_ecl + 3 * o;
    tv.scorer_obs_equ = NULL;
    double *pre = NULL;
    if (c->dedup_earth) {
      pre = (double *)malloc(sizeof(double) * 3 * (n ? n : 1));
      for (size_t i = 0; i < n; i++)
        oo_scorer_observer_position(c->tab, tv.mjd_tt[i], &tv.geo_ecl[3 * i], &pre[3 * i]);
      tv.scorer_obs_equ = pre;
    }
    const double *nz = (c->noise_z && c->noise_offset) ? c->noise_z + c->noise_offset[t] : NULL;
    oo_estimate_best_orbit(&tv, c->tab, c->p, nz, &c->out[t]);
    free(pre);
  }
}
void oo_fit_full_iod(size_t n_traj, const uint64_t *traj_offset, const double *mjd_tt,
                     const double *ra, const double *dec, const double *sigma_ra,
                     const double *sigma_dec, const double *helio_equ, const double *geo_ecl,
                     const oo_ephem_table *tab, const oo_iod_params *p, const double *noise_z,
                     const uint64_t *noise_offset, oo_iod_result *out, int n_threads,
                     int dedup_earth) {
  iod_ctx c = {traj_offset, mjd_tt, ra, dec, sigma_ra, sigma_dec, helio_equ, geo_ecl,
               tab, p, noise_z, noise_offset, out, dedup_earth};
  parallel_for((long long)n_traj, 1, n_threads, iod_task, &c);
}

typedef struct {
  size_t n;
  const double *rv, *t0, *t1;
  int kind;
  double convergency;
  double *out;
  int32_t *status;
} pu_ctx;
static void pu_task(long long lo, long long hi, void *vctx) {
  pu_ctx *c = (pu_ctx *)vctx;
  size_t n = c->n;
  for (long long i = lo; i < hi; i++) {
    double r[3] = {c->rv[i], c->rv[n + i], c->rv[2 * n + i]};
    double v[3] = {c->rv[3 * n + i], c->rv[4 * n + i], c->rv[5 * n + i]};
    double o[11];
    for (int k = 0; k < 11; k++) o[k] = NAN;
    c->status[i] = oo_propagate_universal(r, v, c->t0[i], c->t1[i], c->kind, c->convergency, o);
    for (int k = 0; k < 11; k++) c->out[(size_t)k * n + i] = o[k];
  }
}
void oo_propagate_universal_batch(size_t n, const double *rv, const double *t0, const double *t1,
                                  int kind, double convergency, double *out, int32_t *status,
                                  int n_threads) {
  pu_ctx c = {n, rv, t0, t1, kind, convergency, out, status};
  parallel_for((long long)n, 4096, n_threads, pu_task, &c);
}

/* ---- FullOrbitResultExt::compute_ephemerides[_parallel] (ephemeris/batch.rs:134-183) over a flat
 *      batch: one task per orbit (rayon par_iter over the result map), every (orbit, epoch) entry
 *      recomputes the observer state like the reference unless dedup_observer != 0.
 *      elem: [6][n_orbits] plane-major; out: [9][n_epochs][n_orbits]; status: [n_epochs][n_orbits] */
typedef struct {
  const oo_ephem_table *tab;
  size_t n, n_epochs;
  const int32_t *kind;
  const double *epoch, *elem, *mjd_tt, *mjd_ut1, *r_bf;
  const double *obs_pv; /* [9][n_epochs] precomputed observer state or NULL */
  const int32_t *obs_st;
  double *out;
  int32_t *status;
} eph_ctx;
static void eph_task(long long lo, long long hi, void *vctx) {
  eph_ctx *c = (eph_ctx *)vctx;
  size_t n = c->n, E = c->n_epochs;
  for (long long i = lo; i < hi; i++) {
    oo_elements orb, equi;
    orb.kind = c->kind[i];
    orb.epoch = c->epoch[i];
    for (int q = 0; q < 6; q++) orb.e[q] = c->elem[(size_t)q * n + i];
    int rc = oo_to_equinoctial(&orb, &equi);
    if (rc == OO_OK) {
      double h = equi.e[1], k = equi.e[2];
      if (sqrt(h * h + k * k) >= 1.0) rc = OO_ERR_INVALID_CONVERSION;
    } else {
      rc = OO_ERR_INVALID_CONVERSION;
    }
    for (size_t e = 0; e < E; e++) {
      double o[9], op[3], ov[3], ep[3];
      int st = rc;
      if (st == OO_OK) {
        if (c->obs_pv) {
          st = c->obs_st[e];
          for (int q = 0; q < 3; q++) {
            op[q] = c->obs_pv[(size_t)q * E + e];
            ov[q] = c->obs_pv[(size_t)(3 + q) * E + e];
            ep[q] = c->obs_pv[(size_t)(6 + q) * E + e];
          }
        } else {
          st = oo_ephemeris_observer_pv(c->tab, c->mjd_tt[e], c->mjd_ut1[e], c->r_bf, op, ov, ep);
        }
        if (st == OO_OK) st = oo_ephemeris_entry(&equi, c->mjd_tt[e], op, ov, ep, o);
      }
      if (st != OO_OK)
        for (int q = 0; q < 9; q++) o[q] = NAN;
      for (int q = 0; q < 9; q++) c->out[((size_t)q * E + e) * n + i] = o[q];
      c->status[e * n + i] = st;
    }
  }
}
void oo_ephemeris_twobody_batch(const oo_ephem_table *tab, size_t n_orbits, const int32_t *kind,
                                const double *epoch, const double *elem, size_t n_epochs,
                                const double *mjd_tt, const double *mjd_ut1, const double r_bf[3],
                                double *out, int32_t *status, int n_threads, int dedup_observer) {
  eph_ctx c = {tab, n_orbits, n_epochs, kind, epoch, elem, mjd_tt, mjd_ut1, r_bf, NULL, NULL, out, status};
  double *pv = NULL;
  int32_t *st = NULL;
  if (dedup_observer) {
    pv = (double *)malloc(9 * n_epochs * sizeof(double));
    st = (int32_t *)malloc(n_epochs * sizeof(int32_t));
    for (size_t e = 0; e < n_epochs; e++) {
      double op[3] = {NAN, NAN, NAN}, ov[3] = {NAN, NAN, NAN}, ep[3] = {NAN, NAN, NAN};
      st[e] = oo_ephemeris_observer_pv(tab, mjd_tt[e], mjd_ut1[e], r_bf, op, ov, ep);
      for (int q = 0; q < 3; q++) {
        pv[(size_t)q * n_epochs + e] = op[q];
        pv[(size_t)(3 + q) * n_epochs + e] = ov[q];
        pv[(size_t)(6 + q) * n_epochs + e] = ep[q];
      }
    }
    c.obs_pv = pv;
    c.obs_st = st;
  }
  parallel_for((long long)n_orbits, 64, n_threads, eph_task, &c);
  free(pv);
  free(st);
}

/* FitLSQ::fit_lsq (differential_orbit_correction/obs_dataset_api.rs:113-190) over a flat SoA batch */
typedef struct {
  const uint64_t *traj_offset;
  const double *mjd_tt, *ra, *dec, *sigma_ra, *sigma_dec, *geo_ecl;
  const oo_ephem_table *tab;
  const oo_lsq_config *cfg;
  const oo_iod_result *iod;
  oo_lsq_result *out;
  oo_obs_fit_data *fit;
  const oo_perturber *pert;  /* [n_traj][n_pert] or NULL (two-body) */
  size_t n_pert;
  double atol, rtol;
} lsq_ctx;
static void lsq_task(long long lo, long long hi, void *vctx) {
  lsq_ctx *c = (lsq_ctx *)vctx;
  for (long long t = lo; t < hi; t++) {
    size_t o = (size_t)c->traj_offset[t], n = (size_t)(c->traj_offset[t + 1] - c->traj_offset[t]);
    oo_traj_view tv;
    tv.n = n;
    tv.mjd_tt = c->mjd_tt + o; tv.ra = c->ra + o; tv.dec = c->dec + o;
    tv.sigma_ra = c->sigma_ra + o; tv.sigma_dec = c->sigma_dec + o;
    tv.helio_equ = NULL; tv.geo_ecl = c->geo_ecl + 3 * o;
    tv.scorer_obs_equ = NULL;
    if (c->pert) oo_lsq_set_nbody(c->pert + (size_t)t * c->n_pert, c->n_pert, c->atol, c->rtol);
    oo_differential_correction(&tv, c->tab, &c->iod[t], c->cfg, &c->out[t], c->fit + o);
    if (c->pert) oo_lsq_set_nbody(NULL, 0, 0.0, 0.0);
  }
}
void oo_fit_lsq(size_t n_traj, const uint64_t *traj_offset, const double *mjd_tt, const double *ra,
                const double *dec, const double *sigma_ra, const double *sigma_dec,
                const double *geo_ecl, const oo_ephem_table *tab, const oo_lsq_config *cfg,
                const oo_iod_result *iod, oo_lsq_result *out, oo_obs_fit_data *fit, int n_threads) {
  lsq_ctx c = {traj_offset, mjd_tt, ra, dec, sigma_ra, sigma_dec, geo_ecl, tab, cfg, iod, out, fit, NULL, 0, 0.0, 0.0};
  parallel_for((long long)n_traj, 1, n_threads, lsq_task, &c);
}
void oo_fit_lsq_nbody(size_t n_traj, const uint64_t *traj_offset, const double *mjd_tt, const double *ra,
                      const double *dec, const double *sigma_ra, const double *sigma_dec, const double *geo_ecl,
                      const oo_ephem_table *tab, const oo_lsq_config *cfg, const oo_iod_result *iod,
                      const oo_perturber *pert, size_t n_pert, double atol, double rtol, oo_lsq_result *out,
                      oo_obs_fit_data *fit, int n_threads) {
  lsq_ctx c = {traj_offset, mjd_tt, ra, dec, sigma_ra, sigma_dec, geo_ecl, tab, cfg, iod, out, fit, pert, n_pert, atol, rtol};
  parallel_for((long long)n_traj, 1, n_threads, lsq_task, &c);
}
