// outfit_b200.hpp -- header-only C++17 host mirror of the reference's interface for the hot path,
// over the C-ABI of include/outfit_b200.h (no CUDA or torch types; link with -loutfit_b200).
//
// Mirrors (names, argument meaning, error behaviour; /root/reference/src):
//   IODParams / IODParamsBuilder::build     initial_orbit_determination/mod.rs:225-344, 360-624
//   FitIOD::fit_full_iod on an ObsDataset   initial_orbit_determination/obs_dataset_api.rs:145-207
//   FitLSQ::fit_lsq on an ObsDataset        differential_orbit_correction/obs_dataset_api.rs:113-190
//   FitOrbitResult / GaussResult            constants.rs:134-175, gauss_result.rs:99-102
//   kepler::propagate_universal             kepler/propagation.rs:114-174
//   OrbitalElements::compute::<Combined>    ephemeris/mod.rs:189-292
// Argument / device failures throw outfit::Error (the reference's outer `Err`); per-trajectory
// failures are VALUES inside the returned vector, like the `Err(..)` entries of FullOrbitResult.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/outfit_b200.h"

namespace outfit {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error("outfit_b200 error " + std::to_string(c) + ": " + m), code(c) {}
};

// IODParamsBuilder (mod.rs:360-624): chained setters with the reference's names, build() validates.
class IODParamsBuilder {
 public:
  IODParamsBuilder() { outfit_b200_iod_params_default(&p_); }
  static IODParamsBuilder from_params(const OutfitIodParams &p) { IODParamsBuilder b; b.p_ = p; return b; }
#define OUTFIT_SETTER(name, type) \
  IODParamsBuilder &name(type v) { p_.name = v; return *this; }
  OUTFIT_SETTER(n_noise_realizations, uint64_t) OUTFIT_SETTER(noise_scale, double) OUTFIT_SETTER(extf, double)
  OUTFIT_SETTER(dtmax, double) OUTFIT_SETTER(dt_min, double) OUTFIT_SETTER(dt_max_triplet, double)
  OUTFIT_SETTER(optimal_interval_time, double) OUTFIT_SETTER(max_obs_for_triplets, uint64_t)
  OUTFIT_SETTER(max_triplets, uint32_t) OUTFIT_SETTER(gap_max, double) OUTFIT_SETTER(max_ecc, double)
  OUTFIT_SETTER(max_perihelion_au, double) OUTFIT_SETTER(min_rho2_au, double) OUTFIT_SETTER(r2_min_au, double)
  OUTFIT_SETTER(r2_max_au, double) OUTFIT_SETTER(aberth_max_iter, uint32_t) OUTFIT_SETTER(aberth_eps, double)
  OUTFIT_SETTER(kepler_eps, double) OUTFIT_SETTER(max_tested_solutions, uint64_t) OUTFIT_SETTER(newton_eps, double)
  OUTFIT_SETTER(newton_max_it, uint64_t) OUTFIT_SETTER(root_imag_eps, double)
#undef OUTFIT_SETTER
  // Err(OutfitError::InvalidIODParameter) -> throws
  OutfitIodParams build() const {
    const int rc = outfit_b200_iod_params_validate(&p_);
    if (rc != OUTFIT_OK) throw Error(rc, "InvalidIODParameter");
    return p_;
  }

 private:
  OutfitIodParams p_;
};

// One observation after the error model (what `Observation` + `OutfitCache` hold for the path).
struct Observation {
  double mjd_tt, ra, dec, sigma_ra, sigma_dec;
  double body_fixed[3];  // Earth-fixed observer position, AU (observer_extension.rs:159-171)
  double mjd_ut1;        // epoch.to_ut1(provider).to_mjd_tai_days() (observer_extension.rs:191-192)
};

// Flattens trajectories into the SoA OutfitObsBatch: every trajectory sorted by mjd_tt with the
// total order of f64::total_cmp (obs_dataset_api.rs:222-223), 3-vectors plane-major.
class ObsBatchBuilder {
 public:
  void add_trajectory(std::vector<Observation> obs) {
    std::stable_sort(obs.begin(), obs.end(), [](const Observation &a, const Observation &b) { return total_less(a.mjd_tt, b.mjd_tt); });
    for (const Observation &o : obs) rows_.push_back(o);
    offsets_.push_back(rows_.size());
  }
  // noise_z: [n_traj][max_triplets][n_noise][6] standard normal deviates in draw order, or empty
  void set_noise(std::vector<double> z) { noise_ = std::move(z); }
  // per-trajectory SmallRng seeds (base_seed ^ traj_id.stable_hash()): deviates generated on the device
  void set_seeds(std::vector<uint64_t> seeds) { seeds_ = std::move(seeds); }
  size_t n_traj() const { return offsets_.size() - 1; }
  size_t n_obs() const { return rows_.size(); }
  // The returned struct points into this builder: keep it alive during the call.
  OutfitObsBatch finish() {
    const size_t n = rows_.size();
    mjd_.resize(n); ra_.resize(n); dec_.resize(n); sra_.resize(n); sdec_.resize(n); ut1_.resize(n); bf_.resize(3 * n);
    uint64_t longest = 0;
    for (size_t i = 0; i < n; ++i) {
      const Observation &o = rows_[i];
      mjd_[i] = o.mjd_tt; ra_[i] = o.ra; dec_[i] = o.dec; sra_[i] = o.sigma_ra; sdec_[i] = o.sigma_dec; ut1_[i] = o.mjd_ut1;
      for (int c = 0; c < 3; ++c) bf_[c * n + i] = o.body_fixed[c];
    }
    for (size_t t = 0; t + 1 < offsets_.size(); ++t) longest = std::max<uint64_t>(longest, offsets_[t + 1] - offsets_[t]);
    OutfitObsBatch b;
    std::memset(&b, 0, sizeof b);
    b.n_traj = offsets_.size() - 1; b.n_obs = n; b.traj_offset = offsets_.data();
    b.mjd_tt = mjd_.data(); b.ra = ra_.data(); b.dec = dec_.data(); b.sigma_ra = sra_.data(); b.sigma_dec = sdec_.data();
    b.observer_body_fixed = bf_.data(); b.mjd_ut1 = ut1_.data();
    b.noise_z = noise_.empty() ? nullptr : noise_.data();
    b.traj_seed = seeds_.empty() ? nullptr : seeds_.data();
    b.max_obs_per_traj = longest;
    return b;
  }

 private:
  static bool total_less(double a, double b) {  // f64::total_cmp
    int64_t x, y;
    std::memcpy(&x, &a, 8); std::memcpy(&y, &b, 8);
    x ^= (int64_t)((uint64_t)(x >> 63) >> 1); y ^= (int64_t)((uint64_t)(y >> 63) >> 1);
    return x < y;
  }
  std::vector<Observation> rows_;
  std::vector<uint64_t> offsets_{0}, seeds_;
  std::vector<double> mjd_, ra_, dec_, sra_, sdec_, ut1_, bf_, noise_;
};

// OrbitalElements::{Keplerian, Cometary} with uncertainty / covariance = None
struct OrbitalElements {
  enum Kind { Keplerian = 0, Equinoctial = 1, Cometary = 2 } kind;
  double reference_epoch;
  double elem[6];
};
// FitOrbitResult::IODGauss((GaussResult, rms)) or the error variant (as OUTFIT_ST_* codes)
struct FitOrbitResult {
  bool ok;
  bool corrected;          // GaussResult::CorrectedOrbit vs PrelimOrbit
  OrbitalElements orbit;
  double rms;
  int error;               // OUTFIT_ST_NO_FEASIBLE_TRIPLETS | NO_VIABLE_ORBIT | INVALID_CONVERSION | INVALID_ORBIT
  int cause;               // NoViableOrbit.cause
  double cause_value;      // NonFiniteScore payload
  uint64_t attempts;       // NoViableOrbit.attempts
  double span;             // NoFeasibleTriplets.span
  uint32_t triplet[3];
  uint32_t realization;
};

// FitOrbitResult::DifferentialCorrection((OrbitalElements::Equinoctial{elements, uncertainty, covariance}, rms))
// (differential_orbit_correction/mod.rs:60-115), or -- when the loop failed -- the IOD orbit it started from
struct LsqOrbitResult {
  bool ok;                 // an orbit is present (corrected or IOD fallback)
  bool corrected;          // true: DifferentialCorrection; false with ok: the IOD orbit was returned (mod.rs:113)
  int error;               // !ok: the IOD / conversion error (OUTFIT_ST_*)
  int fallback_cause;      // ok && !corrected: OUTFIT_ST_LSQ_INVERSION | LSQ_BIZARRE | LSQ_DIVERGED
  double reference_epoch;
  double elem[6];          // corrected: equinoctial (a, h, k, p, q, lambda)
  double sigma[6];         // EquinoctialUncertainty
  double covariance[36];   // column-major 6x6
  double normal_matrix[36];
  double normalised_rms;
  uint64_t total_newton_iterations, num_measurements;
};

class Context {
 public:
  explicit Context(int device = -1) {
    const int rc = outfit_b200_init(device, &h_);
    if (rc != OUTFIT_OK) throw Error(rc, outfit_b200_strerror(rc));
  }
  ~Context() { outfit_b200_destroy(h_); }
  Context(const Context &) = delete;
  Context &operator=(const Context &) = delete;
  OutfitCtx *raw() const { return h_; }

  // &JPLEphem: the Chebyshev blocks of EMB, Moon and Sun (jpl_ephem/mod.rs:145-174)
  void load_ephemeris(const double *cheb, size_t n_blocks, size_t block_stride, double jd_start, double block_days,
                      const uint32_t ipt[3][3], double emrat) {
    check(outfit_b200_load_ephemeris(h_, cheb, n_blocks, block_stride, jd_start, block_days, ipt, emrat));
  }

  // FitIOD::fit_full_iod: one result per trajectory, in batch order
  std::vector<FitOrbitResult> fit_full_iod(const OutfitObsBatch &batch, const OutfitIodParams &params) {
    std::vector<OutfitIodResult> raw(batch.n_traj);
    check(outfit_b200_fit_full_iod(h_, &params, &batch, raw.data()));
    return from_raw(raw);
  }
  // FitIOD::fit_iod (obs_dataset_api.rs:118-143): one trajectory of the batch
  FitOrbitResult fit_iod(const OutfitObsBatch &batch, const OutfitIodParams &params, uint64_t traj_index) {
    std::vector<OutfitIodResult> raw(1);
    check(outfit_b200_fit_iod(h_, &params, &batch, traj_index, raw.data()));
    return from_raw(raw)[0];
  }
  // EphemerisConfig (ephemeris/mod.rs:124-142): aberration OUTFIT_ABERRATION_FIRST | _SECOND, two-body propagator
  void set_ephemeris_config(const OutfitEphemerisConfig &c) { check(outfit_b200_set_ephemeris_config(h_, &c)); }
  static std::vector<FitOrbitResult> from_raw(const std::vector<OutfitIodResult> &raw) {
    std::vector<FitOrbitResult> out(raw.size());
    for (size_t t = 0; t < raw.size(); ++t) {
      const OutfitIodResult &r = raw[t];
      FitOrbitResult &f = out[t];
      f.ok = r.status == OUTFIT_ST_OK;
      f.corrected = r.corrected != 0;
      f.orbit.kind = (OrbitalElements::Kind)r.element_kind;
      f.orbit.reference_epoch = r.epoch;
      std::memcpy(f.orbit.elem, r.elem, sizeof r.elem);
      f.rms = r.rms; f.error = r.status; f.cause = r.cause; f.cause_value = r.cause_value;
      f.attempts = r.attempts; f.span = r.span;
      std::memcpy(f.triplet, r.triplet_idx, sizeof r.triplet_idx);
      f.realization = r.realization;
    }
    return out;
  }

  // FitLSQ::fit_lsq (differential_orbit_correction/obs_dataset_api.rs:113-190): `initial_orbits` = the raw IOD
  // records of the same batch, or nullptr to run the IOD first with `params`.  fit (optional): per-observation
  // residuals / chi / selection flags after the fit.
  std::vector<LsqOrbitResult> fit_lsq(const OutfitObsBatch &batch, const OutfitIodParams &params,
                                      const OutfitLsqConfig &cfg, const OutfitIodResult *initial_orbits = nullptr,
                                      std::vector<OutfitObsFit> *fit = nullptr) {
    std::vector<OutfitLsqResult> raw(batch.n_traj);
    if (fit) fit->resize(batch.n_obs);
    check(outfit_b200_fit_lsq(h_, &params, &cfg, &batch, initial_orbits, raw.data(), fit ? fit->data() : nullptr));
    return lsq_records(raw);
  }
  // The same with DifferentialCorrectionConfig::propagator = PropagatorKind::NBody(nbody) (diff_cor.rs:160-173):
  // gm[n_perturbers], perturber_pos[n_perturbers][3][n_traj] = the perturbers at the epochs of `initial_orbits`
  // (build_perturber_snapshots, propagator/nbody.rs:453-476).
  std::vector<LsqOrbitResult> fit_lsq_nbody(const OutfitObsBatch &batch, const OutfitLsqConfig &cfg, const OutfitNBodyConfig &nbody,
                                            const double *gm, const double *perturber_pos, const OutfitIodResult *initial_orbits,
                                            std::vector<OutfitObsFit> *fit = nullptr) {
    std::vector<OutfitLsqResult> raw(batch.n_traj);
    if (fit) fit->resize(batch.n_obs);
    check(outfit_b200_fit_lsq_nbody(h_, &cfg, &nbody, gm, perturber_pos, &batch, initial_orbits, raw.data(),
                                    fit ? fit->data() : nullptr));
    return lsq_records(raw);
  }
  static std::vector<LsqOrbitResult> lsq_records(const std::vector<OutfitLsqResult> &raw) {
    std::vector<LsqOrbitResult> out(raw.size());
    for (size_t t = 0; t < raw.size(); ++t) {
      const OutfitLsqResult &r = raw[t];
      LsqOrbitResult &f = out[t];
      f.ok = r.kind != OUTFIT_LSQ_NONE;
      f.corrected = r.kind == OUTFIT_LSQ_CORRECTED;
      f.error = r.status; f.fallback_cause = r.fallback_cause;
      f.reference_epoch = r.epoch;
      std::memcpy(f.elem, r.elem, sizeof r.elem);
      std::memcpy(f.sigma, r.sigma, sizeof r.sigma);
      std::memcpy(f.covariance, r.covariance, sizeof r.covariance);
      std::memcpy(f.normal_matrix, r.normal_matrix, sizeof r.normal_matrix);
      f.normalised_rms = r.normalised_rms;
      f.total_newton_iterations = r.total_newton_iterations; f.num_measurements = r.num_measurements;
    }
    return out;
  }
  static OutfitLsqConfig default_lsq_config() {  // DifferentialCorrectionConfig::default()
    OutfitLsqConfig c;
    outfit_b200_lsq_config_default(&c);
    return c;
  }

  // kepler::propagate_universal over n states: rv [6][n], out [11][n] (r1, v1, f, g, fdot, gdot, psi)
  void propagate_universal(size_t n, const double *rv, const double *t0, const double *t1, const OutfitSolverType &solver,
                           double *out, int32_t *status, const double *psi_guess = nullptr) {
    check(outfit_b200_propagate_universal(h_, n, rv, t0, t1, psi_guess, &solver, out, status));
  }

  // OrbitalElements::compute::<Combined> for many orbits x the epochs of one observer
  void ephemeris_twobody(const std::vector<OrbitalElements> &orbits, const std::vector<double> &mjd_tt,
                         const std::vector<double> &mjd_ut1, const double body_fixed[3], std::vector<double> &out,
                         std::vector<int32_t> &status) {
    const size_t n = orbits.size(), E = mjd_tt.size();
    std::vector<int32_t> kind(n);
    std::vector<double> epoch(n), elem(6 * n);
    for (size_t i = 0; i < n; ++i) {
      kind[i] = (int32_t)orbits[i].kind; epoch[i] = orbits[i].reference_epoch;
      for (int q = 0; q < 6; ++q) elem[q * n + i] = orbits[i].elem[q];
    }
    out.assign(9 * E * n, 0.0);
    status.assign(E * n, 0);
    check(outfit_b200_ephemeris_twobody(h_, n, kind.data(), epoch.data(), elem.data(), E, mjd_tt.data(), mjd_ut1.data(),
                                        body_fixed, out.data(), status.data()));
  }

 private:
  void check(int rc) {
    if (rc != OUTFIT_OK) {
      const char *m = outfit_b200_last_error(h_);
      throw Error(rc, (m && *m) ? m : outfit_b200_strerror(rc));
    }
  }
  OutfitCtx *h_ = nullptr;
};

// Every GPU of the box behind one call: the host side of FitIOD::fit_full_iod_parallel (obs_dataset_api.rs:175-207).
// The library cuts the batch into contiguous trajectory ranges of near-equal estimated work, runs one host thread and
// one context per GPU and writes every record at its global trajectory index; results do not depend on the GPU count.
class Group {
 public:
  // n_gpus <= 0: all visible devices; device_ids may repeat an id (several contexts on one GPU)
  explicit Group(int n_gpus = 0, const int *device_ids = nullptr) {
    const int rc = outfit_b200_init_multi(n_gpus, device_ids, &g_);
    if (rc != OUTFIT_OK) throw Error(rc, outfit_b200_strerror(rc));
  }
  ~Group() { outfit_b200_group_destroy(g_); }
  Group(const Group &) = delete;
  Group &operator=(const Group &) = delete;
  int size() const { return outfit_b200_group_size(g_); }
  void load_ephemeris(const double *cheb, size_t n_blocks, size_t block_stride, double jd_start, double block_days,
                      const uint32_t ipt[3][3], double emrat) {
    check(outfit_b200_group_load_ephemeris(g_, cheb, n_blocks, block_stride, jd_start, block_days, ipt, emrat));
  }
  std::vector<FitOrbitResult> fit_full_iod(const OutfitObsBatch &batch, const OutfitIodParams &params) {
    std::vector<OutfitIodResult> raw(batch.n_traj);
    check(outfit_b200_group_fit_full_iod(g_, &params, &batch, raw.data()));
    return Context::from_raw(raw);
  }
  // FitLSQ::fit_lsq with PropagatorKind::NBody over every GPU of the group (raw records, batch order)
  std::vector<OutfitLsqResult> fit_lsq_nbody(const OutfitObsBatch &batch, const OutfitLsqConfig &cfg, const OutfitNBodyConfig &nbody,
                                             const double *gm, const double *perturber_pos, const OutfitIodResult *initial_orbits,
                                             std::vector<OutfitObsFit> *fit = nullptr) {
    std::vector<OutfitLsqResult> raw(batch.n_traj);
    if (fit) fit->resize(batch.n_obs);
    check(outfit_b200_group_fit_lsq_nbody(g_, &cfg, &nbody, gm, perturber_pos, &batch, initial_orbits, raw.data(),
                                          fit ? fit->data() : nullptr));
    return raw;
  }
  void propagate_universal(size_t n, const double *rv, const double *t0, const double *t1, const OutfitSolverType &solver,
                           double *out, int32_t *status, const double *psi_guess = nullptr) {
    check(outfit_b200_group_propagate_universal(g_, n, rv, t0, t1, psi_guess, &solver, out, status));
  }
  // the cut of the last call and every shard's wall time (load imbalance)
  void last_shards(std::vector<uint64_t> &cuts, std::vector<float> &ms) {
    cuts.assign((size_t)size() + 1, 0);
    ms.assign((size_t)size(), 0.f);
    check(outfit_b200_group_last_shards(g_, cuts.data(), ms.data()));
  }

 private:
  void check(int rc) {
    if (rc != OUTFIT_OK) {
      const char *m = outfit_b200_group_last_error(g_);
      throw Error(rc, (m && *m) ? m : outfit_b200_strerror(rc));
    }
  }
  OutfitGroup *g_ = nullptr;
};

}  // namespace outfit
