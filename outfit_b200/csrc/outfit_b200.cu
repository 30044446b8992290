// outfit_b200.cu -- host side of the C-ABI (include/outfit_b200.h) of the B200-native batched IOD path: contexts,
// arenas, streams, launch configuration, the host-buffer entries and the multi-GPU group.
//
// Kernels (all scalar FP64, sm_100a) live in the headers included below:
//   k_iod.cuh         triplets / roots / correct / score / select: the full-IOD pipeline (one warp per trajectory
//                     for selection + fold, one lane per (triplet, realization))
//   dev_geometry.cuh  scorer_observer_kernel, observer_cache_kernel: Chebyshev Earth position + frame rotations
//   dev_ephemeris.cuh two-body Combined ephemeris (first / second-order aberration)
//   k_bulk.cuh        propagate_universal_kernel, arithmetic self-test, fp64_peak_kernel
//   k_lsq.cuh         lsq_quad_kernel (FitLSQ, four lanes per trajectory)
//   k_lsq_nbody.cuh   FitLSQ with PropagatorKind::NBody: init / partials / step kernels driven trip by trip
// There is no CPU fallback anywhere in this library.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/outfit_b200.h"
#include "dev_iod.cuh"
#include "dev_correct.cuh"
#include "dev_geometry.cuh"
#include "dev_ephemeris.cuh"
#include "dev_rng.cuh"
#include "dev_lsq.cuh"

using namespace ofb;

#include "k_iod.cuh"
#include "k_bulk.cuh"
#include "k_lsq.cuh"
#include "k_nbody.cuh"
#include "k_lsq_nbody.cuh"

// =================================================================================================
// context + C-ABI
// =================================================================================================
struct OutfitCtx {
  std::recursive_mutex mu;  // one call at a time per context (the scratch, arena and streams are per context)
  int device = 0;
  int sm_count = 0;
  std::string last_error;
  EphemDev eph{};
  double *d_cheb = nullptr;
  bool have_eph = false;
  unsigned long long *d_counters = nullptr;  // [0] trajectory fetch counter, [1..] work counters
  // scratch owned by the context (grown on demand)
  void *scratch = nullptr;
  size_t scratch_bytes = 0;
  void *h_scratch = nullptr;  // page-locked staging (re-based offsets of a trajectory range)
  size_t h_scratch_bytes = 0;
  void *iod_scratch = nullptr;  // per-candidate arrays of the phase pipeline
  size_t iod_scratch_bytes = 0;
  // host entry point: cached input arena, copy / compute streams, per-slice copy events
  unsigned char *arena = nullptr;
  size_t arena_bytes = 0;
  cudaStream_t copy_stream = nullptr, compute_stream = nullptr, d2h_stream = nullptr;
  cudaEvent_t ring_ev[9] = {};  // bulk host entries: (uploaded, computed, downloaded) per ring slot
  std::vector<cudaEvent_t> copy_ev;
  double *d_zig = nullptr;  // ziggurat tables x[257], f[257] of the on-device StandardNormal (dev_rng.cuh)
  bool triplets_per_thread = true;  // OUTFIT_B200_TRIPLETS_WARP=1: the warp-per-trajectory selection kernel
  int n_streams = 8;  // passes in flight (outfit_b200_set_pass_streams; 1 = one pass on the caller's stream)
  cudaStream_t aux_stream[7] = {};  // extra compute streams (pass overlap)
  cudaEvent_t fork_ev = nullptr, join_ev[7] = {};
  // CUDA events bracketing every phase of the last full-IOD launch (outfit_b200_last_iod_phase_ms)
  std::vector<cudaEvent_t> phase_ev;
  unsigned phase_chunks = 0;
  unsigned phase_observer_kernels = 0;
  bool phase_valid = false;
  bool count_work = true;  // work counters on (outfit_b200_set_work_counters)
  int aberration_order = 1;  // EphemerisConfig::aberration (outfit_b200_set_ephemeris_config)
  void *nbody_state = nullptr;  // propagated states of the N-body ephemeris [6][E][n] + status
  size_t nbody_state_bytes = 0;
};

static int fail(OutfitCtx *ctx, int code, const char *what, cudaError_t e = cudaSuccess) {
  if (ctx) {
    ctx->last_error = what;
    if (e != cudaSuccess) { ctx->last_error += ": "; ctx->last_error += cudaGetErrorString(e); }
  }
  return code;
}
#define CK(call)                                                        \
  do {                                                                  \
    cudaError_t e__ = (call);                                           \
    if (e__ != cudaSuccess) return fail(ctx, OUTFIT_E_CUDA, #call, e__); \
  } while (0)

extern "C" int outfit_b200_abi_version(void) { return OUTFIT_B200_ABI_VERSION; }

extern "C" const char *outfit_b200_strerror(int code) {
  switch (code) {
    case OUTFIT_OK: return "ok";
    case OUTFIT_E_INVALID_ARGUMENT: return "invalid argument";
    case OUTFIT_E_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    case OUTFIT_E_CUDA: return "CUDA runtime error";
    case OUTFIT_E_ALLOC: return "allocation failed";
    case OUTFIT_E_INVALID_IOD_PARAMETER: return "invalid IOD parameter";
    case OUTFIT_E_NO_EPHEMERIS: return "ephemeris table not loaded";
    case OUTFIT_E_UNSUPPORTED: return "size exceeds a kernel limit";
    default: return "unknown error";
  }
}
extern "C" const char *outfit_b200_last_error(OutfitCtx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }

extern "C" void outfit_b200_iod_params_default(OutfitIodParams *p) {
  memset(p, 0, sizeof *p);
  p->n_noise_realizations = 20; p->noise_scale = 1.0; p->extf = -1.0; p->dtmax = 30.0;
  p->dt_min = 0.03; p->dt_max_triplet = 150.0; p->optimal_interval_time = 20.0;
  p->max_obs_for_triplets = 100; p->max_triplets = 10; p->gap_max = 8.0 / 24.0;
  p->max_ecc = 5.0; p->max_perihelion_au = 1.0e3; p->min_rho2_au = 0.01;
  p->aberth_max_iter = 50; p->aberth_eps = 1.0e-6; p->kepler_eps = 1e3 * 2.220446049250313e-16;
  p->max_tested_solutions = 3; p->r2_min_au = 0.05; p->r2_max_au = 200.0;
  p->newton_eps = 1.0e-10; p->newton_max_it = 50; p->root_imag_eps = 1.0e-6;
}
extern "C" int outfit_b200_iod_params_validate(const OutfitIodParams *p) {
  if (!p) return OUTFIT_E_INVALID_ARGUMENT;
  const bool ok = p->noise_scale >= 0.0 && p->dt_min >= 0.0 && p->dt_max_triplet >= 0.0 && p->dtmax >= 0.0 &&
                  p->max_ecc >= 0.0 && p->root_imag_eps >= 0.0 && p->max_perihelion_au > 0.0 &&
                  p->min_rho2_au > 0.0 && p->aberth_eps > 0.0 && p->kepler_eps > 0.0 && p->newton_eps > 0.0 &&
                  p->newton_max_it != 0 && p->aberth_max_iter != 0 && p->max_tested_solutions >= 1 &&
                  p->r2_min_au > 0.0 && p->r2_max_au > 0.0 && p->r2_min_au <= p->r2_max_au;
  return ok ? OUTFIT_OK : OUTFIT_E_INVALID_IOD_PARAMETER;
}
extern "C" void outfit_b200_solver_type_default(OutfitSolverType *s) {
  s->kind = OUTFIT_SOLVER_NEWTON; s->parabolic_method = 0;
  s->convergency = 100.0 * 2.220446049250313e-16; s->max_iter_prelim_kepuni = 20;
}

extern "C" int outfit_b200_init(int device, OutfitCtx **out) {
  if (!out) return OUTFIT_E_INVALID_ARGUMENT;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return OUTFIT_E_NO_DEVICE;
  OutfitCtx *ctx = new (std::nothrow) OutfitCtx();
  if (!ctx) return OUTFIT_E_ALLOC;
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) { delete ctx; return OUTFIT_E_CUDA; }
  }
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return OUTFIT_E_CUDA; }
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return OUTFIT_E_CUDA; }
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaMalloc(&ctx->d_counters, 32 * sizeof(unsigned long long)) != cudaSuccess) { delete ctx; return OUTFIT_E_ALLOC; }
  cudaMemset(ctx->d_counters, 0, 32 * sizeof(unsigned long long));
  // Aberth starting directions from the HOST libm (see dev_gauss.cuh)
  double dir[16];
  for (int k = 0; k < 8; ++k) {
    const double theta = (6.283185307179586476925286766559 / 8.0) * (double)k + (3.14159265358979323846 / 2.0) / 8.0;
    dir[2 * k] = cos(theta);
    dir[2 * k + 1] = sin(theta);
  }
  if (cudaMemcpyToSymbol(c_aberth_dir, dir, sizeof dir) != cudaSuccess) { cudaFree(ctx->d_counters); delete ctx; return OUTFIT_E_CUDA; }
  // correctly rounded reciprocals of the Stumpff-series denominators (see div_by_const)
  double rcp[2 * kSeriesTable];
  for (int j = 0; j < kSeriesTable; ++j) {
    const double d = 3.0 + 2.0 * j;
    rcp[2 * j] = 1.0 / (d * (d + 1.0));
    rcp[2 * j + 1] = 1.0 / ((d + 1.0) * (d + 2.0));
  }
  if (cudaMemcpyToSymbol(c_series_rcp, rcp, sizeof rcp) != cudaSuccess) { cudaFree(ctx->d_counters); delete ctx; return OUTFIT_E_CUDA; }
  {
    // ziggurat tables from the published recurrence (rand_distr's ziggurat_tables.py): x[0] = V / f(R),
    // x[1] = R, x[i] = f^-1(V / x[i-1] + f(x[i-1])), x[256] = 0; f(x) = exp(-x^2 / 2)
    double zig[2 * (kZigN + 1)];
    double *xt = zig, *ft = zig + kZigN + 1;
    xt[0] = kZigV / exp(-kZigR * kZigR / 2.0);
    xt[1] = kZigR;
    for (int i = 2; i < kZigN; ++i) {
      const double last = xt[i - 1];
      xt[i] = sqrt(-2.0 * log(kZigV / last + exp(-last * last / 2.0)));
    }
    xt[kZigN] = 0.0;
    for (int i = 0; i <= kZigN; ++i) ft[i] = exp(-xt[i] * xt[i] / 2.0);
    if (cudaMalloc(&ctx->d_zig, sizeof zig) != cudaSuccess || cudaMemcpy(ctx->d_zig, zig, sizeof zig, cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaFree(ctx->d_counters);
      delete ctx;
      return OUTFIT_E_ALLOC;
    }
  }
  if (const char *ev = getenv("OUTFIT_B200_TRIPLETS_WARP")) ctx->triplets_per_thread = atoi(ev) == 0;
  if (const char *ev = getenv("OUTFIT_B200_STREAMS")) {
    const int v = atoi(ev);
    if (v >= 1 && v <= 8) ctx->n_streams = v;
  }
  *out = ctx;
  return OUTFIT_OK;
}

extern "C" int outfit_b200_set_pass_streams(OutfitCtx *ctx, int n_streams) {
  if (!ctx || n_streams < 1 || n_streams > 8) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  ctx->n_streams = n_streams;
  return OUTFIT_OK;
}

extern "C" void outfit_b200_destroy(OutfitCtx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->d_cheb) cudaFree(ctx->d_cheb);
  if (ctx->d_counters) cudaFree(ctx->d_counters);
  if (ctx->d_zig) cudaFree(ctx->d_zig);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->iod_scratch) cudaFree(ctx->iod_scratch);
  if (ctx->nbody_state) cudaFree(ctx->nbody_state);
  if (ctx->h_scratch) cudaFreeHost(ctx->h_scratch);
  for (cudaEvent_t e : ctx->phase_ev) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->copy_ev) cudaEventDestroy(e);
  if (ctx->arena) cudaFree(ctx->arena);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->compute_stream) cudaStreamDestroy(ctx->compute_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  for (cudaEvent_t e : ctx->ring_ev)
    if (e) cudaEventDestroy(e);
  for (int i = 0; i < 7; ++i) {
    if (ctx->aux_stream[i]) cudaStreamDestroy(ctx->aux_stream[i]);
    if (ctx->join_ev[i]) cudaEventDestroy(ctx->join_ev[i]);
  }
  if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
  delete ctx;
}

extern "C" int outfit_b200_load_ephemeris(OutfitCtx *ctx, const double *cheb, size_t n_blocks,
                                          size_t block_stride, double jd_start, double block_days,
                                          const uint32_t ipt[3][3], double emrat) {
  if (!ctx || !cheb || n_blocks == 0 || block_stride == 0 || !(block_days > 0.0)) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  for (int b = 0; b < 3; ++b) {
    if (ipt[b][1] < 3 || ipt[b][1] > kMaxCheb || ipt[b][2] == 0) return fail(ctx, OUTFIT_E_UNSUPPORTED, "ipt: n_coeff must be in [3, 18]");
    if ((size_t)ipt[b][0] + (size_t)ipt[b][1] * ipt[b][2] * 3 > block_stride) return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "ipt exceeds block_stride");
  }
  CK(cudaSetDevice(ctx->device));
  if (ctx->d_cheb) { cudaFree(ctx->d_cheb); ctx->d_cheb = nullptr; }
  CK(cudaMalloc(&ctx->d_cheb, n_blocks * block_stride * sizeof(double)));
  CK(cudaMemcpy(ctx->d_cheb, cheb, n_blocks * block_stride * sizeof(double), cudaMemcpyHostToDevice));
  ctx->eph.cheb = ctx->d_cheb;
  ctx->eph.n_blocks = n_blocks;
  ctx->eph.block_stride = block_stride;
  ctx->eph.jd_start = jd_start;
  ctx->eph.jd_end = jd_start + block_days * (double)n_blocks;
  ctx->eph.block_days = block_days;
  for (int b = 0; b < 3; ++b)
    for (int j = 0; j < 3; ++j) ctx->eph.ipt[b][j] = ipt[b][j];
  ctx->eph.emrat = emrat;
  ctx->have_eph = true;
  return OUTFIT_OK;
}

static int ensure_scratch(OutfitCtx *ctx, size_t bytes) {
  if (ctx->scratch_bytes >= bytes) return OUTFIT_OK;
  if (ctx->scratch) { cudaFree(ctx->scratch); ctx->scratch = nullptr; ctx->scratch_bytes = 0; }
  if (cudaMalloc(&ctx->scratch, bytes) != cudaSuccess) return fail(ctx, OUTFIT_E_ALLOC, "cudaMalloc(scratch)");
  ctx->scratch_bytes = bytes;
  return OUTFIT_OK;
}

static int ensure_iod_scratch(OutfitCtx *ctx, size_t bytes) {
  if (ctx->iod_scratch_bytes >= bytes) return OUTFIT_OK;
  if (ctx->iod_scratch) { cudaFree(ctx->iod_scratch); ctx->iod_scratch = nullptr; ctx->iod_scratch_bytes = 0; }
  if (cudaMalloc(&ctx->iod_scratch, bytes) != cudaSuccess) return fail(ctx, OUTFIT_E_ALLOC, "cudaMalloc(candidate scratch)");
  ctx->iod_scratch_bytes = bytes;
  return OUTFIT_OK;
}

static IodDevParams to_dev_params(const OutfitIodParams &p) {
  IodDevParams d;
  d.noise_scale = p.noise_scale; d.extf = p.extf; d.dtmax = p.dtmax; d.dt_min = p.dt_min;
  d.dt_max_triplet = p.dt_max_triplet; d.inv_optimal_interval = 1.0 / p.optimal_interval_time;
  d.max_ecc = p.max_ecc; d.max_perihelion_au = p.max_perihelion_au; d.min_rho2_au = p.min_rho2_au;
  d.aberth_eps = p.aberth_eps; d.kepler_eps = p.kepler_eps; d.r2_min_au = p.r2_min_au; d.r2_max_au = p.r2_max_au;
  d.newton_eps = p.newton_eps; d.root_imag_eps = p.root_imag_eps;
  d.n_noise = (unsigned)p.n_noise_realizations; d.max_triplets = p.max_triplets;
  d.max_obs_for_triplets = p.max_obs_for_triplets > 0xffffffffull ? 0xffffffffu : (unsigned)p.max_obs_for_triplets;
  d.aberth_max_iter = p.aberth_max_iter;
  d.max_tested_solutions = p.max_tested_solutions > 0xffffffffull ? 0xffffffffu : (unsigned)p.max_tested_solutions;
  d.newton_max_it = p.newton_max_it > 0xffffffffull ? 0xffffffffu : (unsigned)p.newton_max_it;
  return d;
}

// max observations per trajectory the shared-memory staging supports (per-warp planes)
constexpr unsigned kMaxObsPerTraj = 448;
constexpr unsigned kMaxTriplets = 1024;

// The host entry copies the noise deviates in trajectory slices; a pass waits for the slice holding its last
// trajectory (events recorded on the copy stream, one per slice).
struct SliceWait {
  const cudaEvent_t *ev;  // [n_slices] or null (nothing to wait for)
  unsigned long long slice;
  size_t n_slices;
  void wait(unsigned long long t0, unsigned long long tn, cudaStream_t s) const {
    if (!ev) return;
    const size_t last = (size_t)((t0 + tn - 1) / slice);
    cudaStreamWaitEvent(s, ev[last < n_slices ? last : n_slices - 1], 0);
  }
};

// Launch the device pipeline on device-resident buffers.  `max_obs` = longest trajectory.
// `max_chunk` (0 = as large as the scratch budget allows) bounds the trajectories per pipeline pass;
// `before_chunk->wait(t0, tn, s)` is called before the kernels of trajectories [t0, t0 + tn) are enqueued on
// stream s (the host entry point uses it to make s wait for that slice of the input copy).
// With `stream2` the passes alternate between the two streams (each with its own candidate scratch):
// a handful of candidates per pass run ~100x longer than the rest (f-g loops whose every Kepler solve
// exhausts its 50 Newton steps -- reference behaviour), and a pass boundary on ONE stream leaves the
// GPU idle until they finish; on two streams the next pass fills the machine meanwhile.
static int launch_iod(OutfitCtx *ctx, const OutfitIodParams *params, const OutfitObsBatch *b,
                      OutfitIodResult *d_out, unsigned max_obs, cudaStream_t stream,
                      unsigned long long max_chunk = 0, const SliceWait *before_chunk = nullptr, int n_streams = 1) {
  if (!ctx->have_eph) return fail(ctx, OUTFIT_E_NO_EPHEMERIS, "outfit_b200_load_ephemeris must be called first");
  if (max_obs > kMaxObsPerTraj) return fail(ctx, OUTFIT_E_UNSUPPORTED, "trajectory longer than 448 observations");
  if (params->max_triplets > kMaxTriplets) return fail(ctx, OUTFIT_E_UNSUPPORTED, "max_triplets > 1024");
  if (params->n_noise_realizations > 0 && !b->noise_z && !b->traj_seed)
    return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "n_noise_realizations > 0 needs noise_z (host-drawn deviates) or traj_seed (on-device generator)");
  const bool device_noise = params->n_noise_realizations > 0 && !b->noise_z;
  if (params->n_noise_realizations > 65535) return fail(ctx, OUTFIT_E_UNSUPPORTED, "n_noise_realizations > 65535");
  const size_t n = b->n_obs;
  const bool have_cache = b->obs_helio_equ && b->obs_geo_ecl;
  const bool have_bf = b->observer_body_fixed && b->mjd_ut1;
  if (!have_cache && !have_bf) return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "need obs_helio_equ+obs_geo_ecl or observer_body_fixed+mjd_ut1");
  // scratch: scorer[3][n] (+ geo[3][n] + helio[3][n] when computed on device) + status[n]
  const size_t planes = have_cache ? 3 : 9;
  int rc = ensure_scratch(ctx, planes * n * sizeof(double) + n * sizeof(int) + 256);
  if (rc) return rc;
  double *d_scorer = reinterpret_cast<double *>(ctx->scratch);
  const double *d_geo = b->obs_geo_ecl;
  const double *d_helio = b->obs_helio_equ;
  int *d_status = reinterpret_cast<int *>(d_scorer + planes * n);
  const int tpb = 128;
  const unsigned gblocks = (unsigned)((n + tpb - 1) / tpb);
  // phase events: [0] start, [1] after the observer kernels, then 5 per chunk (after each phase)
  ctx->phase_valid = false;
  size_t ev_next = 0;
  auto mark = [&]() {
    if (n_streams > 1) return;
    if (ev_next == ctx->phase_ev.size()) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return;
      ctx->phase_ev.push_back(e);
    }
    cudaEventRecord(ctx->phase_ev[ev_next++], stream);
  };
  mark();
  if (!have_cache) {
    double *geo = d_scorer + 3 * n, *helio = d_scorer + 6 * n;
    if (n) observer_cache_kernel<<<gblocks, tpb, 0, stream>>>(ctx->eph, n, b->mjd_tt, b->mjd_ut1, b->observer_body_fixed, geo, helio, d_status);
    d_geo = geo;
    d_helio = helio;
  }
  if (n) scorer_observer_kernel<<<gblocks, tpb, 0, stream>>>(ctx->eph, n, b->mjd_tt, d_geo, d_scorer, d_status);
  mark();

  const IodDevParams P = to_dev_params(*params);
  const unsigned M = P.n_noise + 1;
  const unsigned long long cand_per_traj = (unsigned long long)P.max_triplets * M;
  // per-candidate scratch (bytes): code 4 + nroots 1 + roots 64 + state_kind 4 + state 56 + score 20
  const size_t per_cand = 4 + 1 + 64 + 4 + 56 + 4 + 4 + 8 + 4;
  const size_t per_traj = (size_t)cand_per_traj * per_cand + (size_t)P.max_triplets * 4 + 4 +
                          (device_noise ? (size_t)P.max_triplets * P.n_noise * 48 : 0);
  // candidate scratch per pass: 6 GB for a single stream, 2 GB per stream (16 GB in total) with 8 in flight
  const size_t budget = n_streams > 1 ? (size_t)2 << 30 : (size_t)6 << 30;
  unsigned long long chunk = b->n_traj;
  if (per_traj * chunk > budget) chunk = budget / per_traj ? budget / per_traj : 1;
  if (max_chunk && chunk > max_chunk) chunk = max_chunk;
  constexpr int kMaxStreams = 8;
  const int ns = n_streams < 1 ? 1 : (n_streams > kMaxStreams ? kMaxStreams : n_streams);
  cudaStream_t st[kMaxStreams];
  st[0] = stream;
  for (int i = 1; i < ns; ++i) {
    if (!ctx->aux_stream[i - 1]) CK(cudaStreamCreateWithFlags(&ctx->aux_stream[i - 1], cudaStreamNonBlocking));
    if (!ctx->join_ev[i - 1]) CK(cudaEventCreateWithFlags(&ctx->join_ev[i - 1], cudaEventDisableTiming));
    st[i] = ctx->aux_stream[i - 1];
  }
  const size_t scratch_half = (per_traj * chunk + 64 * 256 + 255) & ~(size_t)255;
  rc = ensure_iod_scratch(ctx, scratch_half * ns);
  if (rc) return rc;
  const unsigned cap = max_obs < 3 ? 4 : ((max_obs + 1) & ~1u);
  size_t per_warp = ((size_t)cap * sizeof(double) + (size_t)P.max_triplets * (8 + 4) + 15) & ~(size_t)15;
  const size_t smem0 = per_warp * kWarpsPerBlock;
  if (smem0 > 200 * 1024) return fail(ctx, OUTFIT_E_UNSUPPORTED, "shared memory per block exceeds 200 KB");
  CK(cudaFuncSetAttribute(triplets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
  CK(cudaMemsetAsync(ctx->d_counters, 0, 32 * sizeof(unsigned long long), stream));
  if (ns > 1) {  // fork: the other streams start after the observer kernels and the counter reset
    if (!ctx->fork_ev) CK(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
    CK(cudaEventRecord(ctx->fork_ev, st[0]));
    for (int i = 1; i < ns; ++i) CK(cudaStreamWaitEvent(st[i], ctx->fork_ev, 0));
  }
  unsigned n_chunks = 0;
  for (unsigned long long t0 = 0; t0 < b->n_traj; t0 += chunk) {
    const unsigned long long tn = b->n_traj - t0 < chunk ? b->n_traj - t0 : chunk;
    cudaStream_t stream = st[n_chunks % ns];  // shadows the argument inside the pass
    if (before_chunk) before_chunk->wait(t0, tn, stream);
    IodBatchDev B;
    B.n_traj = tn; B.n_obs = n;
    B.traj_offset = reinterpret_cast<const unsigned long long *>(b->traj_offset) + t0;
    B.mjd_tt = b->mjd_tt; B.ra = b->ra; B.dec = b->dec; B.sigma_ra = b->sigma_ra; B.sigma_dec = b->sigma_dec;
    B.helio = d_helio; B.scorer = d_scorer; B.obs_status = d_status;
    B.noise_z = b->noise_z ? b->noise_z + (size_t)t0 * P.max_triplets * P.n_noise * 6 : nullptr;
    IodScratch S;
    S.n_cand = tn * cand_per_traj;
    unsigned char *p = reinterpret_cast<unsigned char *>(ctx->iod_scratch) + (n_chunks % ns) * scratch_half;
    auto take = [&](size_t bytes) { void *q = p; p += (bytes + 255) & ~(size_t)255; return q; };
    S.roots = (double *)take(8 * S.n_cand * 8);
    S.state = (double *)take(7 * S.n_cand * 8);
    S.score_sum = (double *)take(S.n_cand * 8);
    S.code = (int *)take(S.n_cand * 4);
    S.state_kind = (int *)take(S.n_cand * 4);
    S.score_kind = (int *)take(S.n_cand * 4);
    S.score_code = (int *)take(S.n_cand * 4);
    S.score_narc = (unsigned *)take(S.n_cand * 4);
    S.trip = (unsigned *)take(tn * P.max_triplets * 4);
    S.ktraj = (unsigned *)take(tn * 4);
    S.nroots = (unsigned char *)take(S.n_cand);
    if (device_noise) {
      // this pass's deviates, generated on the device in draw order (dev_rng.cuh)
      double *nz = (double *)take((size_t)tn * P.max_triplets * P.n_noise * 48);
      noise_kernel<<<(unsigned)((tn + 127) / 128), 128, 0, stream>>>(tn, reinterpret_cast<const unsigned long long *>(b->traj_seed) + t0,
                                                                   ctx->d_zig, P.max_triplets * P.n_noise, nz);
      B.noise_z = nz;
    }
    const unsigned tblocks = (unsigned)((tn + kWarpsPerBlock - 1) / kWarpsPerBlock);
    const unsigned cblocks = (unsigned)((S.n_cand + kCandThreads - 1) / kCandThreads);
    if (S.n_cand == 0) {
      // max_triplets == 0: no candidate exists; select_kernel reports NoFeasibleTriplets like the reference
      CK(cudaMemsetAsync(S.ktraj, 0, tn * 4, stream));
      for (int q = 0; q < 4; ++q) mark();
    } else {
      // thread per trajectory while the block's heap columns (12 B per slot and thread) fit in shared memory:
      // 128 threads up to K = 106, 64 up to 213, 32 up to 426; beyond that the warp-per-trajectory kernel
      unsigned ttpb = 128;
      while (ttpb > 32 && (size_t)P.max_triplets * 12 * ttpb > 160 * 1024) ttpb >>= 1;
      if ((size_t)P.max_triplets * 12 * ttpb <= 160 * 1024 && ctx->triplets_per_thread) {
        const size_t tsm = (size_t)P.max_triplets * 12 * ttpb;
        if (tsm > 48 * 1024) CK(cudaFuncSetAttribute(triplets_thread_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
        triplets_thread_kernel<<<(unsigned)((tn + ttpb - 1) / ttpb), ttpb, tsm, stream>>>(B, P, S);
      } else {
        triplets_kernel<<<tblocks, kWarpsPerBlock * 32, smem0, stream>>>(B, P, S, cap);
      }
      mark();
      roots_kernel<<<cblocks, kCandThreads, 32 * kCandThreads * sizeof(double), stream>>>(B, P, S, ctx->d_counters + 1);
      mark();
      const unsigned kblocks = (unsigned)((S.n_cand + kCorrectThreads - 1) / kCorrectThreads);
      if (ctx->count_work) correct_kernel<true><<<kblocks, kCorrectThreads, kCorrectSmemBytes, stream>>>(B, P, S, ctx->d_counters + 1);
      else correct_kernel<false><<<kblocks, kCorrectThreads, kCorrectSmemBytes, stream>>>(B, P, S, ctx->d_counters + 1);
      mark();
      if (ctx->count_work) score_kernel<true><<<cblocks, kCandThreads, 0, stream>>>(B, P, S, ctx->d_counters + 1);
      else score_kernel<false><<<cblocks, kCandThreads, 0, stream>>>(B, P, S, ctx->d_counters + 1);
      mark();
    }
    select_kernel<<<tblocks, kWarpsPerBlock * 32, 0, stream>>>(B, P, S, d_out + t0);
    mark();
    ++n_chunks;
  }
  for (int i = 1; i < ns; ++i) {  // join
    CK(cudaEventRecord(ctx->join_ev[i - 1], st[i]));
    CK(cudaStreamWaitEvent(st[0], ctx->join_ev[i - 1], 0));
  }
  CK(cudaGetLastError());
  ctx->phase_chunks = n_chunks;
  ctx->phase_observer_kernels = n ? (have_cache ? 1u : 2u) : 0u;
  ctx->phase_valid = ns == 1 && ev_next == 2 + 5 * (size_t)n_chunks;  // per-phase times need ONE stream
  return OUTFIT_OK;
}

extern "C" int outfit_b200_fit_full_iod_device(OutfitCtx *ctx, const OutfitIodParams *params,
                                               const OutfitObsBatch *batch, OutfitIodResult *out,
                                               void *cuda_stream) {
  if (!ctx || !params || !batch || (!out && batch->n_traj)) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  int rc = outfit_b200_iod_params_validate(params);
  if (rc) return fail(ctx, rc, "IODParams validation failed (mod.rs:544-624)");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  // longest trajectory: read the offsets back once (T+1 words); callers that know it can avoid
  // this by using the host entry point, which computes it from the host copy.
  unsigned max_obs = batch->max_obs_per_traj > 0xffffffffull ? 0xffffffffu : (unsigned)batch->max_obs_per_traj;
  if (batch->n_traj && max_obs == 0) {
    unsigned long long *h = (unsigned long long *)malloc((batch->n_traj + 1) * sizeof(unsigned long long));
    if (!h) return fail(ctx, OUTFIT_E_ALLOC, "malloc(offsets)");
    cudaError_t e = cudaMemcpyAsync(h, batch->traj_offset, (batch->n_traj + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { free(h); return fail(ctx, OUTFIT_E_CUDA, "read traj_offset", e); }
    for (unsigned long long t = 0; t < batch->n_traj; ++t) {
      const unsigned long long c = h[t + 1] - h[t];
      if (c > max_obs) max_obs = c > 0xffffffffull ? 0xffffffffu : (unsigned)c;
    }
    free(h);
  }
  // large batches run as n_streams passes in flight (see launch_iod); per-phase timings need n_streams = 1
  const unsigned long long passes = (unsigned long long)ctx->n_streams;
  if (passes > 1 && batch->n_traj >= passes * 1024)
    return launch_iod(ctx, params, batch, out, max_obs, stream, (batch->n_traj + passes - 1) / passes, nullptr, ctx->n_streams);
  return launch_iod(ctx, params, batch, out, max_obs, stream);
}

// (outfit_b200_fit_iod is defined after fit_full_iod_range below)
// ---- host-buffer plumbing shared by the host entry points ---------------------------------------------
// A host entry works on the trajectory range [tb, te) of the caller's batch (the whole batch for the
// single-GPU entries, one shard of it for the multi-GPU group): the per-observation arrays of the range are a
// contiguous slice [o0, o1) of the caller's arrays, the plane-major 3-vectors are three such slices at the
// caller's plane stride (one 2-D copy), and traj_offset is re-based into a page-locked staging buffer.
static int ensure_arena(OutfitCtx *ctx, size_t bytes) {
  if (ctx->arena_bytes >= bytes) return OUTFIT_OK;
  if (ctx->arena) { cudaFree(ctx->arena); ctx->arena = nullptr; ctx->arena_bytes = 0; }
  if (cudaMalloc(&ctx->arena, bytes) != cudaSuccess) return fail(ctx, OUTFIT_E_ALLOC, "cudaMalloc(batch arena)");
  ctx->arena_bytes = bytes;
  return OUTFIT_OK;
}
static int ensure_host_staging(OutfitCtx *ctx, size_t bytes) {
  if (ctx->h_scratch_bytes >= bytes) return OUTFIT_OK;
  if (ctx->h_scratch) { cudaFreeHost(ctx->h_scratch); ctx->h_scratch = nullptr; ctx->h_scratch_bytes = 0; }
  if (cudaMallocHost(&ctx->h_scratch, bytes) != cudaSuccess) return fail(ctx, OUTFIT_E_ALLOC, "cudaMallocHost(staging)");
  ctx->h_scratch_bytes = bytes;
  return OUTFIT_OK;
}
static int ensure_host_streams(OutfitCtx *ctx) {
  if (!ctx->copy_stream) CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  if (!ctx->compute_stream) CK(cudaStreamCreateWithFlags(&ctx->compute_stream, cudaStreamNonBlocking));
  return OUTFIT_OK;
}

// traj_offset monotone within n_obs over [tb, te]; longest trajectory of the range
static int check_offsets(OutfitCtx *ctx, const OutfitObsBatch *hb, size_t tb, size_t te, unsigned *max_obs) {
  unsigned mo = 0;
  for (size_t t = tb; t < te; ++t) {
    if (hb->traj_offset[t + 1] < hb->traj_offset[t] || hb->traj_offset[t + 1] > hb->n_obs)
      return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "traj_offset is not monotone within n_obs");
    const unsigned long long c = hb->traj_offset[t + 1] - hb->traj_offset[t];
    if (c > mo) mo = c > 0xffffffffull ? 0xffffffffu : (unsigned)c;
  }
  if (max_obs) *max_obs = mo;
  return OUTFIT_OK;
}

// carves the arena and enqueues the H2D copies of a trajectory range on `cs`
struct ArenaPut {
  unsigned char *arena;
  size_t off = 0;
  cudaStream_t cs;
  cudaError_t err = cudaSuccess;
  void *raw(size_t nbytes) {
    void *dst = arena + off;
    off += (nbytes + 255) & ~(size_t)255;
    return dst;
  }
  void *put(const void *src, size_t nbytes) {
    void *dst = raw(nbytes);
    if (src && nbytes && err == cudaSuccess) err = cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyHostToDevice, cs);
    return dst;
  }
  // [3][n_all] plane-major source, columns [o0, o0 + n) -> [3][n]
  void *put_planes(const double *src, size_t n_all, size_t o0, size_t n, int planes = 3) {
    void *dst = raw((size_t)planes * n * 8);
    if (src && n && err == cudaSuccess)
      err = cudaMemcpy2DAsync(dst, n * 8, src + o0, n_all * 8, n * 8, (size_t)planes, cudaMemcpyHostToDevice, cs);
    return dst;
  }
};

static size_t iod_arena_bytes(const OutfitIodParams *params, const OutfitObsBatch *hb, size_t T, size_t n) {
  const bool have_cache = hb->obs_helio_equ && hb->obs_geo_ecl;
  const size_t n_noise_doubles = hb->noise_z ? T * (size_t)params->max_triplets * (size_t)params->n_noise_realizations * 6 : 0;
  return (T + 1) * 8 + T * 8 + 5 * n * 8 + (have_cache ? 6 : 4) * n * 8 + n_noise_doubles * 8 + T * sizeof(OutfitIodResult) + 16 * 256;
}

// FitIOD::fit_full_iod on the trajectories [tb, te) of the host batch; out = &results[tb]
static int fit_full_iod_range(OutfitCtx *ctx, const OutfitIodParams *params, const OutfitObsBatch *hb, size_t tb, size_t te,
                              OutfitIodResult *out) {
  const size_t T = te - tb;
  if (T == 0) return OUTFIT_OK;
  unsigned max_obs = 0;
  int rc = check_offsets(ctx, hb, tb, te, &max_obs);
  if (rc) return rc;
  const size_t o0 = hb->traj_offset[tb], n = hb->traj_offset[te] - o0, n_all = hb->n_obs;
  const bool have_cache = hb->obs_helio_equ && hb->obs_geo_ecl;
  const bool have_bf = hb->observer_body_fixed && hb->mjd_ut1;
  if (!have_cache && !have_bf) return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "need obs_helio_equ+obs_geo_ecl or observer_body_fixed+mjd_ut1");
  const size_t per_traj_noise = hb->noise_z ? (size_t)params->max_triplets * (size_t)params->n_noise_realizations * 6 : 0;
  const size_t n_noise_doubles = T * per_traj_noise;
  // one cached device arena for the inputs and the results (every sub-buffer 256-B aligned)
  rc = ensure_arena(ctx, iod_arena_bytes(params, hb, T, n));
  if (rc) return rc;
  rc = ensure_host_streams(ctx);
  if (rc) return rc;
  cudaStream_t cs = ctx->copy_stream, stream = ctx->compute_stream;
  ArenaPut A{ctx->arena, 0, cs};
  OutfitObsBatch db = *hb;
  db.n_traj = T; db.n_obs = n; db.max_obs_per_traj = max_obs;
  if (o0 == 0) {
    db.traj_offset = (const uint64_t *)A.put(hb->traj_offset + tb, (T + 1) * 8);
  } else {  // re-base the offsets of the range
    rc = ensure_host_staging(ctx, (T + 1) * 8);
    if (rc) return rc;
    uint64_t *h = reinterpret_cast<uint64_t *>(ctx->h_scratch);
    for (size_t t = 0; t <= T; ++t) h[t] = hb->traj_offset[tb + t] - o0;
    db.traj_offset = (const uint64_t *)A.put(h, (T + 1) * 8);
  }
  db.mjd_tt = (const double *)A.put(hb->mjd_tt + o0, n * 8);
  db.ra = (const double *)A.put(hb->ra + o0, n * 8);
  db.dec = (const double *)A.put(hb->dec + o0, n * 8);
  db.sigma_ra = (const double *)A.put(hb->sigma_ra + o0, n * 8);
  db.sigma_dec = (const double *)A.put(hb->sigma_dec + o0, n * 8);
  if (have_cache) {
    db.obs_helio_equ = (const double *)A.put_planes(hb->obs_helio_equ, n_all, o0, n);
    db.obs_geo_ecl = (const double *)A.put_planes(hb->obs_geo_ecl, n_all, o0, n);
    db.observer_body_fixed = nullptr; db.mjd_ut1 = nullptr;
  } else {
    db.observer_body_fixed = (const double *)A.put_planes(hb->observer_body_fixed, n_all, o0, n);
    db.mjd_ut1 = (const double *)A.put(hb->mjd_ut1 + o0, n * 8);
    db.obs_helio_equ = nullptr; db.obs_geo_ecl = nullptr;
  }
  // The observation stream is small (88 B per observation); the noise deviates are the bulk of the
  // input (48 B per noisy candidate).  They are copied in trajectory slices on the copy stream while
  // the compute stream works on the slices that already arrived.
  const size_t want = ctx->n_streams > 1 ? (n_noise_doubles ? 2 * (size_t)ctx->n_streams : (size_t)ctx->n_streams) : 1;
  size_t n_slices = T < want * 1024 ? (T + 1023) / 1024 : want;
  if (const char *ev = getenv("OUTFIT_B200_COPY_SLICES")) {  // tuning knob: 1 = no copy/compute overlap
    const long v = atol(ev);
    if (v >= 1) n_slices = (size_t)v < T ? (size_t)v : T;
  }
  const unsigned long long slice = (T + n_slices - 1) / n_slices;
  while (ctx->copy_ev.size() < n_slices + 1) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->copy_ev.push_back(e);
  }
  db.traj_seed = (!n_noise_doubles && hb->traj_seed) ? (const uint64_t *)A.put(hb->traj_seed + tb, T * 8) : nullptr;
  CK(cudaEventRecord(ctx->copy_ev[0], cs));  // observation arrays (and seeds) are in flight up to here
  double *d_noise = nullptr;
  if (n_noise_doubles) {
    d_noise = (double *)A.raw(n_noise_doubles * 8);
    const double *h_noise = hb->noise_z + tb * per_traj_noise;
    for (size_t sidx = 0; sidx < n_slices; ++sidx) {
      const size_t t0 = sidx * slice, t1 = t0 + slice < T ? t0 + slice : T;
      if (t1 > t0 && A.err == cudaSuccess)
        A.err = cudaMemcpyAsync(d_noise + t0 * per_traj_noise, h_noise + t0 * per_traj_noise, (t1 - t0) * per_traj_noise * 8, cudaMemcpyHostToDevice, cs);
      CK(cudaEventRecord(ctx->copy_ev[sidx + 1], cs));
    }
  }
  db.noise_z = d_noise;
  OutfitIodResult *d_out = (OutfitIodResult *)A.raw(T * sizeof(OutfitIodResult));
  if (A.err != cudaSuccess) { cudaStreamSynchronize(cs); return fail(ctx, OUTFIT_E_CUDA, "fit_full_iod: H2D", A.err); }
  CK(cudaStreamWaitEvent(stream, ctx->copy_ev[0], 0));
  const SliceWait wait_slice{n_noise_doubles ? ctx->copy_ev.data() + 1 : nullptr, slice, n_slices};
  rc = launch_iod(ctx, params, &db, d_out, max_obs, stream, n_slices > 1 ? slice : 0, &wait_slice,
                  n_slices > 1 ? ctx->n_streams : 1);
  if (rc == OUTFIT_OK) {
    cudaError_t e = cudaMemcpyAsync(out, d_out, T * sizeof(OutfitIodResult), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
    if (e != cudaSuccess) rc = fail(ctx, OUTFIT_E_CUDA, "fit_full_iod: copy back / kernel", e);
  } else {
    cudaStreamSynchronize(cs);
  }
  return rc;
}

static int check_iod_host_args(OutfitCtx *ctx, const OutfitIodParams *params, const OutfitObsBatch *hb, const void *out) {
  if (!ctx || !params || !hb || (!out && hb->n_traj)) return OUTFIT_E_INVALID_ARGUMENT;
  int rc = outfit_b200_iod_params_validate(params);
  if (rc) return fail(ctx, rc, "IODParams validation failed (mod.rs:544-624)");
  if (hb->n_traj && (!hb->traj_offset || !hb->mjd_tt || !hb->ra || !hb->dec || !hb->sigma_ra || !hb->sigma_dec))
    return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "NULL observation array");
  return OUTFIT_OK;
}

extern "C" int outfit_b200_fit_full_iod(OutfitCtx *ctx, const OutfitIodParams *params,
                                        const OutfitObsBatch *hb, OutfitIodResult *out) {
  if (!ctx) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  int rc = check_iod_host_args(ctx, params, hb, out);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  return fit_full_iod_range(ctx, params, hb, 0, hb->n_traj, out);
}

extern "C" void outfit_b200_lsq_config_default(OutfitLsqConfig *c) {  // diff_cor.rs:175-192
  if (!c) return;
  memset(c, 0, sizeof *c);
  c->max_newton_iterations = 30;
  c->max_outlier_rejection_passes = 10;
  c->convergence_threshold = 1e-4;
  c->convergence_before_rejection_threshold = 2.0;
  c->rms_stagnation_ratio = 0.98;
  c->rms_divergence_ratio = 1.5;
  c->max_stagnation_iterations = 3;
  c->enable_outlier_rejection = 1;
  c->chi2_rejection_threshold = 25.0;
  c->chi2_recovery_threshold = 9.0;
  c->eccentricity_limit = 1.2;
  c->min_semi_major_axis = 1e-6;
  c->max_semi_major_axis = 1e4;
  c->min_periapsis_distance = 1e-6;
  c->max_apoapsis_distance = 1e4;
  for (int j = 0; j < 6; ++j) c->free_elements[j] = 1;
}

static LsqCfgDev to_lsq_dev(const OutfitLsqConfig &c) {
  LsqCfgDev d;
  d.max_newton_iterations = c.max_newton_iterations;
  d.max_outlier_rejection_passes = c.max_outlier_rejection_passes;
  d.max_stagnation_iterations = c.max_stagnation_iterations;
  d.convergence_threshold = c.convergence_threshold;
  d.convergence_before_rejection_threshold = c.convergence_before_rejection_threshold;
  d.rms_stagnation_ratio = c.rms_stagnation_ratio;
  d.rms_divergence_ratio = c.rms_divergence_ratio;
  d.chi2_reject = c.chi2_rejection_threshold;
  d.chi2_recover = c.chi2_recovery_threshold;
  d.ecc_limit = c.eccentricity_limit;
  d.min_a = c.min_semi_major_axis; d.max_a = c.max_semi_major_axis;
  d.min_q = c.min_periapsis_distance; d.max_Q = c.max_apoapsis_distance;
  d.enable_outlier_rejection = c.enable_outlier_rejection;
  for (int j = 0; j < 6; ++j) d.free_el[j] = c.free_elements[j];
  return d;
}

extern "C" int outfit_b200_fit_lsq_device(OutfitCtx *ctx, const OutfitLsqConfig *cfg, const OutfitObsBatch *b,
                                          const OutfitIodResult *iod, OutfitLsqResult *out, OutfitObsFit *fit,
                                          void *cuda_stream) {
  if (!ctx || !cfg || !b) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (b->n_traj && (!iod || !out || !fit)) return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "fit_lsq: iod, out and fit are required");
  if (!ctx->have_eph) return fail(ctx, OUTFIT_E_NO_EPHEMERIS, "outfit_b200_load_ephemeris must be called first");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  const size_t n = b->n_obs;
  if (b->n_traj == 0) return OUTFIT_OK;
  const bool have_geo = b->obs_geo_ecl != nullptr;
  const bool have_bf = b->observer_body_fixed && b->mjd_ut1;
  if (!have_geo && !have_bf) return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "need obs_geo_ecl or observer_body_fixed+mjd_ut1");
  // scratch: scorer[3][n] + tentative residuals[3][n] (+ geo[3][n] + helio[3][n] from pvobs) + status[n]
  const size_t planes = have_geo ? 6 : 12;
  int rc = ensure_scratch(ctx, planes * n * sizeof(double) + (n + 2) * sizeof(int) + 256);  // + the work counter
  if (rc) return rc;
  double *d_scorer = reinterpret_cast<double *>(ctx->scratch);
  double *d_tmp = d_scorer + 3 * n;
  int *d_status = reinterpret_cast<int *>(d_scorer + planes * n);
  const double *d_geo = b->obs_geo_ecl;
  const int tpb = 128;
  const unsigned gblocks = (unsigned)((n + tpb - 1) / tpb);
  if (!have_geo) {
    double *geo = d_scorer + 6 * n, *helio = d_scorer + 9 * n;
    if (n) observer_cache_kernel<<<gblocks, tpb, 0, stream>>>(ctx->eph, n, b->mjd_tt, b->mjd_ut1, b->observer_body_fixed, geo, helio, d_status);
    d_geo = geo;
  }
  if (n) scorer_observer_kernel<<<gblocks, tpb, 0, stream>>>(ctx->eph, n, b->mjd_tt, d_geo, d_scorer, d_status);
  LsqBatchDev B;
  B.n_traj = b->n_traj; B.n_obs = n; B.traj_offset = (const unsigned long long *)b->traj_offset;
  B.mjd_tt = b->mjd_tt; B.ra = b->ra; B.dec = b->dec; B.sigma_ra = b->sigma_ra; B.sigma_dec = b->sigma_dec;
  B.scorer = d_scorer; B.obs_status = d_status;
  unsigned long long *d_next = reinterpret_cast<unsigned long long *>(d_status + ((n + 1) & ~(size_t)1));
  CK(cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), stream));
  const unsigned long long want = (b->n_traj + kLsqQuads - 1) / kLsqQuads;
  const unsigned long long cap = (unsigned long long)ctx->sm_count * OUTFIT_LSQQ_BPS;  // resident blocks
  lsq_quad_kernel<<<(unsigned)(want < cap ? want : cap), kLsqQThreads, 0, stream>>>(B, to_lsq_dev(*cfg), iod, out, fit, d_tmp,
                                                                                    d_next);
  CK(cudaGetLastError());
  return OUTFIT_OK;
}

// FitLSQ::fit_lsq on the trajectories [tb, te) of the host batch; iod / out = &records[tb], fit = &fit[0]
// (indexed by the batch's global observation index).  The results (776 B per trajectory + 32 B per
// observation) are ~8x the input: the two arrays come back concurrently on two streams.
static int fit_lsq_range(OutfitCtx *ctx, const OutfitIodParams *iod_params, const OutfitLsqConfig *cfg,
                         const OutfitObsBatch *hb, size_t tb, size_t te, const OutfitIodResult *iod,
                         OutfitLsqResult *out, OutfitObsFit *fit) {
  const size_t T = te - tb;
  if (T == 0) return OUTFIT_OK;
  int rc = check_offsets(ctx, hb, tb, te, nullptr);
  if (rc) return rc;
  std::vector<OutfitIodResult> own_iod;
  if (!iod) {  // initial_orbits = None: run the IOD first (mod.rs:80)
    own_iod.resize(T);
    rc = fit_full_iod_range(ctx, iod_params, hb, tb, te, own_iod.data());
    if (rc) return rc;
    iod = own_iod.data();
  }
  const size_t o0 = hb->traj_offset[tb], n = hb->traj_offset[te] - o0, n_all = hb->n_obs;
  const bool have_geo = hb->obs_geo_ecl != nullptr;
  const bool have_bf = hb->observer_body_fixed && hb->mjd_ut1;
  if (!have_geo && !have_bf) return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "need obs_geo_ecl or observer_body_fixed+mjd_ut1");
  rc = ensure_host_streams(ctx);
  if (rc) return rc;
  cudaStream_t stream = ctx->compute_stream, cs = ctx->copy_stream;
  const size_t bytes = (T + 1) * 8 + 5 * n * 8 + (have_geo ? 3 : 4) * n * 8 + T * sizeof(OutfitIodResult) +
                       T * sizeof(OutfitLsqResult) + n * sizeof(OutfitObsFit) + 16 * 256;
  // the context's cached input arena (shared with the IOD host entry, whose use of it has ended by now)
  rc = ensure_arena(ctx, bytes);
  if (rc) return rc;
  ArenaPut A{ctx->arena, 0, stream};
  OutfitObsBatch db = *hb;
  db.n_traj = T; db.n_obs = n;
  if (o0 == 0) {
    db.traj_offset = (const uint64_t *)A.put(hb->traj_offset + tb, (T + 1) * 8);
  } else {
    rc = ensure_host_staging(ctx, (T + 1) * 8);
    if (rc) return rc;
    uint64_t *h = reinterpret_cast<uint64_t *>(ctx->h_scratch);
    for (size_t t = 0; t <= T; ++t) h[t] = hb->traj_offset[tb + t] - o0;
    db.traj_offset = (const uint64_t *)A.put(h, (T + 1) * 8);
  }
  db.mjd_tt = (const double *)A.put(hb->mjd_tt + o0, n * 8);
  db.ra = (const double *)A.put(hb->ra + o0, n * 8);
  db.dec = (const double *)A.put(hb->dec + o0, n * 8);
  db.sigma_ra = (const double *)A.put(hb->sigma_ra + o0, n * 8);
  db.sigma_dec = (const double *)A.put(hb->sigma_dec + o0, n * 8);
  db.obs_helio_equ = nullptr; db.noise_z = nullptr; db.traj_seed = nullptr;
  if (have_geo) {
    db.obs_geo_ecl = (const double *)A.put_planes(hb->obs_geo_ecl, n_all, o0, n);
    db.observer_body_fixed = nullptr; db.mjd_ut1 = nullptr;
  } else {
    db.obs_geo_ecl = nullptr;
    db.observer_body_fixed = (const double *)A.put_planes(hb->observer_body_fixed, n_all, o0, n);
    db.mjd_ut1 = (const double *)A.put(hb->mjd_ut1 + o0, n * 8);
  }
  const OutfitIodResult *d_iod = (const OutfitIodResult *)A.put(iod, T * sizeof(OutfitIodResult));
  OutfitLsqResult *d_out = (OutfitLsqResult *)A.raw(T * sizeof(OutfitLsqResult));
  OutfitObsFit *d_fit = (OutfitObsFit *)A.raw(n * sizeof(OutfitObsFit));
  if (A.err != cudaSuccess) { cudaStreamSynchronize(stream); return fail(ctx, OUTFIT_E_CUDA, "fit_lsq: H2D", A.err); }
  rc = outfit_b200_fit_lsq_device(ctx, cfg, &db, d_iod, d_out, d_fit, stream);
  if (rc == OUTFIT_OK) {
    // two copy engines: the per-trajectory records on the compute stream, the per-observation fit array on the
    // copy stream behind an event
    while (ctx->copy_ev.size() < 1) {
      cudaEvent_t e;
      CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ctx->copy_ev.push_back(e);
    }
    cudaError_t e = cudaEventRecord(ctx->copy_ev[0], stream);
    if (e == cudaSuccess && fit) {
      e = cudaStreamWaitEvent(cs, ctx->copy_ev[0], 0);
      if (e == cudaSuccess) e = cudaMemcpyAsync(fit + o0, d_fit, n * sizeof(OutfitObsFit), cudaMemcpyDeviceToHost, cs);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, T * sizeof(OutfitLsqResult), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
    if (e != cudaSuccess) rc = fail(ctx, OUTFIT_E_CUDA, "fit_lsq: copy back / kernel", e);
  } else {
    cudaStreamSynchronize(stream);
  }
  return rc;
}

extern "C" int outfit_b200_fit_iod(OutfitCtx *ctx, const OutfitIodParams *params, const OutfitObsBatch *hb,
                                   uint64_t traj_index, OutfitIodResult *out) {
  if (!ctx) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  int rc = check_iod_host_args(ctx, params, hb, out);
  if (rc) return rc;
  if (traj_index >= hb->n_traj) return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "fit_iod: no such trajectory");
  CK(cudaSetDevice(ctx->device));
  return fit_full_iod_range(ctx, params, hb, traj_index, traj_index + 1, out);
}

static int check_lsq_host_args(OutfitCtx *ctx, const OutfitIodParams *iod_params, const OutfitLsqConfig *cfg,
                               const OutfitObsBatch *hb, const OutfitIodResult *iod, const void *out) {
  if (!ctx || !cfg || !hb || (!out && hb->n_traj)) return OUTFIT_E_INVALID_ARGUMENT;
  if (!iod && !iod_params) return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "fit_lsq: iod results or iod_params are required");
  if (!iod) {
    const int rc = outfit_b200_iod_params_validate(iod_params);
    if (rc) return fail(ctx, rc, "IODParams validation failed (mod.rs:544-624)");
  }
  if (hb->n_traj && (!hb->traj_offset || !hb->mjd_tt || !hb->ra || !hb->dec || !hb->sigma_ra || !hb->sigma_dec))
    return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "NULL observation array");
  return OUTFIT_OK;
}

extern "C" int outfit_b200_fit_lsq(OutfitCtx *ctx, const OutfitIodParams *iod_params, const OutfitLsqConfig *cfg,
                                   const OutfitObsBatch *hb, const OutfitIodResult *iod, OutfitLsqResult *out,
                                   OutfitObsFit *fit) {
  if (!ctx) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  int rc = check_lsq_host_args(ctx, iod_params, cfg, hb, iod, out);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  return fit_lsq_range(ctx, iod_params, cfg, hb, 0, hb->n_traj, iod, out, fit);
}

extern "C" int outfit_b200_last_iod_counters(OutfitCtx *ctx, OutfitIodCounters *out) {
  if (!ctx || !out) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  unsigned long long h[32];
  // the launch may sit on non-blocking streams, which a legacy-stream copy does not wait for
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost));
  // order of struct Work (dev_kepler.cuh)
  out->gauss_solves = h[1]; out->aberth_sweeps = h[2]; out->roots_accepted = h[3]; out->fg_iterations = h[4];
  out->kepler_universal_solves = h[5]; out->newton_steps = h[6]; out->sfunct_terms = h[7];
  out->scorer_evals = h[8]; out->scorer_newton_steps = h[9]; out->candidates = h[10];
  out->fg_iterations_skipped = h[11];
  return OUTFIT_OK;
}

// debug builds only (make EXTRA=-DOUTFIT_DEBUG_STRAGGLERS): the raw counter slots
extern "C" int outfit_b200_debug_counters(OutfitCtx *ctx, unsigned long long *out32) {
  if (!ctx || !out32) return OUTFIT_E_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpy(out32, ctx->d_counters, 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return OUTFIT_OK;
}

#ifdef OUTFIT_DEBUG_FGHIST
extern "C" int outfit_b200_debug_fghist(OutfitCtx *ctx, unsigned long long *out384, int reset) {
  if (!ctx || !out384) return OUTFIT_E_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpyFromSymbol(out384, g_fghist, 3 * 128 * sizeof(unsigned long long)));
  if (reset) {
    static const unsigned long long zero[3 * 128] = {0};
    CK(cudaMemcpyToSymbol(g_fghist, zero, sizeof zero));
  }
  return OUTFIT_OK;
}
#endif

extern "C" int outfit_b200_set_work_counters(OutfitCtx *ctx, int enabled) {
  if (!ctx) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  ctx->count_work = enabled != 0;
  return OUTFIT_OK;
}

extern "C" int outfit_b200_last_iod_phase_ms(OutfitCtx *ctx, OutfitIodPhaseMs *out) {
  if (!ctx || !out) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  memset(out, 0, sizeof *out);
  out->n_chunks = ctx->phase_chunks;
  out->kernel_launches = ctx->phase_chunks * 5u + ctx->phase_observer_kernels;
  if (!ctx->phase_valid) {  // several passes in flight: only the launch counts are meaningful
    out->observer_ms = out->triplets_ms = out->roots_ms = out->correct_ms = out->score_ms = out->select_ms = out->total_ms = -1.f;
    return ctx->phase_chunks ? OUTFIT_OK : fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "no full-IOD launch recorded on this context");
  }
  CK(cudaSetDevice(ctx->device));
  const size_t n_ev = 2 + 5 * (size_t)ctx->phase_chunks;
  CK(cudaEventSynchronize(ctx->phase_ev[n_ev - 1]));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ctx->phase_ev[0], ctx->phase_ev[1]));
  out->observer_ms = ms;
  float *acc[5] = {&out->triplets_ms, &out->roots_ms, &out->correct_ms, &out->score_ms, &out->select_ms};
  for (unsigned c = 0; c < ctx->phase_chunks; ++c)
    for (int q = 0; q < 5; ++q) {
      CK(cudaEventElapsedTime(&ms, ctx->phase_ev[1 + 5 * c + q], ctx->phase_ev[2 + 5 * c + q]));
      *acc[q] += ms;
    }
  CK(cudaEventElapsedTime(&ms, ctx->phase_ev[0], ctx->phase_ev[n_ev - 1]));
  out->total_ms = ms;
  out->n_chunks = ctx->phase_chunks;
  out->kernel_launches = ctx->phase_chunks * 5u + ctx->phase_observer_kernels;
  return OUTFIT_OK;
}

extern "C" int outfit_b200_observer_cache_device(OutfitCtx *ctx, size_t n, const double *mjd_tt,
                                                 const double *mjd_ut1, const double *bf, double *geo_ecl,
                                                 double *helio_equ, int32_t *status, void *cuda_stream) {
  if (!ctx || (n && (!mjd_tt || !mjd_ut1 || !bf || !geo_ecl || !helio_equ))) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (!ctx->have_eph) return fail(ctx, OUTFIT_E_NO_EPHEMERIS, "outfit_b200_load_ephemeris must be called first");
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return OUTFIT_OK;
  const int tpb = 128;
  observer_cache_kernel<<<(unsigned)((n + tpb - 1) / tpb), tpb, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(
      ctx->eph, n, mjd_tt, mjd_ut1, bf, geo_ecl, helio_equ, status);
  CK(cudaGetLastError());
  return OUTFIT_OK;
}

extern "C" int outfit_b200_propagate_universal_device(OutfitCtx *ctx, size_t n, const double *rv, const double *t0,
                                                      const double *t1, const double *psi_guess,
                                                      const OutfitSolverType *solver, double *out, int32_t *status,
                                                      void *cuda_stream) {
  if (!ctx || !solver || (n && (!rv || !t0 || !t1 || !out || !status))) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return OUTFIT_OK;
  propagate_universal_kernel<<<(unsigned)((n + kPropTile - 1) / kPropTile), kPropThreads, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(
      n, rv, t0, t1, psi_guess, *solver, out, status);
  CK(cudaGetLastError());
  return OUTFIT_OK;
}

// ---- three-stage ring for the bulk host entries: H2D (copy stream) -> kernels (compute stream) -> D2H
// (d2h stream), kRing chunks in flight, every chunk in its own slot of the cached arena.  PCIe is full duplex,
// so the upload of chunk j+1 and the download of chunk j-1 run under the kernels of chunk j.
constexpr int kRing = 3;
static int ensure_ring(OutfitCtx *ctx) {
  int rc = ensure_host_streams(ctx);
  if (rc) return rc;
  if (!ctx->d2h_stream) CK(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 3 * kRing; ++i)
    if (!ctx->ring_ev[i]) CK(cudaEventCreateWithFlags(&ctx->ring_ev[i], cudaEventDisableTiming));
  return OUTFIT_OK;
}

// kepler::propagate_universal over the columns [i0, i1) of the caller's plane-major arrays (n_all columns)
static int propagate_range(OutfitCtx *ctx, size_t n_all, size_t i0, size_t i1, const double *rv, const double *t0,
                           const double *t1, const double *psi_guess, const OutfitSolverType *solver, double *out,
                           int32_t *status) {
  if (i1 <= i0) return OUTFIT_OK;
  int rc = ensure_ring(ctx);
  if (rc) return rc;
  const size_t n = i1 - i0;
  size_t chunk = (size_t)1 << 20;
  if (const char *ev = getenv("OUTFIT_B200_PROP_CHUNK")) { const long v = atol(ev); if (v >= 1024) chunk = (size_t)v; }
  if (chunk > n) chunk = n;
  const size_t in_pl = 8 + (psi_guess ? 1 : 0);
  const size_t slot_bytes = ((in_pl + 11) * chunk * 8 + chunk * 4 + 4 * 256 + 255) & ~(size_t)255;
  rc = ensure_arena(ctx, slot_bytes * kRing);
  if (rc) return rc;
  cudaStream_t up = ctx->copy_stream, run = ctx->compute_stream, down = ctx->d2h_stream;
  cudaError_t e = cudaSuccess;
  size_t j = 0;
  for (size_t c0 = i0; c0 < i1 && e == cudaSuccess; c0 += chunk, ++j) {
    const size_t c = i1 - c0 < chunk ? i1 - c0 : chunk;
    const int sl = (int)(j % kRing);
    cudaEvent_t ev_up = ctx->ring_ev[3 * sl], ev_run = ctx->ring_ev[3 * sl + 1], ev_down = ctx->ring_ev[3 * sl + 2];
    unsigned char *base = ctx->arena + (size_t)sl * slot_bytes;
    double *d_rv = reinterpret_cast<double *>(base), *d_t0 = d_rv + 6 * c, *d_t1 = d_rv + 7 * c;
    double *d_pg = psi_guess ? d_rv + 8 * c : nullptr, *d_out = d_rv + in_pl * c;
    int *d_st = reinterpret_cast<int *>(d_out + 11 * c);
    if (j >= (size_t)kRing) e = cudaStreamWaitEvent(up, ev_down, 0);  // the slot's previous results have left
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(d_rv, c * 8, rv + c0, n_all * 8, c * 8, 6, cudaMemcpyHostToDevice, up);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_t0, t0 + c0, c * 8, cudaMemcpyHostToDevice, up);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_t1, t1 + c0, c * 8, cudaMemcpyHostToDevice, up);
    if (e == cudaSuccess && psi_guess) e = cudaMemcpyAsync(d_pg, psi_guess + c0, c * 8, cudaMemcpyHostToDevice, up);
    if (e == cudaSuccess) e = cudaEventRecord(ev_up, up);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(run, ev_up, 0);
    if (e != cudaSuccess) break;
    rc = outfit_b200_propagate_universal_device(ctx, c, d_rv, d_t0, d_t1, d_pg, solver, d_out, d_st, run);
    if (rc) break;
    e = cudaEventRecord(ev_run, run);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(down, ev_run, 0);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(out + c0, n_all * 8, d_out, c * 8, c * 8, 11, cudaMemcpyDeviceToHost, down);
    if (e == cudaSuccess) e = cudaMemcpyAsync(status + c0, d_st, c * sizeof(int), cudaMemcpyDeviceToHost, down);
    if (e == cudaSuccess) e = cudaEventRecord(ev_down, down);
  }
  cudaError_t e2 = cudaStreamSynchronize(up);
  if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(run);
  if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(down);
  if (rc) return rc;
  if (e != cudaSuccess || e2 != cudaSuccess) return fail(ctx, OUTFIT_E_CUDA, "propagate_universal: copy / kernel", e != cudaSuccess ? e : e2);
  return OUTFIT_OK;
}

extern "C" int outfit_b200_propagate_universal(OutfitCtx *ctx, size_t n, const double *rv, const double *t0,
                                               const double *t1, const double *psi_guess,
                                               const OutfitSolverType *solver, double *out, int32_t *status) {
  if (!ctx || !solver || (n && (!rv || !t0 || !t1 || !out || !status))) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  return propagate_range(ctx, n, 0, n, rv, t0, t1, psi_guess, solver, out, status);
}

// ---- two-body Combined ephemeris (ephemeris/mod.rs:189-292) -------------------------------------------
// observer table [9][e_stride] + status[e_stride] in the context scratch (rows 16-byte aligned), one thread per
// epoch.  d_bf = device [3][n_epochs] per-epoch body-fixed positions (a request with several observers) or null
// (one observer: bf[3]).
static int ephemeris_observer_table(OutfitCtx *ctx, size_t n_epochs, const double *mjd_tt, const double *mjd_ut1,
                                    const double *d_bf, const double bf[3], cudaStream_t stream, size_t *e_stride_out,
                                    double **d_table_out, int **d_ost_out) {
  const size_t e_stride = (n_epochs + 1) & ~(size_t)1;
  int rc = ensure_scratch(ctx, 9 * e_stride * sizeof(double) + e_stride * sizeof(int) + 256);
  if (rc) return rc;
  double *d_table = reinterpret_cast<double *>(ctx->scratch);
  int *d_ost = reinterpret_cast<int *>(d_table + 9 * e_stride);
  ephemeris_observer_kernel<<<(unsigned)((e_stride + 127) / 128), 128, 0, stream>>>(
      ctx->eph, n_epochs, e_stride, mjd_tt, mjd_ut1, d_bf, bf ? bf[0] : 0.0, bf ? bf[1] : 0.0, bf ? bf[2] : 0.0, d_table, d_ost);
  *e_stride_out = e_stride; *d_table_out = d_table; *d_ost_out = d_ost;
  return OUTFIT_OK;
}

static int ephemeris_device_impl(OutfitCtx *ctx, size_t n_orbits, const int32_t *kind, const double *epoch,
                                 const double *elem, size_t n_epochs, const double *mjd_tt, const double *mjd_ut1,
                                 const double *d_bf, const double bf[3], double *out, int32_t *status, cudaStream_t stream) {
  if (n_orbits && n_epochs && (!kind || !epoch || !elem || !mjd_tt || !mjd_ut1 || !out || !status)) return OUTFIT_E_INVALID_ARGUMENT;
  if (!ctx->have_eph) return fail(ctx, OUTFIT_E_NO_EPHEMERIS, "outfit_b200_load_ephemeris must be called first");
  CK(cudaSetDevice(ctx->device));
  if (n_orbits == 0 || n_epochs == 0) return OUTFIT_OK;
  size_t e_stride;
  double *d_table;
  int *d_ost;
  int rc = ephemeris_observer_table(ctx, n_epochs, mjd_tt, mjd_ut1, d_bf, bf, stream, &e_stride, &d_table, &d_ost);
  if (rc) return rc;
  if (ctx->aberration_order == 2)
    ephemeris_twobody_kernel<true><<<(unsigned)((n_orbits + kEphThreads - 1) / kEphThreads), kEphThreads, 0, stream>>>(
        n_orbits, kind, epoch, elem, n_epochs, e_stride, mjd_tt, d_table, d_ost, out, status);
  else
    ephemeris_twobody_kernel<false><<<(unsigned)((n_orbits + kEphThreads - 1) / kEphThreads), kEphThreads, 0, stream>>>(
        n_orbits, kind, epoch, elem, n_epochs, e_stride, mjd_tt, d_table, d_ost, out, status);
  CK(cudaGetLastError());
  return OUTFIT_OK;
}

extern "C" void outfit_b200_ephemeris_config_default(OutfitEphemerisConfig *c) {  // EphemerisConfig::default()
  if (!c) return;
  c->propagator = OUTFIT_PROPAGATOR_TWOBODY;
  c->aberration = OUTFIT_ABERRATION_FIRST;
}
extern "C" int outfit_b200_set_ephemeris_config(OutfitCtx *ctx, const OutfitEphemerisConfig *c) {
  if (!ctx || !c) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (c->propagator != OUTFIT_PROPAGATOR_TWOBODY)
    return fail(ctx, OUTFIT_E_UNSUPPORTED, "PropagatorKind::NBody needs the perturber snapshots: call outfit_b200_ephemeris_nbody, which takes them");
  if (c->aberration != OUTFIT_ABERRATION_FIRST && c->aberration != OUTFIT_ABERRATION_SECOND)
    return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "aberration must be OUTFIT_ABERRATION_FIRST or _SECOND");
  ctx->aberration_order = c->aberration == OUTFIT_ABERRATION_SECOND ? 2 : 1;
  return OUTFIT_OK;
}
extern "C" int outfit_b200_group_set_ephemeris_config(OutfitGroup *g, const OutfitEphemerisConfig *c);

extern "C" int outfit_b200_ephemeris_twobody_device(OutfitCtx *ctx, size_t n_orbits, const int32_t *kind,
                                                    const double *epoch, const double *elem, size_t n_epochs,
                                                    const double *mjd_tt, const double *mjd_ut1,
                                                    const double body_fixed[3], double *out, int32_t *status,
                                                    void *cuda_stream) {
  if (!ctx || !body_fixed) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  return ephemeris_device_impl(ctx, n_orbits, kind, epoch, elem, n_epochs, mjd_tt, mjd_ut1, nullptr, body_fixed, out, status,
                               reinterpret_cast<cudaStream_t>(cuda_stream));
}

extern "C" int outfit_b200_ephemeris_request_device(OutfitCtx *ctx, size_t n_orbits, const int32_t *kind,
                                                    const double *epoch, const double *elem, size_t n_epochs,
                                                    const double *mjd_tt, const double *mjd_ut1,
                                                    const double *epoch_body_fixed, double *out, int32_t *status,
                                                    void *cuda_stream) {
  if (!ctx || (n_epochs && !epoch_body_fixed)) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  return ephemeris_device_impl(ctx, n_orbits, kind, epoch, elem, n_epochs, mjd_tt, mjd_ut1, epoch_body_fixed, nullptr, out, status,
                               reinterpret_cast<cudaStream_t>(cuda_stream));
}

// Host entry over the orbits [i0, i1) of the caller's arrays (n_all orbits = the plane stride of elem / out /
// status): epochs and the observer table once, then the orbits in chunks through the three-stage ring.  The
// output is 76 B per (orbit, epoch) entry against ~600 flop: end to end this call is bound by the D2H copy.
// h_bf: host [3][n_epochs] per-epoch body-fixed observer positions (several observers) or null (-> bf[3]).
static int ephemeris_range(OutfitCtx *ctx, size_t n_all, size_t i0, size_t i1, const int32_t *kind, const double *epoch,
                           const double *elem, size_t n_epochs, const double *mjd_tt, const double *mjd_ut1,
                           const double *h_bf, const double bf[3], double *out, int32_t *status) {
  if (i1 <= i0 || n_epochs == 0) return OUTFIT_OK;
  if (!ctx->have_eph) return fail(ctx, OUTFIT_E_NO_EPHEMERIS, "outfit_b200_load_ephemeris must be called first");
  int rc = ensure_ring(ctx);
  if (rc) return rc;
  const size_t n = i1 - i0;
  // chunk: ~256 MB of output per slot
  size_t chunk = ((size_t)256 << 20) / (76 * n_epochs);
  chunk = chunk < 1024 ? 1024 : (chunk & ~(size_t)127);
  if (const char *ev = getenv("OUTFIT_B200_EPH_CHUNK")) { const long v = atol(ev); if (v >= 128) chunk = (size_t)v; }
  if (chunk > n) chunk = n;
  const size_t head_bytes = ((5 * n_epochs * 8 + 255) & ~(size_t)255) + 256;
  const size_t slot_bytes = ((chunk * 4 + 255) & ~(size_t)255) + ((7 * chunk * 8 + 255) & ~(size_t)255) +
                            ((9 * n_epochs * chunk * 8 + 255) & ~(size_t)255) + ((n_epochs * chunk * 4 + 255) & ~(size_t)255);
  rc = ensure_arena(ctx, head_bytes + slot_bytes * kRing);
  if (rc) return rc;
  cudaStream_t up = ctx->copy_stream, run = ctx->compute_stream, down = ctx->d2h_stream;
  double *d_tt = reinterpret_cast<double *>(ctx->arena), *d_ut = d_tt + n_epochs, *d_bf = h_bf ? d_ut + n_epochs : nullptr;
  cudaError_t e = cudaMemcpyAsync(d_tt, mjd_tt, n_epochs * 8, cudaMemcpyHostToDevice, run);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_ut, mjd_ut1, n_epochs * 8, cudaMemcpyHostToDevice, run);
  if (e == cudaSuccess && h_bf) e = cudaMemcpyAsync(d_bf, h_bf, 3 * n_epochs * 8, cudaMemcpyHostToDevice, run);
  if (e != cudaSuccess) return fail(ctx, OUTFIT_E_CUDA, "ephemeris_twobody: H2D", e);
  size_t e_stride;
  double *d_table;
  int *d_ost;
  rc = ephemeris_observer_table(ctx, n_epochs, d_tt, d_ut, d_bf, bf, run, &e_stride, &d_table, &d_ost);
  if (rc) return rc;
  size_t j = 0;
  for (size_t c0 = i0; c0 < i1 && e == cudaSuccess; c0 += chunk, ++j) {
    const size_t c = i1 - c0 < chunk ? i1 - c0 : chunk;
    const int sl = (int)(j % kRing);
    cudaEvent_t ev_up = ctx->ring_ev[3 * sl], ev_run = ctx->ring_ev[3 * sl + 1], ev_down = ctx->ring_ev[3 * sl + 2];
    unsigned char *base = ctx->arena + head_bytes + (size_t)sl * slot_bytes;
    int32_t *d_kind = reinterpret_cast<int32_t *>(base);
    double *d_epoch = reinterpret_cast<double *>(base + ((chunk * 4 + 255) & ~(size_t)255));
    double *d_elem = d_epoch + c;
    double *d_out = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(d_epoch) + ((7 * chunk * 8 + 255) & ~(size_t)255));
    int32_t *d_st = reinterpret_cast<int32_t *>(reinterpret_cast<unsigned char *>(d_out) + ((9 * n_epochs * chunk * 8 + 255) & ~(size_t)255));
    if (j >= (size_t)kRing) e = cudaStreamWaitEvent(up, ev_down, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_kind, kind + c0, c * 4, cudaMemcpyHostToDevice, up);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_epoch, epoch + c0, c * 8, cudaMemcpyHostToDevice, up);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(d_elem, c * 8, elem + c0, n_all * 8, c * 8, 6, cudaMemcpyHostToDevice, up);
    if (e == cudaSuccess) e = cudaEventRecord(ev_up, up);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(run, ev_up, 0);
    if (e != cudaSuccess) break;
    if (ctx->aberration_order == 2)
      ephemeris_twobody_kernel<true><<<(unsigned)((c + kEphThreads - 1) / kEphThreads), kEphThreads, 0, run>>>(
          c, d_kind, d_epoch, d_elem, n_epochs, e_stride, d_tt, d_table, d_ost, d_out, d_st);
    else
      ephemeris_twobody_kernel<false><<<(unsigned)((c + kEphThreads - 1) / kEphThreads), kEphThreads, 0, run>>>(
          c, d_kind, d_epoch, d_elem, n_epochs, e_stride, d_tt, d_table, d_ost, d_out, d_st);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaEventRecord(ev_run, run);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(down, ev_run, 0);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(out + c0, n_all * 8, d_out, c * 8, c * 8, 9 * n_epochs, cudaMemcpyDeviceToHost, down);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(status + c0, n_all * 4, d_st, c * 4, c * 4, n_epochs, cudaMemcpyDeviceToHost, down);
    if (e == cudaSuccess) e = cudaEventRecord(ev_down, down);
  }
  cudaError_t e2 = cudaStreamSynchronize(up);
  if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(run);
  if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(down);
  if (e != cudaSuccess || e2 != cudaSuccess) return fail(ctx, OUTFIT_E_CUDA, "ephemeris_twobody: copy / kernel", e != cudaSuccess ? e : e2);
  return OUTFIT_OK;
}

extern "C" int outfit_b200_ephemeris_twobody(OutfitCtx *ctx, size_t n_orbits, const int32_t *kind, const double *epoch,
                                             const double *elem, size_t n_epochs, const double *mjd_tt,
                                             const double *mjd_ut1, const double body_fixed[3], double *out,
                                             int32_t *status) {
  if (!ctx || !body_fixed) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (n_orbits && n_epochs && (!kind || !epoch || !elem || !mjd_tt || !mjd_ut1 || !out || !status)) return OUTFIT_E_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  return ephemeris_range(ctx, n_orbits, 0, n_orbits, kind, epoch, elem, n_epochs, mjd_tt, mjd_ut1, nullptr, body_fixed, out, status);
}

// EphemerisRequest with several (observer, epochs) pairs (ephemeris/request.rs:276-340, mod.rs:242-290): the
// epochs of all observers concatenated in request order, observer o owning [epoch_offset[o], epoch_offset[o+1]).
static int expand_request(OutfitCtx *ctx, size_t n_observers, const double *observer_body_fixed, const uint64_t *epoch_offset,
                          std::vector<double> &bf_planes, size_t *n_epochs) {
  if (!observer_body_fixed || !epoch_offset || epoch_offset[0] != 0) return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "ephemeris_request: observers / epoch_offset");
  for (size_t o = 0; o < n_observers; ++o)
    if (epoch_offset[o + 1] < epoch_offset[o]) return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "ephemeris_request: epoch_offset is not monotone");
  const size_t E = epoch_offset[n_observers];
  bf_planes.resize(3 * E);
  for (size_t o = 0; o < n_observers; ++o)
    for (size_t e = epoch_offset[o]; e < epoch_offset[o + 1]; ++e)
      for (int q = 0; q < 3; ++q) bf_planes[(size_t)q * E + e] = observer_body_fixed[3 * o + q];
  *n_epochs = E;
  return OUTFIT_OK;
}

extern "C" int outfit_b200_ephemeris_request(OutfitCtx *ctx, size_t n_orbits, const int32_t *kind, const double *epoch,
                                             const double *elem, size_t n_observers, const double *observer_body_fixed,
                                             const uint64_t *epoch_offset, const double *mjd_tt, const double *mjd_ut1,
                                             double *out, int32_t *status) {
  if (!ctx) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (n_observers == 0 || n_orbits == 0) return OUTFIT_OK;
  std::vector<double> bfp;
  size_t E = 0;
  int rc = expand_request(ctx, n_observers, observer_body_fixed, epoch_offset, bfp, &E);
  if (rc) return rc;
  if (E && (!kind || !epoch || !elem || !mjd_tt || !mjd_ut1 || !out || !status)) return OUTFIT_E_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  return ephemeris_range(ctx, n_orbits, 0, n_orbits, kind, epoch, elem, E, mjd_tt, mjd_ut1, bfp.data(), nullptr, out, status);
}

// =================================================================================================
// multi-GPU group: ONE call drives every GPU of the box, like fit_full_iod_parallel drives every core
// (obs_dataset_api.rs:175-207).  Trajectories are independent, so the batch is cut into contiguous
// trajectory ranges of near-equal estimated work, one per GPU; one host thread per GPU runs the
// single-GPU host entry on its range (own context, arena, streams), and every result lands at its global
// trajectory index in the caller's array.  No collective, no peer traffic: the only shared resource is the
// host memory the ranges are read from.
// =================================================================================================
struct OutfitGroup {
  std::vector<OutfitCtx *> ctx;
  std::vector<float> shard_ms;                 // wall time of every shard in the last group call
  std::vector<unsigned long long> shard_cut;   // trajectory / item cuts of the last group call [n + 1]
  std::string last_error;
  std::mutex mu;
};

// Relative cost of a trajectory of n observations: candidates x (Gauss solve + f-g + arc scoring) + enumeration
static double traj_work(double n, double K, double M) {
  const double feasible = n * (n - 1.0) * (n - 2.0) / 6.0;
  const double k = feasible < K ? (feasible > 0.0 ? feasible : 0.0) : K;
  return k * M * (60.0 + n) + 0.05 * (feasible > 0.0 ? feasible : 0.0) + 1.0;
}

extern "C" int outfit_b200_shard_ranges(uint64_t n_traj, const uint64_t *traj_offset, uint32_t max_triplets,
                                        uint64_t n_noise_realizations, int n_parts, uint64_t *cuts) {
  if (n_parts < 1 || !cuts || (n_traj && !traj_offset)) return OUTFIT_E_INVALID_ARGUMENT;
  const double K = (double)max_triplets, M = 1.0 + (double)n_noise_realizations;
  double total = 0.0;
  for (uint64_t t = 0; t < n_traj; ++t) total += traj_work((double)(traj_offset[t + 1] - traj_offset[t]), K, M);
  cuts[0] = 0;
  double acc = 0.0;
  int r = 1;
  for (uint64_t t = 0; t < n_traj && r < n_parts; ++t) {
    acc += traj_work((double)(traj_offset[t + 1] - traj_offset[t]), K, M);
    // first prefix that reaches r / n_parts of the total work ends part r - 1 BEFORE trajectory t
    // (numpy searchsorted(cumsum, total * r / n, side="left") of outfit_b200/shard.py)
    while (r < n_parts && acc >= total * (double)r / (double)n_parts) cuts[r++] = t;
  }
  while (r < n_parts) cuts[r++] = n_traj;
  cuts[n_parts] = n_traj;
  for (int i = 1; i <= n_parts; ++i)
    if (cuts[i] < cuts[i - 1]) cuts[i] = cuts[i - 1];
  return OUTFIT_OK;
}

extern "C" int outfit_b200_init_multi(int n_gpus, const int *device_ids, OutfitGroup **out) {
  if (!out) return OUTFIT_E_INVALID_ARGUMENT;
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return OUTFIT_E_NO_DEVICE;
  if (n_gpus <= 0) { n_gpus = n_dev; device_ids = nullptr; }
  if (!device_ids && n_gpus > n_dev) return OUTFIT_E_INVALID_ARGUMENT;
  OutfitGroup *g = new (std::nothrow) OutfitGroup();
  if (!g) return OUTFIT_E_ALLOC;
  for (int i = 0; i < n_gpus; ++i) {
    const int dev = device_ids ? device_ids[i] : i;
    OutfitCtx *c = nullptr;
    const int rc = (dev < 0 || dev >= n_dev) ? OUTFIT_E_INVALID_ARGUMENT : outfit_b200_init(dev, &c);
    if (rc) {
      for (OutfitCtx *q : g->ctx) outfit_b200_destroy(q);
      delete g;
      return rc;
    }
    g->ctx.push_back(c);
  }
  g->shard_ms.assign(n_gpus, 0.f);
  g->shard_cut.assign(n_gpus + 1, 0ull);
  *out = g;
  return OUTFIT_OK;
}
extern "C" void outfit_b200_group_destroy(OutfitGroup *g) {
  if (!g) return;
  for (OutfitCtx *c : g->ctx) outfit_b200_destroy(c);
  delete g;
}
extern "C" int outfit_b200_group_size(OutfitGroup *g) { return g ? (int)g->ctx.size() : 0; }
extern "C" OutfitCtx *outfit_b200_group_ctx(OutfitGroup *g, int i) {
  return (g && i >= 0 && (size_t)i < g->ctx.size()) ? g->ctx[i] : nullptr;
}
extern "C" const char *outfit_b200_group_last_error(OutfitGroup *g) { return g ? g->last_error.c_str() : ""; }
extern "C" int outfit_b200_group_last_shards(OutfitGroup *g, uint64_t *cuts, float *ms) {
  if (!g) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(g->mu);
  for (size_t i = 0; i < g->ctx.size(); ++i) {
    if (ms) ms[i] = g->shard_ms[i];
    if (cuts) cuts[i] = g->shard_cut[i];
  }
  if (cuts) cuts[g->ctx.size()] = g->shard_cut[g->ctx.size()];
  return OUTFIT_OK;
}

extern "C" int outfit_b200_group_load_ephemeris(OutfitGroup *g, const double *cheb, size_t n_blocks, size_t block_stride,
                                                double jd_start, double block_days, const uint32_t ipt[3][3], double emrat) {
  if (!g) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(g->mu);
  for (OutfitCtx *c : g->ctx) {  // replicated: 0.5 - 46 MB, read-only
    const int rc = outfit_b200_load_ephemeris(c, cheb, n_blocks, block_stride, jd_start, block_days, ipt, emrat);
    if (rc) { g->last_error = c->last_error; return rc; }
  }
  return OUTFIT_OK;
}
extern "C" int outfit_b200_group_set_ephemeris_config(OutfitGroup *g, const OutfitEphemerisConfig *c) {
  if (!g) return OUTFIT_E_INVALID_ARGUMENT;
  for (OutfitCtx *x : g->ctx) {
    const int rc = outfit_b200_set_ephemeris_config(x, c);
    if (rc) { g->last_error = x->last_error; return rc; }
  }
  return OUTFIT_OK;
}
extern "C" int outfit_b200_group_set_pass_streams(OutfitGroup *g, int n_streams) {
  if (!g) return OUTFIT_E_INVALID_ARGUMENT;
  for (OutfitCtx *c : g->ctx) {
    const int rc = outfit_b200_set_pass_streams(c, n_streams);
    if (rc) return rc;
  }
  return OUTFIT_OK;
}

// run fn(i, ctx, begin, end) for every shard on its own host thread; first failure wins
template <class F>
static int group_run(OutfitGroup *g, const std::vector<unsigned long long> &cuts, F fn) {
  const size_t n = g->ctx.size();
  std::vector<int> rcs(n, OUTFIT_OK);
  std::vector<std::thread> th;
  g->shard_cut = cuts;
  auto body = [&](size_t i) {
    OutfitCtx *c = g->ctx[i];
    std::lock_guard<std::recursive_mutex> lock(c->mu);
    timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    if (cudaSetDevice(c->device) != cudaSuccess) rcs[i] = fail(c, OUTFIT_E_CUDA, "cudaSetDevice");
    else rcs[i] = fn(i, c, cuts[i], cuts[i + 1]);
    clock_gettime(CLOCK_MONOTONIC, &b);
    g->shard_ms[i] = (float)((b.tv_sec - a.tv_sec) * 1e3 + (b.tv_nsec - a.tv_nsec) * 1e-6);
  };
  for (size_t i = 1; i < n; ++i) th.emplace_back(body, i);
  body(0);
  for (std::thread &t : th) t.join();
  for (size_t i = 0; i < n; ++i)
    if (rcs[i]) { g->last_error = "shard " + std::to_string(i) + ": " + g->ctx[i]->last_error; return rcs[i]; }
  return OUTFIT_OK;
}

static std::vector<unsigned long long> even_cuts(size_t n_items, size_t parts, size_t align) {
  std::vector<unsigned long long> cuts(parts + 1, n_items);
  cuts[0] = 0;
  for (size_t i = 1; i < parts; ++i) {
    size_t c = n_items * i / parts;
    c = (c + align - 1) / align * align;
    cuts[i] = c < n_items ? c : n_items;
  }
  return cuts;
}

extern "C" int outfit_b200_group_fit_full_iod(OutfitGroup *g, const OutfitIodParams *params, const OutfitObsBatch *hb,
                                              OutfitIodResult *out) {
  if (!g || g->ctx.empty()) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(g->mu);
  int rc = check_iod_host_args(g->ctx[0], params, hb, out);
  if (rc == OUTFIT_OK && hb->n_traj) rc = check_offsets(g->ctx[0], hb, 0, hb->n_traj, nullptr);
  if (rc) { g->last_error = g->ctx[0]->last_error; return rc; }
  std::vector<unsigned long long> cuts(g->ctx.size() + 1, 0ull);
  static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "cut type");
  outfit_b200_shard_ranges(hb->n_traj, hb->traj_offset, params->max_triplets, params->n_noise_realizations, (int)g->ctx.size(),
                           reinterpret_cast<uint64_t *>(cuts.data()));
  return group_run(g, cuts, [&](size_t, OutfitCtx *c, unsigned long long tb, unsigned long long te) {
    return fit_full_iod_range(c, params, hb, tb, te, out + tb);
  });
}

extern "C" int outfit_b200_group_fit_lsq(OutfitGroup *g, const OutfitIodParams *iod_params, const OutfitLsqConfig *cfg,
                                         const OutfitObsBatch *hb, const OutfitIodResult *iod, OutfitLsqResult *out,
                                         OutfitObsFit *fit) {
  if (!g || g->ctx.empty()) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(g->mu);
  int rc = check_lsq_host_args(g->ctx[0], iod_params, cfg, hb, iod, out);
  if (rc == OUTFIT_OK && hb->n_traj) rc = check_offsets(g->ctx[0], hb, 0, hb->n_traj, nullptr);
  if (rc) { g->last_error = g->ctx[0]->last_error; return rc; }
  std::vector<unsigned long long> cuts(g->ctx.size() + 1, 0ull);
  // the IOD dominates when it runs first; the correction alone costs ~ the number of observations
  outfit_b200_shard_ranges(hb->n_traj, hb->traj_offset, iod ? 1u : iod_params->max_triplets, iod ? 0ull : iod_params->n_noise_realizations,
                           (int)g->ctx.size(), reinterpret_cast<uint64_t *>(cuts.data()));
  return group_run(g, cuts, [&](size_t, OutfitCtx *c, unsigned long long tb, unsigned long long te) {
    return fit_lsq_range(c, iod_params, cfg, hb, tb, te, iod ? iod + tb : nullptr, out + tb, fit);
  });
}

extern "C" int outfit_b200_group_propagate_universal(OutfitGroup *g, size_t n, const double *rv, const double *t0,
                                                     const double *t1, const double *psi_guess, const OutfitSolverType *solver,
                                                     double *out, int32_t *status) {
  if (!g || g->ctx.empty() || !solver || (n && (!rv || !t0 || !t1 || !out || !status))) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(g->mu);
  return group_run(g, even_cuts(n, g->ctx.size(), 128), [&](size_t, OutfitCtx *c, unsigned long long i0, unsigned long long i1) {
    return propagate_range(c, n, i0, i1, rv, t0, t1, psi_guess, solver, out, status);
  });
}

extern "C" int outfit_b200_group_ephemeris_request(OutfitGroup *g, size_t n_orbits, const int32_t *kind, const double *epoch,
                                                   const double *elem, size_t n_observers, const double *observer_body_fixed,
                                                   const uint64_t *epoch_offset, const double *mjd_tt, const double *mjd_ut1,
                                                   double *out, int32_t *status) {
  if (!g || g->ctx.empty()) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(g->mu);
  if (n_observers == 0 || n_orbits == 0) return OUTFIT_OK;
  std::vector<double> bfp;
  size_t E = 0;
  int rc = expand_request(g->ctx[0], n_observers, observer_body_fixed, epoch_offset, bfp, &E);
  if (rc) { g->last_error = g->ctx[0]->last_error; return rc; }
  if (E && (!kind || !epoch || !elem || !mjd_tt || !mjd_ut1 || !out || !status)) return OUTFIT_E_INVALID_ARGUMENT;
  return group_run(g, even_cuts(n_orbits, g->ctx.size(), 128), [&](size_t, OutfitCtx *c, unsigned long long i0, unsigned long long i1) {
    return ephemeris_range(c, n_orbits, i0, i1, kind, epoch, elem, E, mjd_tt, mjd_ut1, bfp.data(), nullptr, out, status);
  });
}

// page-locked host buffers for callers without their own pinned allocator (the host entries copy
// asynchronously only from / to page-locked memory); portable across the GPUs of a group
extern "C" void *outfit_b200_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) return nullptr;
  return p;
}
extern "C" void outfit_b200_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

// ---- N-body propagation (propagator/nbody.rs, equinoctial_element.rs:908-968) ------------------------------------
extern "C" void outfit_b200_nbody_config_default(OutfitNBodyConfig *c) {  // NBodyConfig::default(): [Sun], 1e-12, 1e-12
  if (!c) return;
  c->abs_tol = 1e-12; c->rel_tol = 1e-12; c->n_perturbers = 1; c->max_steps = 0;
}
extern "C" double outfit_b200_planet_gm(int body) {  // planet_gm.rs:10-60 (DE440 values, km^3/s^2 -> AU^3/day^2)
  static const double km3_s2[11] = {1.32712440041e11, 2.203178e4, 3.2485857e5, 4.03503235e5, 4.28283736e4, 1.267127648e8,
                                    3.79406252e7, 5.7945564e6, 6.8365271e6, 9.755e2, 4.902800066e3};
  const double au_km = 1.495978707e8;
  const double conv = (86400.0 * 86400.0) / (au_km * au_km * au_km);
  return (body >= 0 && body < 11) ? km3_s2[body] * conv : NAN;
}
static int check_nbody_cfg(OutfitCtx *ctx, const OutfitNBodyConfig *cfg) {
  if (cfg->n_perturbers == 0 || cfg->n_perturbers > (unsigned)kNbMaxPert)
    return fail(ctx, OUTFIT_E_UNSUPPORTED, "n_perturbers must be in [1, 12]");
  if (!(cfg->abs_tol > 0.0) || !(cfg->rel_tol > 0.0)) return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "abs_tol and rel_tol must be positive");
  return OUTFIT_OK;
}
extern "C" int outfit_b200_propagate_nbody_device(OutfitCtx *ctx, size_t n, const int32_t *kind, const double *epoch,
                                                  const double *elem, const double *t1, const OutfitNBodyConfig *cfg,
                                                  const double *gm, const double *pert_pos, double *out, double *stm,
                                                  int32_t *status, uint32_t *steps, void *cuda_stream) {
  if (!ctx || !cfg || (n && (!kind || !epoch || !elem || !t1 || !gm || !pert_pos || !out || !status))) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  int rc = check_nbody_cfg(ctx, cfg);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return OUTFIT_OK;
  CK(cudaFuncSetAttribute(propagate_nbody_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNbSmemBytes));
  const NbCfgDev c{cfg->abs_tol, cfg->rel_tol, cfg->n_perturbers, cfg->max_steps ? cfg->max_steps : 100000u};
  const size_t threads = n * 8;
  propagate_nbody_kernel<<<(unsigned)((threads + kNbThreads - 1) / kNbThreads), kNbThreads, kNbSmemBytes, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(
      n, kind, epoch, elem, t1, c, gm, pert_pos, out, stm, status, steps);
  CK(cudaGetLastError());
  return OUTFIT_OK;
}
extern "C" int outfit_b200_propagate_nbody(OutfitCtx *ctx, size_t n, const int32_t *kind, const double *epoch, const double *elem,
                                           const double *t1, const OutfitNBodyConfig *cfg, const double *gm, const double *pert_pos,
                                           double *out, double *stm, int32_t *status, uint32_t *steps) {
  if (!ctx || !cfg || (n && (!kind || !epoch || !elem || !t1 || !gm || !pert_pos || !out || !status))) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  int rc = check_nbody_cfg(ctx, cfg);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return OUTFIT_OK;
  rc = ensure_host_streams(ctx);
  if (rc) return rc;
  const size_t P = cfg->n_perturbers;
  const size_t bytes = n * 4 + n * 8 * (1 + 6 + 1 + 3 * P + 6 + 36) + P * 8 + n * 8 + 16 * 256;
  rc = ensure_arena(ctx, bytes);
  if (rc) return rc;
  cudaStream_t stream = ctx->compute_stream;
  ArenaPut A{ctx->arena, 0, stream};
  const int32_t *d_kind = (const int32_t *)A.put(kind, n * 4);
  const double *d_epoch = (const double *)A.put(epoch, n * 8);
  const double *d_elem = (const double *)A.put(elem, 6 * n * 8);
  const double *d_t1 = (const double *)A.put(t1, n * 8);
  const double *d_gm = (const double *)A.put(gm, P * 8);
  const double *d_pp = (const double *)A.put(pert_pos, 3 * P * n * 8);
  double *d_out = (double *)A.raw(6 * n * 8);
  double *d_stm = stm ? (double *)A.raw(36 * n * 8) : nullptr;
  int32_t *d_st = (int32_t *)A.raw(n * 4);
  uint32_t *d_steps = steps ? (uint32_t *)A.raw(n * 4) : nullptr;
  if (A.err != cudaSuccess) { cudaStreamSynchronize(stream); return fail(ctx, OUTFIT_E_CUDA, "propagate_nbody: H2D", A.err); }
  rc = outfit_b200_propagate_nbody_device(ctx, n, d_kind, d_epoch, d_elem, d_t1, cfg, d_gm, d_pp, d_out, d_stm, d_st, d_steps, stream);
  if (rc == OUTFIT_OK) {
    cudaError_t e = cudaMemcpyAsync(out, d_out, 6 * n * 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && stm) e = cudaMemcpyAsync(stm, d_stm, 36 * n * 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(status, d_st, n * 4, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && steps) e = cudaMemcpyAsync(steps, d_steps, n * 4, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) rc = fail(ctx, OUTFIT_E_CUDA, "propagate_nbody: copy back / kernel", e);
  } else {
    cudaStreamSynchronize(stream);
  }
  return rc;
}

// OrbitalElements::compute::<Combined> with PropagatorKind::NBody (ephemeris/mod.rs:189-292, propagator/mod.rs:93-101):
// DEVICE buffers.  d_bf = [3][n_epochs] per-epoch body-fixed observer positions.
extern "C" int outfit_b200_ephemeris_nbody_device(OutfitCtx *ctx, size_t n_orbits, const int32_t *kind, const double *epoch,
                                                  const double *elem, size_t n_epochs, const double *mjd_tt,
                                                  const double *mjd_ut1, const double *epoch_body_fixed,
                                                  const OutfitNBodyConfig *cfg, const double *gm, const double *pert_pos,
                                                  double *out, int32_t *status, void *cuda_stream) {
  if (!ctx || !cfg) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (n_orbits && n_epochs && (!kind || !epoch || !elem || !mjd_tt || !mjd_ut1 || !epoch_body_fixed || !gm || !pert_pos || !out || !status))
    return OUTFIT_E_INVALID_ARGUMENT;
  if (!ctx->have_eph) return fail(ctx, OUTFIT_E_NO_EPHEMERIS, "outfit_b200_load_ephemeris must be called first");
  int rc = check_nbody_cfg(ctx, cfg);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  if (n_orbits == 0 || n_epochs == 0) return OUTFIT_OK;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  const size_t n_ent = n_orbits * n_epochs;
  // propagated states [6][E][n] + their status: a separate slab (the observer table lives in ctx->scratch)
  if (ctx->nbody_state_bytes < 6 * n_ent * 8 + n_ent * 4 + 256) {
    if (ctx->nbody_state) { cudaFree(ctx->nbody_state); ctx->nbody_state = nullptr; ctx->nbody_state_bytes = 0; }
    if (cudaMalloc(&ctx->nbody_state, 6 * n_ent * 8 + n_ent * 4 + 256) != cudaSuccess) return fail(ctx, OUTFIT_E_ALLOC, "cudaMalloc(N-body states)");
    ctx->nbody_state_bytes = 6 * n_ent * 8 + n_ent * 4 + 256;
  }
  double *d_state = reinterpret_cast<double *>(ctx->nbody_state);
  int *d_sst = reinterpret_cast<int *>(d_state + 6 * n_ent);
  CK(cudaFuncSetAttribute(ephemeris_nbody_state_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNbSmemBytes));
  const NbCfgDev c{cfg->abs_tol, cfg->rel_tol, cfg->n_perturbers, cfg->max_steps ? cfg->max_steps : 100000u};
  ephemeris_nbody_state_kernel<<<(unsigned)((n_ent * 8 + kNbThreads - 1) / kNbThreads), kNbThreads, kNbSmemBytes, stream>>>(
      n_orbits, kind, epoch, elem, n_epochs, mjd_tt, c, gm, pert_pos, d_state, d_sst);
  size_t e_stride;
  double *d_table;
  int *d_ost;
  rc = ephemeris_observer_table(ctx, n_epochs, mjd_tt, mjd_ut1, epoch_body_fixed, nullptr, stream, &e_stride, &d_table, &d_ost);
  if (rc) return rc;
  const unsigned blocks = (unsigned)((n_orbits + kEphThreads - 1) / kEphThreads);
  if (ctx->aberration_order == 2)
    ephemeris_twobody_kernel<true, true><<<blocks, kEphThreads, 0, stream>>>(n_orbits, kind, epoch, elem, n_epochs, e_stride, mjd_tt, d_table,
                                                                            d_ost, out, status, d_state, d_sst);
  else
    ephemeris_twobody_kernel<false, true><<<blocks, kEphThreads, 0, stream>>>(n_orbits, kind, epoch, elem, n_epochs, e_stride, mjd_tt, d_table,
                                                                             d_ost, out, status, d_state, d_sst);
  CK(cudaGetLastError());
  return OUTFIT_OK;
}

// HOST buffers: an EphemerisRequest with several observers (as outfit_b200_ephemeris_request) under PropagatorKind::NBody
extern "C" int outfit_b200_ephemeris_nbody(OutfitCtx *ctx, size_t n_orbits, const int32_t *kind, const double *epoch,
                                           const double *elem, size_t n_observers, const double *observer_body_fixed,
                                           const uint64_t *epoch_offset, const double *mjd_tt, const double *mjd_ut1,
                                           const OutfitNBodyConfig *cfg, const double *gm, const double *pert_pos, double *out,
                                           int32_t *status) {
  if (!ctx || !cfg) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (n_observers == 0 || n_orbits == 0) return OUTFIT_OK;
  std::vector<double> bfp;
  size_t E = 0;
  int rc = expand_request(ctx, n_observers, observer_body_fixed, epoch_offset, bfp, &E);
  if (rc) return rc;
  if (E == 0) return OUTFIT_OK;
  if (!kind || !epoch || !elem || !mjd_tt || !mjd_ut1 || !gm || !pert_pos || !out || !status) return OUTFIT_E_INVALID_ARGUMENT;
  rc = check_nbody_cfg(ctx, cfg);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  rc = ensure_host_streams(ctx);
  if (rc) return rc;
  const size_t P = cfg->n_perturbers, n = n_orbits, n_ent = n * E;
  rc = ensure_arena(ctx, n * 4 + n * 8 * (1 + 6 + 3 * P) + 5 * E * 8 + P * 8 + 9 * n_ent * 8 + n_ent * 4 + 16 * 256);
  if (rc) return rc;
  cudaStream_t stream = ctx->compute_stream;
  ArenaPut A{ctx->arena, 0, stream};
  const int32_t *d_kind = (const int32_t *)A.put(kind, n * 4);
  const double *d_epoch = (const double *)A.put(epoch, n * 8);
  const double *d_elem = (const double *)A.put(elem, 6 * n * 8);
  const double *d_tt = (const double *)A.put(mjd_tt, E * 8);
  const double *d_ut = (const double *)A.put(mjd_ut1, E * 8);
  const double *d_bf = (const double *)A.put(bfp.data(), 3 * E * 8);
  const double *d_gm = (const double *)A.put(gm, P * 8);
  const double *d_pp = (const double *)A.put(pert_pos, 3 * P * n * 8);
  double *d_out = (double *)A.raw(9 * n_ent * 8);
  int32_t *d_st = (int32_t *)A.raw(n_ent * 4);
  if (A.err != cudaSuccess) { cudaStreamSynchronize(stream); return fail(ctx, OUTFIT_E_CUDA, "ephemeris_nbody: H2D", A.err); }
  rc = outfit_b200_ephemeris_nbody_device(ctx, n, d_kind, d_epoch, d_elem, E, d_tt, d_ut, d_bf, cfg, d_gm, d_pp, d_out, d_st, stream);
  if (rc == OUTFIT_OK) {
    cudaError_t e = cudaMemcpyAsync(out, d_out, 9 * n_ent * 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(status, d_st, n_ent * 4, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) rc = fail(ctx, OUTFIT_E_CUDA, "ephemeris_nbody: copy back / kernel", e);
  } else {
    cudaStreamSynchronize(stream);
  }
  return rc;
}

extern "C" int outfit_b200_selftest_arith(OutfitCtx *ctx, unsigned long long n, unsigned long long seed, int exp_range,
                                          unsigned long long out4[4]) {
  if (!ctx || !out4 || exp_range < 1 || exp_range > 500) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  unsigned long long *d = nullptr;
  CK(cudaMalloc(&d, 4 * sizeof(unsigned long long)));
  CK(cudaMemset(d, 0, 4 * sizeof(unsigned long long)));
  selftest_arith_kernel<<<ctx->sm_count * 8, 256>>>(n, seed, exp_range, d);
  cudaError_t e = cudaMemcpy(out4, d, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  // zero / special values: bf_sqrt(0) must be exactly 0
  cudaFree(d);
  if (e != cudaSuccess) return fail(ctx, OUTFIT_E_CUDA, "selftest_arith", e);
  return OUTFIT_OK;
}

// ---- FitLSQ with PropagatorKind::NBody (k_lsq_nbody.cuh) --------------------------------------------------------------
// DEVICE buffers.  Unlike the two-body entry this one SYNCHRONISES `cuda_stream`: the host drives the state machine
// trip by trip and reads the number of active trajectories (8 bytes) after every trip.
extern "C" int outfit_b200_fit_lsq_nbody_device(OutfitCtx *ctx, const OutfitLsqConfig *cfg, const OutfitNBodyConfig *nb,
                                                const double *gm, const double *pert_pos, const OutfitObsBatch *b,
                                                const OutfitIodResult *iod, OutfitLsqResult *out, OutfitObsFit *fit,
                                                void *cuda_stream) {
  if (!ctx || !cfg || !nb || !b) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  if (b->n_traj && (!iod || !out || !fit || !gm || !pert_pos))
    return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "fit_lsq_nbody: iod, out, fit, gm and perturber_pos are required");
  if (!ctx->have_eph) return fail(ctx, OUTFIT_E_NO_EPHEMERIS, "outfit_b200_load_ephemeris must be called first");
  int rc = check_nbody_cfg(ctx, nb);
  if (rc) return rc;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  const size_t n = b->n_obs, T = b->n_traj;
  if (T == 0) return OUTFIT_OK;
  if (T > 0xffffffffull) return fail(ctx, OUTFIT_E_UNSUPPORTED, "fit_lsq_nbody: more than 2^32 trajectories in one call");
  const bool have_geo = b->obs_geo_ecl != nullptr;
  const bool have_bf = b->observer_body_fixed && b->mjd_ut1;
  if (!have_geo && !have_bf) return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "need obs_geo_ecl or observer_body_fixed+mjd_ut1");
  // scratch: scorer[3][n] + tentative residuals[3][n] (+ geo[3][n] + helio[3][n] from pvobs) | records[n] | states[T] |
  //          status[n] | obs -> trajectory[n] | active counter
  const size_t planes = have_geo ? 6 : 12;
  const size_t off_rec = planes * n * sizeof(double);
  const size_t off_state = off_rec + n * sizeof(LsqNbRec);
  const size_t off_status = off_state + T * sizeof(LsqNbState);
  const size_t off_map = off_status + ((n + 1) & ~(size_t)1) * sizeof(int);
  const size_t off_cnt = off_map + ((n + 1) & ~(size_t)1) * sizeof(unsigned);
  rc = ensure_scratch(ctx, off_cnt + 256);
  if (rc) return rc;
  unsigned char *base = reinterpret_cast<unsigned char *>(ctx->scratch);
  double *d_scorer = reinterpret_cast<double *>(base);
  double *d_tmp = d_scorer + 3 * n;
  LsqNbRec *d_rec = reinterpret_cast<LsqNbRec *>(base + off_rec);
  LsqNbState *d_state = reinterpret_cast<LsqNbState *>(base + off_state);
  int *d_status = reinterpret_cast<int *>(base + off_status);
  unsigned *d_map = reinterpret_cast<unsigned *>(base + off_map);
  unsigned long long *d_active = reinterpret_cast<unsigned long long *>(base + off_cnt);
  const double *d_geo = b->obs_geo_ecl;
  const int tpb = 128;
  const unsigned gblocks = (unsigned)((n + tpb - 1) / tpb);
  if (!have_geo) {
    double *geo = d_scorer + 6 * n, *helio = d_scorer + 9 * n;
    if (n) observer_cache_kernel<<<gblocks, tpb, 0, stream>>>(ctx->eph, n, b->mjd_tt, b->mjd_ut1, b->observer_body_fixed, geo, helio, d_status);
    d_geo = geo;
  }
  if (n) scorer_observer_kernel<<<gblocks, tpb, 0, stream>>>(ctx->eph, n, b->mjd_tt, d_geo, d_scorer, d_status);
  LsqBatchDev B;
  B.n_traj = T; B.n_obs = n; B.traj_offset = (const unsigned long long *)b->traj_offset;
  B.mjd_tt = b->mjd_tt; B.ra = b->ra; B.dec = b->dec; B.sigma_ra = b->sigma_ra; B.sigma_dec = b->sigma_dec;
  B.scorer = d_scorer; B.obs_status = d_status;
  const LsqCfgDev C = to_lsq_dev(*cfg);
  const NbCfgDev nc{nb->abs_tol, nb->rel_tol, nb->n_perturbers, nb->max_steps ? nb->max_steps : 100000u};
  CK(cudaFuncSetAttribute(lsqnb_partials_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNbSmemBytes));
  CK(cudaMemsetAsync(d_active, 0, sizeof(unsigned long long), stream));
  lsqnb_init_kernel<<<(unsigned)((T + 127) / 128), 128, 0, stream>>>(B, C, iod, out, fit, d_state, d_map, d_active);
  unsigned long long active = 0;
  CK(cudaMemcpyAsync(&active, d_active, sizeof active, cudaMemcpyDeviceToHost, stream));
  CK(cudaStreamSynchronize(stream));
  // every trip either ends a trajectory or advances its (outer, inner) counters: bounded by the configuration
  const unsigned long long lim = 1ull << 31;  // saturating: absurd iteration caps must not wrap the bound
  const unsigned long long a_ = cfg->max_newton_iterations < lim ? cfg->max_newton_iterations + 2 : lim;
  const unsigned long long b_ = cfg->max_outlier_rejection_passes < lim ? cfg->max_outlier_rejection_passes + 2 : lim;
  const unsigned long long max_trips = a_ * b_ + 2;
  const size_t pthreads = n * 8;
  for (unsigned long long trip = 0; active != 0 && trip < max_trips; ++trip) {
    CK(cudaMemsetAsync(d_active, 0, sizeof(unsigned long long), stream));
    lsqnb_partials_kernel<<<(unsigned)((pthreads + kNbThreads - 1) / kNbThreads), kNbThreads, kNbSmemBytes, stream>>>(
        B, nc, gm, pert_pos, d_state, fit, d_map, d_rec);
    lsqnb_step_kernel<<<(unsigned)((T + 63) / 64), 64, 0, stream>>>(B, C, iod, out, fit, d_tmp, d_state, d_rec, d_active);
    CK(cudaMemcpyAsync(&active, d_active, sizeof active, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
  }
  CK(cudaGetLastError());
  if (active != 0) return fail(ctx, OUTFIT_E_CUDA, "fit_lsq_nbody: the state machine did not end within its trip bound");
  return OUTFIT_OK;
}

// HOST buffers, trajectories [tb, te) of the batch: their observations, IOD records and perturbers go up, the device
// entry runs, the results come back (iod / out = &records[tb], fit = &fit[0], pert_pos [P][3][hb->n_traj]).
static int fit_lsq_nbody_range(OutfitCtx *ctx, const OutfitLsqConfig *cfg, const OutfitNBodyConfig *nb, const double *gm,
                               const double *pert_pos, const OutfitObsBatch *hb, size_t tb, size_t te,
                               const OutfitIodResult *iod, OutfitLsqResult *out, OutfitObsFit *fit) {
  const size_t T = te - tb;
  if (T == 0) return OUTFIT_OK;
  const size_t o0 = hb->traj_offset[tb], n = hb->traj_offset[te] - o0, n_all = hb->n_obs, T_all = hb->n_traj;
  const bool have_geo = hb->obs_geo_ecl != nullptr;
  int rc = ensure_host_streams(ctx);
  if (rc) return rc;
  cudaStream_t stream = ctx->compute_stream;
  const size_t P = nb->n_perturbers;
  const size_t bytes = (T + 1) * 8 + 5 * n * 8 + (have_geo ? 3 : 4) * n * 8 + T * sizeof(OutfitIodResult) + T * sizeof(OutfitLsqResult) +
                       n * sizeof(OutfitObsFit) + P * 8 + P * 3 * T * 8 + 20 * 256;
  rc = ensure_arena(ctx, bytes);
  if (rc) return rc;
  ArenaPut A{ctx->arena, 0, stream};
  OutfitObsBatch db = *hb;
  db.n_traj = T; db.n_obs = n;
  if (o0 == 0) {
    db.traj_offset = (const uint64_t *)A.put(hb->traj_offset + tb, (T + 1) * 8);
  } else {
    rc = ensure_host_staging(ctx, (T + 1) * 8);
    if (rc) return rc;
    uint64_t *h = reinterpret_cast<uint64_t *>(ctx->h_scratch);
    for (size_t t = 0; t <= T; ++t) h[t] = hb->traj_offset[tb + t] - o0;
    db.traj_offset = (const uint64_t *)A.put(h, (T + 1) * 8);
  }
  db.mjd_tt = (const double *)A.put(hb->mjd_tt + o0, n * 8);
  db.ra = (const double *)A.put(hb->ra + o0, n * 8);
  db.dec = (const double *)A.put(hb->dec + o0, n * 8);
  db.sigma_ra = (const double *)A.put(hb->sigma_ra + o0, n * 8);
  db.sigma_dec = (const double *)A.put(hb->sigma_dec + o0, n * 8);
  db.obs_helio_equ = nullptr; db.noise_z = nullptr; db.traj_seed = nullptr;
  if (have_geo) {
    db.obs_geo_ecl = (const double *)A.put_planes(hb->obs_geo_ecl, n_all, o0, n);
    db.observer_body_fixed = nullptr; db.mjd_ut1 = nullptr;
  } else {
    db.obs_geo_ecl = nullptr;
    db.observer_body_fixed = (const double *)A.put_planes(hb->observer_body_fixed, n_all, o0, n);
    db.mjd_ut1 = (const double *)A.put(hb->mjd_ut1 + o0, n * 8);
  }
  const OutfitIodResult *d_iod = (const OutfitIodResult *)A.put(iod, T * sizeof(OutfitIodResult));
  const double *d_gm = (const double *)A.put(gm, P * 8);
  const double *d_pos = (const double *)A.put_planes(pert_pos, T_all, tb, T, (int)(3 * P));
  OutfitLsqResult *d_out = (OutfitLsqResult *)A.raw(T * sizeof(OutfitLsqResult));
  OutfitObsFit *d_fit = (OutfitObsFit *)A.raw(n * sizeof(OutfitObsFit));
  if (A.err != cudaSuccess) { cudaStreamSynchronize(stream); return fail(ctx, OUTFIT_E_CUDA, "fit_lsq_nbody: H2D", A.err); }
  rc = outfit_b200_fit_lsq_nbody_device(ctx, cfg, nb, d_gm, d_pos, &db, d_iod, d_out, d_fit, stream);
  if (rc != OUTFIT_OK) { cudaStreamSynchronize(stream); return rc; }
  cudaError_t e = cudaMemcpyAsync(out, d_out, T * sizeof(OutfitLsqResult), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess && fit) e = cudaMemcpyAsync(fit + o0, d_fit, n * sizeof(OutfitObsFit), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return fail(ctx, OUTFIT_E_CUDA, "fit_lsq_nbody: copy back / kernel", e);
  return OUTFIT_OK;
}

static int check_lsq_nbody_host_args(OutfitCtx *ctx, const OutfitLsqConfig *cfg, const OutfitNBodyConfig *nb, const double *gm,
                                     const double *pert_pos, const OutfitObsBatch *hb, const OutfitIodResult *iod, const void *out) {
  if (!ctx || !cfg || !nb || !hb) return OUTFIT_E_INVALID_ARGUMENT;
  if (hb->n_traj == 0) return OUTFIT_OK;
  if (!iod || !out || !gm || !pert_pos) return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "fit_lsq_nbody: iod, out, gm and perturber_pos are required");
  if (!hb->traj_offset || !hb->mjd_tt || !hb->ra || !hb->dec || !hb->sigma_ra || !hb->sigma_dec)
    return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "NULL observation array");
  if (!hb->obs_geo_ecl && !(hb->observer_body_fixed && hb->mjd_ut1))
    return fail(ctx, OUTFIT_E_INVALID_ARGUMENT, "need obs_geo_ecl or observer_body_fixed+mjd_ut1");
  const int rc = check_nbody_cfg(ctx, nb);
  if (rc) return rc;
  return check_offsets(ctx, hb, 0, hb->n_traj, nullptr);
}

extern "C" int outfit_b200_fit_lsq_nbody(OutfitCtx *ctx, const OutfitLsqConfig *cfg, const OutfitNBodyConfig *nb,
                                         const double *gm, const double *pert_pos, const OutfitObsBatch *hb,
                                         const OutfitIodResult *iod, OutfitLsqResult *out, OutfitObsFit *fit) {
  if (!ctx) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  const int rc = check_lsq_nbody_host_args(ctx, cfg, nb, gm, pert_pos, hb, iod, out);
  if (rc || hb->n_traj == 0) return rc;
  CK(cudaSetDevice(ctx->device));
  return fit_lsq_nbody_range(ctx, cfg, nb, gm, pert_pos, hb, 0, hb->n_traj, iod, out, fit);
}

// one call, every GPU of the group: trajectory ranges of equal observation counts (the integrations dominate)
extern "C" int outfit_b200_group_fit_lsq_nbody(OutfitGroup *g, const OutfitLsqConfig *cfg, const OutfitNBodyConfig *nb,
                                               const double *gm, const double *pert_pos, const OutfitObsBatch *hb,
                                               const OutfitIodResult *iod, OutfitLsqResult *out, OutfitObsFit *fit) {
  if (!g || g->ctx.empty()) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(g->mu);
  const int rc = check_lsq_nbody_host_args(g->ctx[0], cfg, nb, gm, pert_pos, hb, iod, out);
  if (rc) { g->last_error = g->ctx[0]->last_error; return rc; }
  if (hb->n_traj == 0) return OUTFIT_OK;
  std::vector<unsigned long long> cuts(g->ctx.size() + 1, 0ull);
  outfit_b200_shard_ranges(hb->n_traj, hb->traj_offset, 1u, 0ull, (int)g->ctx.size(), reinterpret_cast<uint64_t *>(cuts.data()));
  return group_run(g, cuts, [&](size_t, OutfitCtx *c, unsigned long long tb, unsigned long long te) {
    return fit_lsq_nbody_range(c, cfg, nb, gm, pert_pos, hb, tb, te, iod + tb, out + tb, fit);
  });
}

extern "C" int outfit_b200_measure_fp64_peak(OutfitCtx *ctx, double *flops_per_s) {
  if (!ctx || !flops_per_s) return OUTFIT_E_INVALID_ARGUMENT;
  std::lock_guard<std::recursive_mutex> lock(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  double *sink = nullptr;
  CK(cudaMalloc(&sink, 8));
  const int iters = 1 << 15, threads = 256;
  const int blocks = ctx->sm_count * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  fp64_peak_kernel<<<blocks, threads>>>(sink, iters);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    fp64_peak_kernel<<<blocks, threads>>>(sink, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8.0 * (double)iters * (double)threads * (double)blocks / (ms * 1e-3);
    if (fl > best) best = fl;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  CK(cudaGetLastError());
  *flops_per_s = best;
  return OUTFIT_OK;
}
