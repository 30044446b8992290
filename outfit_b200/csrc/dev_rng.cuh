// dev_rng.cuh -- on-device noise deviates for GaussObs::realizations_iter (gauss.rs:323-387): the
// per-trajectory generator of obs_dataset_api.rs:285-286, `SmallRng::seed_from_u64(base_seed ^
// traj_id.stable_hash())`, sampled with rand_distr's `StandardNormal`.
//
// The rand / rand_distr crates are NOT vendored under /root/reference; this restates their published
// algorithms (rand 0.9: SmallRng = Xoshiro256++ on 64-bit targets, seeded through SplitMix64;
// rand_distr 0.5: 256-layer ziggurat with R = 3.654152885361009, V = 4.92867323399e-3; u64 -> f64
// conversions of rand::distr::{StandardUniform, Open01}).  PARITY UNPINNED: the reference's only
// known answers for these streams need DE440, and the ziggurat tables are regenerated
// from the published recurrence rather than copied from the crate's literals.  The integer generator
// is pinned by the public xoshiro256++ / splitmix64 test vectors (tests/test_rng.py).  The host-drawn
// `noise_z` stream remains the strict-parity input; this path removes 48 B per noisy candidate of
// host -> device traffic when the caller only needs "a" reproducible normal stream.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ofb {

struct Xoshiro256pp {
  unsigned long long s0, s1, s2, s3;
};
__host__ __device__ inline unsigned long long rotl64(unsigned long long x, int k) { return (x << k) | (x >> (64 - k)); }
__host__ __device__ inline unsigned long long splitmix64_next(unsigned long long &x) {
  x += 0x9e3779b97f4a7c15ull;
  unsigned long long z = x;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__host__ __device__ inline Xoshiro256pp xoshiro_seed_from_u64(unsigned long long seed) {
  Xoshiro256pp r;
  r.s0 = splitmix64_next(seed); r.s1 = splitmix64_next(seed); r.s2 = splitmix64_next(seed); r.s3 = splitmix64_next(seed);
  return r;
}
__host__ __device__ inline unsigned long long xoshiro_next(Xoshiro256pp &r) {
  const unsigned long long result = rotl64(r.s0 + r.s3, 23) + r.s0;
  const unsigned long long t = r.s1 << 17;
  r.s2 ^= r.s0; r.s3 ^= r.s1; r.s1 ^= r.s2; r.s0 ^= r.s3;
  r.s2 ^= t;
  r.s3 = rotl64(r.s3, 45);
  return result;
}

constexpr double kZigR = 3.654152885361008796;
constexpr double kZigV = 4.92867323399e-3;
constexpr int kZigN = 256;

#ifdef __CUDACC__
__device__ __forceinline__ double bits_to_f64(unsigned long long b) { return __longlong_as_double((long long)b); }
// zig: x table [257] then f table [257] (shared memory)
__device__ __forceinline__ double standard_normal(Xoshiro256pp &r, const double *zig) {
  const double *xt = zig, *ft = zig + kZigN + 1;
  for (;;) {
    const unsigned long long bits = xoshiro_next(r);
    const unsigned i = (unsigned)(bits & 0xffull);
    // 52 random bits as a float in [2, 4), minus 3: uniform in [-1, 1)
    const double u = bits_to_f64((1024ull << 52) | (bits >> 12)) - 3.0;
    const double x = u * xt[i];
    if (fabs(x) < xt[i + 1]) return x;
    if (i == 0) {
      // tail (Marsaglia): x = -ln(U1)/R, y = -ln(U2) until 2y >= x^2, U in the open interval (0, 1)
      double xx, yy;
      do {
        const double u1 = bits_to_f64((1023ull << 52) | (xoshiro_next(r) >> 12)) - (1.0 - 2.220446049250313e-16 / 2.0);
        const double u2 = bits_to_f64((1023ull << 52) | (xoshiro_next(r) >> 12)) - (1.0 - 2.220446049250313e-16 / 2.0);
        xx = log(u1) / kZigR;
        yy = log(u2);
      } while (-2.0 * yy < xx * xx);
      return u < 0.0 ? xx - kZigR : kZigR - xx;
    }
    // wedge: uniform f64 in [0, 1) from 53 bits
    const double uf = (double)(xoshiro_next(r) >> 11) * (1.0 / 9007199254740992.0);
    if (ft[i + 1] + (ft[i] - ft[i + 1]) * uf < exp(-x * x / 2.0)) return x;
  }
}

// One thread per trajectory: its whole deviate stream, in draw order, into
// noise[t][max_triplets][n_noise][6] (ra0, ra1, ra2, dec0, dec1, dec2 per realization).
__global__ void __launch_bounds__(128)
noise_kernel(unsigned long long n_traj, const unsigned long long *__restrict__ seeds, const double *__restrict__ zig_g,
             unsigned per_traj_realizations, double *__restrict__ noise) {
  __shared__ double zig[2 * (kZigN + 1)];
  for (unsigned q = threadIdx.x; q < 2 * (kZigN + 1); q += blockDim.x) zig[q] = zig_g[q];
  __syncthreads();
  const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_traj) return;
  Xoshiro256pp r = xoshiro_seed_from_u64(seeds[t]);
  double2 *out = reinterpret_cast<double2 *>(noise + (size_t)t * per_traj_realizations * 6);
  for (unsigned q = 0; q < per_traj_realizations; ++q) {
    const double z0 = standard_normal(r, zig), z1 = standard_normal(r, zig), z2 = standard_normal(r, zig);
    const double z3 = standard_normal(r, zig), z4 = standard_normal(r, zig), z5 = standard_normal(r, zig);
    out[3 * q + 0] = make_double2(z0, z1);
    out[3 * q + 1] = make_double2(z2, z3);
    out[3 * q + 2] = make_double2(z4, z5);
  }
}
#endif

}  // namespace ofb
