/*
 * oo_rng.c -- CPU ORACLE (test infrastructure only) for the noise deviates of
 * GaussObs::realizations_iter (gauss.rs:323-387): SmallRng::seed_from_u64 + StandardNormal
 * (obs_dataset_api.rs:285-286, gauss.rs:355-364).
 *
 * rand 0.9 / rand_distr 0.5 are NOT vendored under /root/reference: this restates their published
 * algorithms (Xoshiro256++ seeded through SplitMix64; 256-layer ziggurat, R = 3.654152885361009,
 * V = 4.92867323399e-3, tables regenerated from the recurrence of the crate's ziggurat_tables.py;
 * StandardUniform / Open01 u64 -> f64 conversions).  PARITY UNPINNED against the crates (their only
 * known answers on this path need DE440); the integer generators are pinned by the public
 * splitmix64 / xoshiro256++ test vectors (tests/test_rng.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "oo.h"

#define ZIG_R 3.654152885361008796
#define ZIG_V 4.92867323399e-3
#define ZIG_N 256

static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
uint64_t oo_splitmix64_next(uint64_t *x) {
  *x += 0x9e3779b97f4a7c15ull;
  uint64_t z = *x;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
void oo_xoshiro_seed_from_u64(uint64_t seed, uint64_t s[4]) {
  for (int i = 0; i < 4; i++) s[i] = oo_splitmix64_next(&seed);
}
uint64_t oo_xoshiro_next(uint64_t s[4]) {
  const uint64_t result = rotl64(s[0] + s[3], 23) + s[0];
  const uint64_t t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
  s[2] ^= t;
  s[3] = rotl64(s[3], 45);
  return result;
}

static double zig_x[ZIG_N + 1], zig_f[ZIG_N + 1];
static int zig_ready = 0;
static void zig_init(void) {
  if (zig_ready) return;
  zig_x[0] = ZIG_V / exp(-ZIG_R * ZIG_R / 2.0);
  zig_x[1] = ZIG_R;
  for (int i = 2; i < ZIG_N; i++) {
    double last = zig_x[i - 1];
    zig_x[i] = sqrt(-2.0 * log(ZIG_V / last + exp(-last * last / 2.0)));
  }
  zig_x[ZIG_N] = 0.0;
  for (int i = 0; i <= ZIG_N; i++) zig_f[i] = exp(-zig_x[i] * zig_x[i] / 2.0);
  zig_ready = 1;
}
void oo_ziggurat_tables(double x[ZIG_N + 1], double f[ZIG_N + 1]) {
  zig_init();
  memcpy(x, zig_x, sizeof zig_x);
  memcpy(f, zig_f, sizeof zig_f);
}
static inline double bits_to_f64(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }

double oo_standard_normal(uint64_t s[4]) {
  zig_init();
  for (;;) {
    uint64_t bits = oo_xoshiro_next(s);
    unsigned i = (unsigned)(bits & 0xff);
    double u = bits_to_f64((1024ull << 52) | (bits >> 12)) - 3.0;
    double x = u * zig_x[i];
    if (fabs(x) < zig_x[i + 1]) return x;
    if (i == 0) {
      double xx, yy;
      do {
        double u1 = bits_to_f64((1023ull << 52) | (oo_xoshiro_next(s) >> 12)) - (1.0 - OO_EPS / 2.0);
        double u2 = bits_to_f64((1023ull << 52) | (oo_xoshiro_next(s) >> 12)) - (1.0 - OO_EPS / 2.0);
        xx = log(u1) / ZIG_R;
        yy = log(u2);
      } while (-2.0 * yy < xx * xx);
      return u < 0.0 ? xx - ZIG_R : ZIG_R - xx;
    }
    double uf = (double)(oo_xoshiro_next(s) >> 11) * (1.0 / 9007199254740992.0);
    if (zig_f[i + 1] + (zig_f[i] - zig_f[i + 1]) * uf < exp(-x * x / 2.0)) return x;
  }
}

/* the whole deviate stream of one trajectory, in draw order */
void oo_draw_noise(uint64_t seed, size_t n, double *out) {
  uint64_t s[4];
  oo_xoshiro_seed_from_u64(seed, s);
  for (size_t i = 0; i < n; i++) out[i] = oo_standard_normal(s);
}
