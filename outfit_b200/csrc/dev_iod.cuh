// dev_iod.cuh -- building blocks of the full-IOD pipeline: warp-per-trajectory triplet selection,
// and the per-candidate (one lane per (triplet, realization)) Gauss geometry / polynomial / root solve.
//
// Reference behaviour:
//   generate_triplets / best-K       triplet_generation/mod.rs:229-440, index_generator.rs:66-271
//   estimate_best_orbit              trajectory.rs:429-545  (sequential loop with running-best
//                                    pruning; reproduced here by an order-preserving warp fold)
//   rms_orbit_error / interval       trajectory.rs:294-427
//   prelim_orbit(_all)               gauss.rs:1119-1247
#pragma once
#include "dev_elements.cuh"

namespace ofb {

// ---- per-warp shared-memory view of one trajectory ------------------------------------------
struct TrajSmem {
  double *t;             // [n_obs] observation epochs of the trajectory
  double *heap_w;        // [max_triplets] best-K heap: weights
  unsigned *heap_x;      // [max_triplets] best-K heap: packed (i<<20 | j<<10 | k)
};

__device__ __forceinline__ unsigned keep_index(unsigned i, unsigned n, unsigned max_keep) {
  // downsample_uniform_with_edges (index_generator.rs:66-75)
  if (max_keep >= n) return i;
  if (max_keep <= 3) return i == 0 ? 0u : (i == 1 ? n / 2 : n - 1);
  return (unsigned)(((unsigned long long)i * (n - 1)) / (max_keep - 1));
}
__device__ __forceinline__ double s_gap(double dt, double inv_dtw) {
  const double r = dt * inv_dtw;
  return r <= 1.0 ? 1.0 / r : 1.0 + r;
}

// Warp-cooperative best-K triplet selection that reproduces the reference's container semantics
// exactly, because exact weight ties are systematic (for gaps longer than the optimal interval the
// weight is 2 + (t_k - t_i)/dtw whatever the middle index): candidates are produced in the
// generator's lexicographic (i<j<k) order, lanes compute 32 weights at a time, and the survivors are
// fed IN ORDER to a std::collections::BinaryHeap replica (push = sift_up; pop = swap-last +
// sift_down_to_bottom + sift_up; replacement only on strict `<`, mod.rs:386-401), followed by the
// stable insertion sort `sort_unstable_by` performs on short slices (mod.rs:405-407).
// Returns the number found (warp-uniform); leaves them, ascending, packed in sm.heap_x.
__device__ __forceinline__ bool heap_le(double wa, double wb) { return !(wa > wb); }  // a <= b by weight
__device__ __forceinline__ void heap_sift_up(double *hw, unsigned *hx, unsigned pos) {
  const double w = hw[pos];
  const unsigned x = hx[pos];
  while (pos > 0) {
    const unsigned parent = (pos - 1) >> 1;
    if (heap_le(w, hw[parent])) break;
    hw[pos] = hw[parent]; hx[pos] = hx[parent];
    pos = parent;
  }
  hw[pos] = w; hx[pos] = x;
}
__device__ __forceinline__ void heap_sift_down_to_bottom(double *hw, unsigned *hx, unsigned end) {
  const double w = hw[0];
  const unsigned x = hx[0];
  unsigned pos = 0, child = 1;
  while (end >= 2 && child <= end - 2) {
    if (heap_le(hw[child], hw[child + 1])) child += 1;
    hw[pos] = hw[child]; hx[pos] = hx[child];
    pos = child;
    child = 2 * pos + 1;
  }
  if (end >= 1 && child == end - 1) {
    hw[pos] = hw[child]; hx[pos] = hx[child];
    pos = child;
  }
  hw[pos] = w; hx[pos] = x;
  heap_sift_up(hw, hx, pos);
}

__device__ __noinline__ unsigned select_triplets(const TrajSmem &sm, unsigned n_obs,
                                                    const IodDevParams &P, unsigned lane) {
  if (P.max_triplets == 0 || n_obs < 3) return 0;
  const unsigned nr = P.max_obs_for_triplets >= n_obs ? n_obs : (P.max_obs_for_triplets <= 3 ? 3u : P.max_obs_for_triplets);
  const unsigned K = P.max_triplets;
  double *hw = sm.heap_w;
  unsigned *hx = sm.heap_x;
  unsigned heap_len = 0;   // warp-uniform
  double worst = INFINITY; // weight at the heap root once the heap is full (warp-uniform)
  // odometer over (i, j, k): this lane visits lexicographic indices lane, lane + 32, ...
  unsigned i = 0, j = 1, o = lane;
  bool done = false;
  for (;;) {
    while (!done && o >= nr - 1 - j) {
      o -= nr - 1 - j;
      ++j;
      if (j + 1 >= nr) { ++i; j = i + 1; if (i + 2 >= nr) done = true; }
    }
    if (__all_sync(0xffffffffu, done)) break;
    double wgt = INFINITY;
    unsigned packed = 0;
    bool cand = false;
    if (!done) {
      const unsigned k = j + 1 + o;
      const double ti = sm.t[keep_index(i, n_obs, P.max_obs_for_triplets)];
      const double tj = sm.t[keep_index(j, n_obs, P.max_obs_for_triplets)];
      const double tk = sm.t[keep_index(k, n_obs, P.max_obs_for_triplets)];
      const double span = tk - ti;
      if (span >= P.dt_min && span <= P.dt_max_triplet) {
        wgt = s_gap(tj - ti, P.inv_optimal_interval) + s_gap(tk - tj, P.inv_optimal_interval);
        cand = isfinite(wgt);
        packed = (i << 20) | (j << 10) | k;
      }
      o += 32;
    }
    // feed this chunk's survivors to the heap in enumeration (= lane) order
    unsigned mask = __ballot_sync(0xffffffffu, cand && (heap_len < K || wgt < worst));
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const double ws = __shfl_sync(0xffffffffu, wgt, src);
      const unsigned xs = __shfl_sync(0xffffffffu, packed, src);
      if (lane == 0) {
        if (heap_len < K) {
          hw[heap_len] = ws; hx[heap_len] = xs;
          heap_sift_up(hw, hx, heap_len);
          ++heap_len;
        } else if (ws < hw[0]) {
          // BinaryHeap::pop then push
          --heap_len;
          if (heap_len > 0) {
            hw[0] = hw[heap_len]; hx[0] = hx[heap_len];
            heap_sift_down_to_bottom(hw, hx, heap_len);
          }
          hw[heap_len] = ws; hx[heap_len] = xs;
          heap_sift_up(hw, hx, heap_len);
          ++heap_len;
        }
        worst = heap_len >= K ? hw[0] : INFINITY;
      }
      heap_len = __shfl_sync(0xffffffffu, heap_len, 0);
      worst = __shfl_sync(0xffffffffu, worst, 0);
      mask &= __ballot_sync(0xffffffffu, cand && (heap_len < K || wgt < worst));
    }
  }
  if (lane == 0) {
    // heap.into_vec() then insertion sort by weight (stable)
    for (unsigned a = 1; a < heap_len; ++a) {
      const double w = hw[a];
      const unsigned x = hx[a];
      unsigned b = a;
      while (b > 0 && w < hw[b - 1]) { hw[b] = hw[b - 1]; hx[b] = hx[b - 1]; --b; }
      hw[b] = w; hx[b] = x;
    }
  }
  __syncwarp();
  return heap_len;
}

// ---- the same selection, one THREAD per trajectory ---------------------------------------------------
// The std::BinaryHeap replica is inherently serial; run by lane 0 of a warp it issues ~30 k warp
// instructions per trajectory with one lane active (ncu r02l: 5.3 of 32 threads per instruction, issue
// slots 74 % busy).  With a thread per trajectory the same instructions serve 32 trajectories at once.
// The heap is a column of a [slot][thread] shared array (bank = thread, whatever the slot), the epochs
// are read through the read-only cache.  Identical operations on identical values: identical output.
struct HeapCol {
  double *w;     // &heap_w[threadIdx.x], stride = threads per block
  unsigned *x;   // &heap_x[threadIdx.x]
  unsigned stride;
  __device__ __forceinline__ double &W(unsigned i) const { return w[i * stride]; }
  __device__ __forceinline__ unsigned &X(unsigned i) const { return x[i * stride]; }
};
__device__ __forceinline__ void heapc_sift_up(const HeapCol &h, unsigned pos) {
  const double w = h.W(pos);
  const unsigned x = h.X(pos);
  while (pos > 0) {
    const unsigned parent = (pos - 1) >> 1;
    if (heap_le(w, h.W(parent))) break;
    h.W(pos) = h.W(parent); h.X(pos) = h.X(parent);
    pos = parent;
  }
  h.W(pos) = w; h.X(pos) = x;
}
__device__ __forceinline__ void heapc_sift_down_to_bottom(const HeapCol &h, unsigned end) {
  const double w = h.W(0);
  const unsigned x = h.X(0);
  unsigned pos = 0, child = 1;
  while (end >= 2 && child <= end - 2) {
    if (heap_le(h.W(child), h.W(child + 1))) child += 1;
    h.W(pos) = h.W(child); h.X(pos) = h.X(child);
    pos = child;
    child = 2 * pos + 1;
  }
  if (end >= 1 && child == end - 1) {
    h.W(pos) = h.W(child); h.X(pos) = h.X(child);
    pos = child;
  }
  h.W(pos) = w; h.X(pos) = x;
  heapc_sift_up(h, pos);
}
// returns the number found; leaves them, ascending, in the heap column
__device__ __forceinline__ unsigned select_triplets_thread(const HeapCol &h, const double *__restrict__ T, unsigned n_obs,
                                                           const IodDevParams &P) {
  if (P.max_triplets == 0 || n_obs < 3) return 0;
  const unsigned nr = P.max_obs_for_triplets >= n_obs ? n_obs : (P.max_obs_for_triplets <= 3 ? 3u : P.max_obs_for_triplets);
  const unsigned K = P.max_triplets;
  unsigned heap_len = 0;
  for (unsigned i = 0; i + 2 < nr; ++i) {
    const double ti = __ldg(T + keep_index(i, n_obs, P.max_obs_for_triplets));
    for (unsigned j = i + 1; j + 1 < nr; ++j) {
      const double tj = __ldg(T + keep_index(j, n_obs, P.max_obs_for_triplets));
      const double gap1 = s_gap(tj - ti, P.inv_optimal_interval);
      for (unsigned k = j + 1; k < nr; ++k) {
        const double tk = __ldg(T + keep_index(k, n_obs, P.max_obs_for_triplets));
        const double span = tk - ti;
        if (!(span >= P.dt_min && span <= P.dt_max_triplet)) continue;
        const double wgt = gap1 + s_gap(tk - tj, P.inv_optimal_interval);
        if (!isfinite(wgt)) continue;
        const unsigned packed = (i << 20) | (j << 10) | k;
        if (heap_len < K) {
          h.W(heap_len) = wgt; h.X(heap_len) = packed;
          heapc_sift_up(h, heap_len);
          ++heap_len;
        } else if (wgt < h.W(0)) {
          // BinaryHeap::pop then push
          --heap_len;
          if (heap_len > 0) {
            h.W(0) = h.W(heap_len); h.X(0) = h.X(heap_len);
            heapc_sift_down_to_bottom(h, heap_len);
          }
          h.W(heap_len) = wgt; h.X(heap_len) = packed;
          heapc_sift_up(h, heap_len);
          ++heap_len;
        }
      }
    }
  }
  // heap.into_vec() then insertion sort by weight (stable)
  for (unsigned a = 1; a < heap_len; ++a) {
    const double w = h.W(a);
    const unsigned x = h.X(a);
    unsigned b = a;
    while (b > 0 && w < h.W(b - 1)) { h.W(b) = h.W(b - 1); h.X(b) = h.X(b - 1); --b; }
    h.W(b) = w; h.X(b) = x;
  }
  return heap_len;
}

// ---- candidate geometry: gauss_prelim + unit matrix + cofactor inverse (gauss.rs:464-549) -------
// false <=> SingularDirectionMatrix
__device__ __noinline__ bool gauss_geometry(const Triplet &g, GaussGeom &gm) {
  gm.tau1 = kGaussK * (g.t[0] - g.t[1]);
  gm.tau3 = kGaussK * (g.t[2] - g.t[1]);
  const double tau13 = gm.tau3 - gm.tau1;
  gm.a0 = gm.tau3 / tau13;
  gm.a2 = -(gm.tau1 / tau13);
  gm.b0 = gm.a0 * (tau13 * tau13 - gm.tau3 * gm.tau3) / 6.0;
  gm.b2 = gm.a2 * (tau13 * tau13 - gm.tau1 * gm.tau1) / 6.0;
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    double sr, cr, sd, cd;
    sincos(g.ra[c], &sr, &cr);
    sincos(g.dec[c], &sd, &cd);
    gm.S[c] = V3{cr * cd, sr * cd, sd};
  }
  const double m11 = gm.S[0].x, m12 = gm.S[1].x, m13 = gm.S[2].x;
  const double m21 = gm.S[0].y, m22 = gm.S[1].y, m23 = gm.S[2].y;
  const double m31 = gm.S[0].z, m32 = gm.S[1].z, m33 = gm.S[2].z;
  const double mi1 = m22 * m33 - m32 * m23;
  const double mi2 = m21 * m33 - m31 * m23;
  const double mi3 = m21 * m32 - m31 * m22;
  const double det = m11 * mi1 - m12 * mi2 + m13 * mi3;
  if (det == 0.0) return false;
  gm.SiR[0] = V3{mi1 / det, (m13 * m32 - m33 * m12) / det, (m12 * m23 - m22 * m13) / det};
  gm.SiR[1] = V3{-mi2 / det, (m11 * m33 - m31 * m13) / det, (m13 * m21 - m23 * m11) / det};
  gm.SiR[2] = V3{mi3 / det, (m12 * m31 - m32 * m11) / det, (m11 * m22 - m21 * m12) / det};
  return true;
}

// coeff_eight_poly (gauss.rs:585-614) + Descartes prefilter (gauss.rs:214-240).
// false <=> no sign change <=> no positive real root (GaussNoRootsFound)
__device__ __forceinline__ bool gauss_polynomial(const Triplet &g, const GaussGeom &gm, double &c0, double &c3,
                                                 double &c6) {
  const V3 ra_v = V3{(g.R[0].x * gm.a0 + g.R[1].x * -1.0) + g.R[2].x * gm.a2,
                     (g.R[0].y * gm.a0 + g.R[1].y * -1.0) + g.R[2].y * gm.a2,
                     (g.R[0].z * gm.a0 + g.R[1].z * -1.0) + g.R[2].z * gm.a2};
  const V3 rb_v = V3{(g.R[0].x * gm.b0 + g.R[1].x * 0.0) + g.R[2].x * gm.b2,
                     (g.R[0].y * gm.b0 + g.R[1].y * 0.0) + g.R[2].y * gm.b2,
                     (g.R[0].z * gm.b0 + g.R[1].z * 0.0) + g.R[2].z * gm.b2};
  const double a2s = dot(gm.SiR[1], ra_v);
  const double b2s = dot(gm.SiR[1], rb_v);
  const double r22 = dot(g.R[1], g.R[1]);
  const double s2r2 = dot(gm.S[1], g.R[1]);
  c6 = -(a2s * a2s) - r22 - (2.0 * a2s * s2r2);
  c3 = -(2.0 * b2s * (a2s + s2r2));
  c0 = -(b2s * b2s);
  int last = 1, count = 0;
  const double cs[3] = {c6, c3, c0};
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int cur = cs[q] == 0.0 ? 0 : (signbit(cs[q]) ? -1 : 1);
    if (cur == 0) continue;
    if (cur != last) ++count;
    last = cur;
  }
  return count != 0;
}

}  // namespace ofb
