/* oo_triplets.c -- ORACLE (test infrastructure only): triplet enumeration and best-K selection.
 * Restates src/initial_orbit_determination/triplet_generation/{index_generator,mod}.rs, plus the
 * behaviour of std::collections::BinaryHeap (push = sift_up, pop = swap-last + sift_down_to_bottom
 * + sift_up) and of slice::sort_unstable_by for short slices (insertion sort for len <= 20; for
 * longer slices the order among EXACTLY equal weights is implementation-defined in the reference
 * too -- a stable merge is used here, documented as unpinned). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oo.h"
#include "oo_linalg.h"

/* index_generator.rs:66-75 */
size_t oo_downsample_uniform_with_edges(size_t n, size_t max_keep, size_t *keep) {
  if (n == 0) return 0;
  if (max_keep >= n) {
    for (size_t i = 0; i < n; i++) keep[i] = i;
    return n;
  }
  if (max_keep <= 3) {
    keep[0] = 0; keep[1] = n / 2; keep[2] = n - 1;
    return 3;
  }
  for (size_t i = 0; i < max_keep; i++) keep[i] = i * (n - 1) / (max_keep - 1);
  return max_keep;
}

typedef struct { size_t lo, hi; } last_window;
/* index_generator.rs:94-114 */
static last_window window_compute(size_t anchor, const double *ep, size_t n, double dt_min,
                                  double dt_max) {
  double t0 = ep[anchor];
  size_t lo = anchor + 2;
  while (lo < n && ep[lo] - t0 < dt_min) lo++;
  size_t hi = (lo == 0 ? 0 : lo - 1);
  if (hi < anchor + 1) hi = anchor + 1;
  while (hi + 1 < n && ep[hi + 1] - t0 <= dt_max) hi++;
  last_window w = {lo, hi};
  return w;
}
static int window_empty(last_window w, size_t anchor, size_t n) {
  return w.lo >= n || w.lo > w.hi || w.hi <= anchor + 1;
}

typedef struct {
  const double *ep;
  size_t n, anchor, middle, last;
  last_window w;
  double dt_min, dt_max;
} tgen;
static size_t zmax(size_t a, size_t b) { return a > b ? a : b; }
/* index_generator.rs:161-180 */
static void tgen_init(tgen *g, const double *ep, size_t n, double dt_min, double dt_max) {
  g->ep = ep; g->n = n; g->dt_min = dt_min; g->dt_max = dt_max;
  if (n >= 3) g->w = window_compute(0, ep, n, dt_min, dt_max);
  else { g->w.lo = n; g->w.hi = 0; }
  g->anchor = 0; g->middle = 1; g->last = zmax(g->w.lo, 2);
}
/* index_generator.rs:211-220 */
static int tgen_advance_anchor(tgen *g) {
  g->anchor += 1;
  if (g->anchor + 2 >= g->n) return 0;
  g->w = window_compute(g->anchor, g->ep, g->n, g->dt_min, g->dt_max);
  g->middle = g->anchor + 1;
  g->last = zmax(g->w.lo, g->middle + 1);
  return 1;
}
/* index_generator.rs:231-270 */
static int tgen_next(tgen *g, size_t *i, size_t *j, size_t *k) {
  for (;;) {
    if (g->anchor + 2 >= g->n) return 0;
    if (window_empty(g->w, g->anchor, g->n)) {
      if (!tgen_advance_anchor(g)) return 0;
      continue;
    }
    if (g->middle >= g->w.hi) {
      if (!tgen_advance_anchor(g)) return 0;
      continue;
    }
    if (g->last <= g->middle) g->last = zmax(g->w.lo, g->middle + 1);
    if (g->last > g->w.hi) {
      g->middle += 1;
      g->last = zmax(g->w.lo, g->middle + 1);
      continue;
    }
    *i = g->anchor; *j = g->middle; *k = g->last;
    g->last += 1;
    return 1;
  }
}

size_t oo_enumerate_triplets(const double *epochs, size_t n, double dt_min, double dt_max,
                             uint64_t *ijk, size_t cap) {
  tgen g;
  tgen_init(&g, epochs, n, dt_min, dt_max);
  size_t cnt = 0, i, j, k;
  while (cnt < cap && tgen_next(&g, &i, &j, &k)) {
    ijk[3 * cnt] = i; ijk[3 * cnt + 1] = j; ijk[3 * cnt + 2] = k;
    cnt++;
  }
  return cnt;
}

/* mod.rs:264-274 */
static double s_gap(double dt, double inv_dtw) {
  double r = dt * inv_dtw;
  if (r <= 1.0) return 1.0 / r;
  return 1.0 + r;
}
/* mod.rs:229-234 */
double oo_triplet_weight_with_inv(double t1, double t2, double t3, double inv_dtw) {
  double dt12 = t2 - t1, dt23 = t3 - t2;
  return s_gap(dt12, inv_dtw) + s_gap(dt23, inv_dtw);
}

/* Ord for WeightedTriplet, mod.rs:135-146: weight, falling back to indices when incomparable */
static int wt_cmp(const oo_weighted_triplet *a, const oo_weighted_triplet *b) {
  if (a->weight < b->weight) return -1;
  if (a->weight > b->weight) return 1;
  if (a->weight == b->weight) return 0;
  if (a->i != b->i) return a->i < b->i ? -1 : 1;
  if (a->j != b->j) return a->j < b->j ? -1 : 1;
  if (a->k != b->k) return a->k < b->k ? -1 : 1;
  return 0;
}
static void heap_sift_up(oo_weighted_triplet *d, size_t start, size_t pos) {
  oo_weighted_triplet el = d[pos];
  while (pos > start) {
    size_t parent = (pos - 1) / 2;
    if (wt_cmp(&el, &d[parent]) <= 0) break;
    d[pos] = d[parent];
    pos = parent;
  }
  d[pos] = el;
}
static void heap_sift_down_to_bottom(oo_weighted_triplet *d, size_t len) {
  size_t pos = 0, end = len;
  oo_weighted_triplet el = d[0];
  size_t child = 1;
  while (child <= (end >= 2 ? end - 2 : 0) && end >= 2) {
    if (wt_cmp(&d[child], &d[child + 1]) <= 0) child += 1;
    d[pos] = d[child];
    pos = child;
    child = 2 * pos + 1;
  }
  if (child == end - 1 && end >= 1) {
    d[pos] = d[child];
    pos = child;
  }
  d[pos] = el;
  heap_sift_up(d, 0, pos);
}

/* mod.rs:328-408 (phase 1 of generate_triplets) */
size_t oo_best_k_triplets(const double *mjd_tt, size_t n_obs, const oo_iod_params *p,
                          oo_weighted_triplet *out) {
  if (p->max_triplets == 0 || n_obs < 3) return 0;
  size_t *keep = (size_t *)malloc(sizeof(size_t) * (n_obs > 3 ? n_obs : 3));
  size_t nk = oo_downsample_uniform_with_edges(n_obs, (size_t)p->max_obs_for_triplets, keep);
  double *ep = (double *)malloc(sizeof(double) * nk);
  for (size_t i = 0; i < nk; i++) ep[i] = mjd_tt[keep[i]];
  size_t kcap = p->max_triplets;
  double inv_dtw = 1.0 / p->optimal_interval_time;
  oo_weighted_triplet *heap = (oo_weighted_triplet *)malloc(sizeof(oo_weighted_triplet) * (kcap + 1));
  size_t len = 0;
  tgen g;
  tgen_init(&g, ep, nk, p->dt_min, p->dt_max_triplet);
  size_t i, j, k;
  while (tgen_next(&g, &i, &j, &k)) {
    double w = oo_triplet_weight_with_inv(ep[i], ep[j], ep[k], inv_dtw);
    if (!isfinite(w)) continue;
    oo_weighted_triplet wt = {w, i, j, k};
    if (len < kcap) {
      heap[len] = wt;
      heap_sift_up(heap, 0, len);
      len++;
    } else if (len > 0 && w < heap[0].weight) {
      /* BinaryHeap::pop */
      oo_weighted_triplet lastel = heap[len - 1];
      len--;
      if (len > 0) {
        heap[0] = lastel;
        heap_sift_down_to_bottom(heap, len);
      }
      /* push */
      heap[len] = wt;
      heap_sift_up(heap, 0, len);
      len++;
    }
  }
  /* sort_unstable_by(weight): insertion sort (what std uses for len <= 20; stable) */
  for (size_t a = 1; a < len; a++) {
    oo_weighted_triplet el = heap[a];
    size_t b = a;
    while (b > 0 && el.weight < heap[b - 1].weight) {
      heap[b] = heap[b - 1];
      b--;
    }
    heap[b] = el;
  }
  memcpy(out, heap, sizeof(oo_weighted_triplet) * len);
  free(heap); free(ep); free(keep);
  return len;
}
