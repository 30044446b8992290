/* oo_linalg.h -- 3-vector / 3x3 helpers that reproduce nalgebra 0.34's operation order
 * (ORACLE: test infrastructure only).  Column-major matrices, m[3*c + r].
 *   dot3     : (a0*b0 + a1*b1) + a2*b2              (nalgebra base/blas.rs dotx, U3 special case)
 *   mat*vec  : (m_i0*v0 + m_i1*v1) + m_i2*v2        (gemv = axcpy column by column)
 *   mat*mat  : column j of C = A * column j of B    (gemm = gemv per column)
 *   norm     : sqrt(dot) ; Matrix3::norm = sqrt((d0 + d1) + d2), d_c = column c . column c
 *   inverse  : cofactor formula of Matrix::try_inverse (linalg/inverse.rs, dim 3)
 */
#ifndef OO_LINALG_H
#define OO_LINALG_H
#include <math.h>
#include "oo.h"

static inline double oo_dot3(const double a[3], const double b[3]) {
  return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}
static inline double oo_norm3(const double a[3]) { return sqrt(oo_dot3(a, a)); }
static inline void oo_cross3(const double a[3], const double b[3], double r[3]) {
  double x = a[1] * b[2] - a[2] * b[1];
  double y = a[2] * b[0] - a[0] * b[2];
  double z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void oo_matvec(const double m[9], const double v[3], double r[3]) {
  double t[3];
  for (int i = 0; i < 3; i++)
    t[i] = (OO_M(m, i, 0) * v[0] + OO_M(m, i, 1) * v[1]) + OO_M(m, i, 2) * v[2];
  r[0] = t[0]; r[1] = t[1]; r[2] = t[2];
}
static inline void oo_matmul(const double a[9], const double b[9], double c[9]) {
  double t[9];
  for (int j = 0; j < 3; j++) oo_matvec(a, &b[3 * j], &t[3 * j]);
  for (int i = 0; i < 9; i++) c[i] = t[i];
}
static inline void oo_transpose(const double a[9], double t[9]) {
  double r[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) OO_M(r, i, j) = OO_M(a, j, i);
  for (int i = 0; i < 9; i++) t[i] = r[i];
}
static inline double oo_matnorm(const double m[9]) {
  double res = 0.0;
  for (int c = 0; c < 3; c++) res += oo_dot3(&m[3 * c], &m[3 * c]);
  return sqrt(res);
}
static inline int oo_inverse3(const double m[9], double inv[9]) {
  double m11 = OO_M(m, 0, 0), m12 = OO_M(m, 0, 1), m13 = OO_M(m, 0, 2);
  double m21 = OO_M(m, 1, 0), m22 = OO_M(m, 1, 1), m23 = OO_M(m, 1, 2);
  double m31 = OO_M(m, 2, 0), m32 = OO_M(m, 2, 1), m33 = OO_M(m, 2, 2);
  double minor_m12_m23 = m22 * m33 - m32 * m23;
  double minor_m11_m23 = m21 * m33 - m31 * m23;
  double minor_m11_m22 = m21 * m32 - m31 * m22;
  double det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
  if (det == 0.0) return 0;
  OO_M(inv, 0, 0) = minor_m12_m23 / det;
  OO_M(inv, 0, 1) = (m13 * m32 - m33 * m12) / det;
  OO_M(inv, 0, 2) = (m12 * m23 - m22 * m13) / det;
  OO_M(inv, 1, 0) = -minor_m11_m23 / det;
  OO_M(inv, 1, 1) = (m11 * m33 - m31 * m13) / det;
  OO_M(inv, 1, 2) = (m13 * m21 - m23 * m11) / det;
  OO_M(inv, 2, 0) = minor_m11_m22 / det;
  OO_M(inv, 2, 1) = (m12 * m31 - m32 * m11) / det;
  OO_M(inv, 2, 2) = (m11 * m22 - m21 * m12) / det;
  return 1;
}
/* Rust f64::rem_euclid */
static inline double oo_rem_euclid(double x, double m) {
  double r = fmod(x, m);
  return (r < 0.0) ? r + fabs(m) : r;
}
static inline double oo_clamp(double x, double lo, double hi) {
  /* f64::clamp: NaN stays NaN */
  if (x < lo) return lo;
  if (x > hi) return hi;
  return x;
}
/* per-thread counters (defined in oo_counters.c) */
extern _Thread_local oo_counters oo_tls_cnt;
#endif
