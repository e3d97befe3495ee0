// Exact orientation / in-circle signs on double inputs, device side.
//
// Needed for rule R2: the reference splits an immersed quad along the
// Delaunay diagonal of its xy-projection with CGAL's exact predicates
// (code/src/intersections.cc:83-84,464), and for non-planar quads the choice changes the
// surface.  A floating-point filter decides almost every quad; the rest is evaluated in
// 512-bit two's-complement integer arithmetic: every coordinate is m * 2^e with a 53-bit m,
// all coordinates of one quad are scaled by the smallest exponent present (spread <= 64 bits,
// otherwise the quad is reported as a tie), after which differences and products are exact
// integers of at most 476 bits.
#pragma once
#include <stdint.h>

namespace nmpred {

struct Big {
  uint32_t w[16];
};

__device__ inline void big_zero(Big &a)
{
#pragma unroll
  for (int i = 0; i < 16; ++i) a.w[i] = 0;
}
__device__ inline void big_neg(Big &a)
{
  uint64_t c = 1;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    uint64_t t = (uint64_t)(~a.w[i]) + c;
    a.w[i] = (uint32_t)t;
    c = t >> 32;
  }
}
__device__ inline Big big_add(const Big &a, const Big &b)
{
  Big r;
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    uint64_t t = (uint64_t)a.w[i] + b.w[i] + c;
    r.w[i] = (uint32_t)t;
    c = t >> 32;
  }
  return r;
}
__device__ inline Big big_sub(const Big &a, const Big &b)
{
  Big nb = b;
  big_neg(nb);
  return big_add(a, nb);
}
__device__ inline Big big_mul(const Big &a, const Big &b)
{  // low 512 bits; correct for two's complement when the true product fits
  Big r;
  big_zero(r);
  for (int i = 0; i < 16; ++i) {
    uint64_t c = 0;
    const uint64_t ai = a.w[i];
    if (ai == 0) continue;
    for (int j = 0; i + j < 16; ++j) {
      uint64_t t = ai * b.w[j] + r.w[i + j] + c;
      r.w[i + j] = (uint32_t)t;
      c = t >> 32;
    }
  }
  return r;
}
__device__ inline int big_sign(const Big &a)
{
  if (a.w[15] >> 31) return -1;
  uint32_t any = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) any |= a.w[i];
  return any ? 1 : 0;
}
// exponent of the last mantissa bit of x (x != 0): x = m * 2^e with integer 53-bit m
__device__ inline int ulp_exp(double x)
{
  int e;
  frexp(x, &e);
  return e - 53;
}
__device__ inline Big big_from(double x, int emin)
{
  Big r;
  big_zero(r);
  if (x == 0.0) return r;
  int e;
  double m = frexp(fabs(x), &e);
  uint64_t mant = (uint64_t)ldexp(m, 53);
  int sh = (e - 53) - emin;  // 0 .. 64
  int word = sh >> 5, bit = sh & 31;
  // mant (53 bits) shifted by `bit` spans at most 3 words
  uint64_t lo = mant << bit;
  uint64_t hi = bit ? (mant >> (64 - bit)) : 0;
  r.w[word] = (uint32_t)lo;
  r.w[word + 1] = (uint32_t)(lo >> 32);
  r.w[word + 2] = (uint32_t)hi;
  if (x < 0) big_neg(r);
  return r;
}

// returns false when the exponents spread over more than 64 bits
__device__ inline bool common_exponent(const double *const *P, int n, int &emin)
{
  int lo = 1 << 30, hi = -(1 << 30);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 2; ++k) {
      double x = P[i][k];
      if (x != 0.0) { int e = ulp_exp(x); lo = e < lo ? e : lo; hi = e > hi ? e : hi; }
    }
  if (hi < lo) { emin = 0; return true; }
  emin = lo;
  return hi - lo <= 64;
}

// sign of (b-a) x (c-a) in the xy plane; 2 = undecidable (exponent spread)
__device__ inline int orient2d(const double *a, const double *b, const double *c)
{
  const double l = (b[0] - a[0]) * (c[1] - a[1]), r = (b[1] - a[1]) * (c[0] - a[0]);
  const double det = l - r, bound = 8.0e-16 * (fabs(l) + fabs(r));
  if (det > bound) return 1;
  if (det < -bound) return -1;
  const double *P[3] = {a, b, c};
  int emin;
  if (!common_exponent(P, 3, emin)) return 2;
  Big ax = big_from(a[0], emin), ay = big_from(a[1], emin);
  Big bx = big_sub(big_from(b[0], emin), ax), by = big_sub(big_from(b[1], emin), ay);
  Big cx = big_sub(big_from(c[0], emin), ax), cy = big_sub(big_from(c[1], emin), ay);
  return big_sign(big_sub(big_mul(bx, cy), big_mul(by, cx)));
}

// > 0 iff d strictly inside the circle through a, b, c (counter-clockwise); 2 = undecidable
__device__ inline int incircle(const double *a, const double *b, const double *c, const double *d)
{
  const double adx = a[0] - d[0], ady = a[1] - d[1], bdx = b[0] - d[0], bdy = b[1] - d[1];
  const double cdx = c[0] - d[0], cdy = c[1] - d[1];
  const double al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;
  const double bc1 = bdx * cdy, bc2 = cdx * bdy, ca1 = cdx * ady, ca2 = adx * cdy, ab1 = adx * bdy, ab2 = bdx * ady;
  const double det = al * (bc1 - bc2) + bl * (ca1 - ca2) + cl * (ab1 - ab2);
  const double perm = al * (fabs(bc1) + fabs(bc2)) + bl * (fabs(ca1) + fabs(ca2)) + cl * (fabs(ab1) + fabs(ab2));
  const double bound = 2.0e-15 * perm;
  if (det > bound) return 1;
  if (det < -bound) return -1;
  const double *P[4] = {a, b, c, d};
  int emin;
  if (!common_exponent(P, 4, emin)) return 2;
  Big dx = big_from(d[0], emin), dy = big_from(d[1], emin);
  Big X[3], Y[3], L[3];
  for (int i = 0; i < 3; ++i) {
    X[i] = big_sub(big_from(P[i][0], emin), dx);
    Y[i] = big_sub(big_from(P[i][1], emin), dy);
    L[i] = big_add(big_mul(X[i], X[i]), big_mul(Y[i], Y[i]));
  }
  Big t0 = big_mul(L[0], big_sub(big_mul(X[1], Y[2]), big_mul(X[2], Y[1])));
  Big t1 = big_mul(L[1], big_sub(big_mul(X[2], Y[0]), big_mul(X[0], Y[2])));
  Big t2 = big_mul(L[2], big_sub(big_mul(X[0], Y[1]), big_mul(X[1], Y[0])));
  return big_sign(big_add(big_add(t0, t1), t2));
}

}  // namespace nmpred
