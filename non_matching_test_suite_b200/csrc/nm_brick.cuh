// Brick index helpers shared by the kernels of nm_index.cu and the host-side unit test (tests/cpp/test_brick.cu):
// pure functions, no device state.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define NMB_ADD(a, b) __dadd_rn((a), (b))
#define NMB_MUL(a, b) __dmul_rn((a), (b))
#define NMB_FN __device__ __forceinline__
#else
static inline double nmb_add_(double a, double b) { volatile double r = a + b; return r; }
static inline double nmb_mul_(double a, double b) { volatile double r = a * b; return r; }
#define NMB_ADD(a, b) nmb_add_((a), (b))
#define NMB_MUL(a, b) nmb_mul_((a), (b))
#ifdef __CUDACC__
#define NMB_FN __host__ __device__ inline
#else
#define NMB_FN inline
#endif
#endif

namespace nmbrick {

NMB_FN unsigned long long ord_encode(double x)
{
  unsigned long long b;
  memcpy(&b, &x, 8);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
NMB_FN double ord_decode(unsigned long long k)
{
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
  double x;
  memcpy(&x, &b, 8);
  return x;
}
// corner i of the level grid; the same two roundings wherever it is evaluated
NMB_FN double grid_corner(double org, long long i, double g) { return NMB_ADD(org, NMB_MUL((double)i, g)); }

template <int D> struct BrickShape {
  static constexpr int SH = D == 3 ? 2 : 3;            // brick edge = 2^SH grid cells
  static constexpr int EDGE = 1 << SH;
};

// One axis of one level: the exact range [L, H] of grid indices whose closed cell [c(i), c(i+1)] meets [qlo, qhi]
// (c = grid_corner, monotone in i).  `top` = largest admissible index.  false: empty.
NMB_FN bool axis_range(double org, double g, double inv, long long top, double qlo, double qhi, long long &L, long long &H)
{
  // smallest i with c(i+1) >= qlo
  double t = (qlo - org) * inv;
  t = fmin(fmax(t, -4.0), (double)top + 4.0);
  long long i = (long long)ceil(t) - 1;
  for (int it = 0; it < 4 && grid_corner(org, i, g) >= qlo; ++it) --i;
  for (int it = 0; it < 4 && grid_corner(org, i + 1, g) < qlo; ++it) ++i;
  L = i < 0 ? 0 : i;
  // largest i with c(i) <= qhi
  t = (qhi - org) * inv;
  t = fmin(fmax(t, -4.0), (double)top + 4.0);
  i = (long long)floor(t);
  for (int it = 0; it < 4 && grid_corner(org, i + 1, g) <= qhi; ++it) ++i;
  for (int it = 0; it < 4 && grid_corner(org, i, g) > qhi; ++it) --i;
  H = i > top ? top : i;
  return L <= H;
}

// bits of one brick that lie inside the index range (bit = lx | ly << SH | lz << 2 SH)
template <int D>
NMB_FN unsigned long long brick_query_mask(const long long *L, const long long *H, long long bx, long long by, long long bz)
{
  constexpr int SH = BrickShape<D>::SH, E = BrickShape<D>::EDGE;
  int a0[3], a1[3];
  const long long bb[3] = {bx, by, bz};
  for (int k = 0; k < D; ++k) {
    const long long o = bb[k] << SH;
    a0[k] = (int)(L[k] > o ? L[k] - o : 0);
    a1[k] = (int)(H[k] < o + E - 1 ? H[k] - o : E - 1);
    if (a0[k] > a1[k]) return 0ULL;
  }
  if (D == 3) {
    const unsigned long long X = (unsigned long long)(((1u << (a1[0] + 1)) - (1u << a0[0])) & 0xFu) * 0x1111111111111111ULL;
    const unsigned long long Y = (unsigned long long)(((1u << (4 * (a1[1] + 1))) - (1u << (4 * a0[1]))) & 0xFFFFu) * 0x0001000100010001ULL;
    const unsigned long long Z = (a1[2] == 3 ? ~0ULL : ((1ULL << (16 * (a1[2] + 1))) - 1ULL)) & ~((1ULL << (16 * a0[2])) - 1ULL);
    return X & Y & Z;
  }
  const unsigned long long X = (unsigned long long)(((1u << (a1[0] + 1)) - (1u << a0[0])) & 0xFFu) * 0x0101010101010101ULL;
  const unsigned long long Y = (a1[1] == 7 ? ~0ULL : ((1ULL << (8 * (a1[1] + 1))) - 1ULL)) & ~((1ULL << (8 * a0[1])) - 1ULL);
  return X & Y;
}

}  // namespace nmbrick
