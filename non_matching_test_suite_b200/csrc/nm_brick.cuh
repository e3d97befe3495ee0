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
NMB_FN double grid_corner(double org, int i, double g) { return NMB_ADD(org, NMB_MUL((double)i, g)); }

template <int D> struct BrickShape {
  static constexpr int SH = D == 3 ? 2 : 3;            // brick edge = 2^SH grid cells
  static constexpr int EDGE = 1 << SH;
};

// One axis of one level: the exact range [L, H] of grid indices whose closed cell [c(i), c(i+1)] meets [qlo, qhi]
// (c = grid_corner, monotone in i).  `top` = largest admissible index.  false: empty.
NMB_FN bool axis_range(double org, double g, double inv, int top, double qlo, double qhi, int &L, int &H)
{
  // smallest i with c(i+1) >= qlo
  double t = (qlo - org) * inv;
  t = fmin(fmax(t, -4.0), (double)top + 4.0);
  int i = (int)ceil(t) - 1;
  for (int it = 0; it < 4 && grid_corner(org, i, g) >= qlo; ++it) --i;
  for (int it = 0; it < 4 && grid_corner(org, i + 1, g) < qlo; ++it) ++i;
  L = i < 0 ? 0 : i;
  // largest i with c(i) <= qhi
  t = (qhi - org) * inv;
  t = fmin(fmax(t, -4.0), (double)top + 4.0);
  i = (int)floor(t);
  for (int it = 0; it < 4 && grid_corner(org, i + 1, g) <= qhi; ++it) ++i;
  for (int it = 0; it < 4 && grid_corner(org, i, g) > qhi; ++it) --i;
  H = i > top ? top : i;
  return L <= H;
}

// bits of one brick that lie inside the index range (bit = lx | ly << SH | lz << 2 SH)
template <int D>
NMB_FN unsigned long long brick_query_mask(const int *L, const int *H, int bx, int by, int bz)
{
  constexpr int SH = BrickShape<D>::SH, E = BrickShape<D>::EDGE;
  int a0[3], a1[3];
  const int bb[3] = {bx, by, bz};
  for (int k = 0; k < D; ++k) {
    const int o = bb[k] << SH;
    a0[k] = L[k] > o ? L[k] - o : 0;
    a1[k] = H[k] < o + E - 1 ? H[k] - o : E - 1;
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

// Level of a cell of extent `ext` on the grid family g0 * 2^k: the smallest k with g0 * 2^k >= ext.
NMB_FN int level_of_extent(double ext, double g0)
{
  if (!(ext > g0)) return 0;
  int k = ilogb(ext) - ilogb(g0);
  if (k < 0) k = 0;
  if (k > 31) k = 31;
  while (k > 0 && ldexp(g0, k - 1) >= ext) --k;
  while (k < 31 && ldexp(g0, k) < ext) ++k;
  return k;
}

// One origin for every level: the lower corner of one coarsest cell, moved down by whole coarsest cells until it is not
// above the smallest lower corner.  On a tree every cell of every level then sits at org + i * g_k.
NMB_FN double aligned_origin(double lo_min, double lo_coarsest, double g_max)
{
  const double k = ceil((lo_coarsest - lo_min) / g_max);
  return NMB_ADD(lo_coarsest, -NMB_MUL(k, g_max));
}

}  // namespace nmbrick
