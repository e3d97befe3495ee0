// Subsystem (1): bounding-box candidate pairs on the device.
//
// Replaces the Boost R-tree query loop of the reference
//   for (immersed_box, immersed_cell) in immersed_tree:
//     for (space_box, space_cell) in space_tree | queried(intersects(immersed_box))
// (code/src/coupling_utilities.cc:40-55) with a multi-level uniform grid hash over the
// background boxes.  The result is the exact set { (im, bg) : closed boxes overlap on the raw
// doubles } -- the grid only produces a superset, the final test is the comparison itself.
//
// Grid: level k has spacing g0 * 2^k (g0 = smallest background cell extent); a background
// cell lives on the smallest level whose spacing covers its extent.  Cell index functions are
// monotone in the coordinate, so two boxes that overlap as closed intervals always share a
// grid cell -- no epsilon is involved:
//     background cell : [ floor(t(lo)), max(floor(t(lo)), ceil(t(hi)) - 1) ]
//     query box       : [ ceil(t(lo)) - 1, floor(t(hi)) ]            t(x) = (x - org) * inv_k
// (the asymmetric ends make a cell of a grid-aligned tree occupy exactly one grid cell while a
// query box that only touches it still finds it).  A pair seen from several grid cells is
// reported from the cell whose index is the per-axis max of the two lower ends.
#include <cub/cub.cuh>

#include <thrust/iterator/transform_iterator.h>
#include <thrust/iterator/transform_output_iterator.h>
#include <thrust/iterator/zip_iterator.h>
#include <thrust/tuple.h>

#include "nm_common.cuh"
#include "nm_brick.cuh"

namespace {

struct GridDev {
  int d;
  double org[3];
  double inv[32];
  double g0;
  uint32_t level_mask;
  const HashSlot *hash;
  uint32_t hash_mask;
  const uint32_t *next;     // chain of the entries of one grid cell: next[entry id]
  const uint32_t *extra_bg; // background id of entry ids >= n_bg (cells spanning several grid cells)
  uint32_t n_bg;
  int idx_bits;
};

__host__ __device__ inline unsigned long long pack_key(int d, int level, long long ix, long long iy, long long iz)
{
  if (d == 2) return ((unsigned long long)level << 58) | ((unsigned long long)iy << 29) | (unsigned long long)ix;
  return ((unsigned long long)level << 57) | ((unsigned long long)iz << 38) | ((unsigned long long)iy << 19) | (unsigned long long)ix;
}

__device__ inline unsigned long long mix64(unsigned long long k)
{
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}

__device__ inline long long g_floor(double x, double org, double inv) { return (long long)floor((x - org) * inv); }
__device__ inline long long g_ceilm1(double x, double org, double inv) { return (long long)ceil((x - org) * inv) - 1; }

__device__ inline int level_of(double ext, double g0)
{
  int k = 0;
  double g = g0;
  while (g < ext && k < 31) { g *= 2.0; ++k; }
  return k;
}

// ---------------------------------------------------------------------------------------
// background layout kernels
// ---------------------------------------------------------------------------------------
// verts [n][2^d][d] -> box [n][8]; flags cells that are not axis-aligned boxes in deal.II
// vertex order (vertex v: bit k selects the upper end along axis k).
template <int D>
__global__ void bg_boxes_from_verts(uint64_t n, const double *__restrict__ verts, double *__restrict__ box, int *bad)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int NV = 1 << D;
  const double *v = verts + i * NV * D;
  double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  bool ok = true;
#pragma unroll
  for (int k = 0; k < D; ++k) { lo[k] = v[k]; hi[k] = v[(NV - 1) * D + k]; ok &= lo[k] <= hi[k]; }
#pragma unroll
  for (int q = 0; q < NV; ++q)
#pragma unroll
    for (int k = 0; k < D; ++k) ok &= v[q * D + k] == (((q >> k) & 1) ? hi[k] : lo[k]);
  if (!ok) atomicAdd(bad, 1);
  double *b = box + i * 8;
  b[0] = lo[0]; b[1] = lo[1]; b[2] = lo[2]; b[3] = 0;
  b[4] = hi[0]; b[5] = hi[1]; b[6] = hi[2]; b[7] = 0;
}

__global__ void fill_line_of(uint64_t n_c, const uint32_t *__restrict__ c_dof, int32_t *__restrict__ line_of, uint64_t n_dofs, int *bad)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_c) return;
  uint32_t d = c_dof[i];
  if (d >= n_dofs) { atomicAdd(bad, 1); return; }
  line_of[d] = (int32_t)i;
}

__global__ void constrained_dofs_kernel(uint64_t n_dofs, const int32_t *__restrict__ line_of, uint32_t *__restrict__ out)
{
  uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d < n_dofs && line_of[d] >= 0) out[line_of[d]] = (uint32_t)d;
}

// per-block partial min of lo, min positive extent, max hi  -> out[block][8]
struct BoxStats {
  double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY}, g = INFINITY, gm = 0;
  double gml[3] = {0, 0, 0};   // lower corner of one cell of the largest extent (anchors the grid of every level)
  template <int D> __device__ __forceinline__ void add(const double *lo, const double *hi)
  {
    double e = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      mn[k] = fmin(mn[k], lo[k]);
      mx[k] = fmax(mx[k], hi[k]);
      e = fmax(e, hi[k] - lo[k]);
    }
    if (e > 0) g = fmin(g, e);
    if (e > gm) {
      gm = e;
#pragma unroll
      for (int k = 0; k < D; ++k) gml[k] = lo[k];
    }
  }
  __device__ inline void reduce_to(double *__restrict__ out) const;
};

constexpr int NM_STATS_W = 12;   // doubles per block: min lo (3), max hi (3), min extent, max extent, corner of a coarsest cell (3), pad

__device__ inline void BoxStats::reduce_to(double *__restrict__ out) const
{
  typedef cub::BlockReduce<double, 256> BR;
  __shared__ typename BR::TempStorage tmp;
  __shared__ double s_gm, s_gml[3];
  double r[8];
  for (int k = 0; k < 3; ++k) { r[k] = BR(tmp).Reduce(mn[k], cub::Min()); __syncthreads(); }
  for (int k = 0; k < 3; ++k) { r[3 + k] = BR(tmp).Reduce(mx[k], cub::Max()); __syncthreads(); }
  r[6] = BR(tmp).Reduce(g, cub::Min()); __syncthreads();
  r[7] = BR(tmp).Reduce(gm, cub::Max());
  if (threadIdx.x == 0) { s_gm = r[7]; s_gml[0] = s_gml[1] = s_gml[2] = 0; }
  __syncthreads();
  if (gm == s_gm && gm > 0) { s_gml[0] = gml[0]; s_gml[1] = gml[1]; s_gml[2] = gml[2]; }   // any one of the coarsest cells (benign race)
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < 8; ++k) out[blockIdx.x * NM_STATS_W + k] = r[k];
    for (int k = 0; k < 3; ++k) out[blockIdx.x * NM_STATS_W + 8 + k] = s_gml[k];
    out[blockIdx.x * NM_STATS_W + 11] = 0;
  }
}

// lo/hi corners -> 64-byte boxes, with the extent statistics of the grid index taken in the same pass
// (grid-stride, 256 threads per block, out[gridDim.x][8])
template <int D>
__global__ void __launch_bounds__(256) bg_boxes_from_lohi_stats(uint64_t n, const double *__restrict__ lo, const double *__restrict__ hi,
                                                                const double *__restrict__ edge, double *__restrict__ box, int *bad,
                                                                double *__restrict__ out)
{
  BoxStats S;
  int n_bad = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    double l[3] = {0, 0, 0}, h[3] = {0, 0, 0};
    const double e = edge ? edge[i] : 0.0;     // cubic cells: the upper corner is fl(lo + edge) on every axis
#pragma unroll
    for (int k = 0; k < D; ++k) { l[k] = lo[i * D + k]; h[k] = edge ? __dadd_rn(l[k], e) : hi[i * D + k]; n_bad += !(l[k] <= h[k]); }
    double4 *b = reinterpret_cast<double4 *>(box + i * 8);
    b[0] = make_double4(l[0], l[1], l[2], 0.0);
    b[1] = make_double4(h[0], h[1], h[2], 0.0);
    S.add<D>(l, h);
  }
  if (n_bad) atomicAdd(bad, n_bad);
  S.reduce_to(out);
}

template <int D>
__global__ void bg_stats(uint64_t n, const double *__restrict__ box, double *__restrict__ out)
{
  BoxStats S;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) S.add<D>(box + i * 8, box + i * 8 + 4);
  S.reduce_to(out);
}

// ---------------------------------------------------------------------------------------
// grid entries
// ---------------------------------------------------------------------------------------
template <int D>
__device__ inline void bg_range(const GridDev &G, const double *b, int &level, long long *L, long long *H)
{
  double e = 0;
#pragma unroll
  for (int k = 0; k < D; ++k) e = fmax(e, b[4 + k] - b[k]);
  level = level_of(e, G.g0);
  const double inv = G.inv[level];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    L[k] = g_floor(b[k], G.org[k], inv);
    long long h = g_ceilm1(b[4 + k], G.org[k], inv);
    H[k] = h > L[k] ? h : L[k];
  }
  for (int k = D; k < 3; ++k) L[k] = H[k] = 0;
}

// number of grid cells every background cell registers in (sum -> table size), levels present.
// Grid-stride with one atomic per block: a single counter cannot take an atomic per warp.
template <int D>
__global__ void entries_count(GridDev G, uint64_t n, const double *__restrict__ box, unsigned long long *total,
                              unsigned *level_mask, int *bad)
{
  unsigned long long c_sum = 0;
  unsigned bits = 0, n_bad = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    int level; long long L[3], H[3];
    bg_range<D>(G, box + i * 8, level, L, H);
    unsigned long long c = 1;
    const long long lim = 1LL << G.idx_bits;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      c *= (unsigned long long)(H[k] - L[k] + 1);
      ok &= L[k] >= 0 && H[k] < lim;
    }
    if (!ok || c > 64) { ++n_bad; c = 0; }
    c_sum += c;
    bits |= 1u << level;
  }
  typedef cub::BlockReduce<unsigned long long, 256> BR;
  __shared__ typename BR::TempStorage tmp;
  __shared__ unsigned s_bits, s_bad;
  if (threadIdx.x == 0) { s_bits = 0; s_bad = 0; }
  __syncthreads();
  const unsigned long long tot = BR(tmp).Sum(c_sum);
  bits = __reduce_or_sync(0xffffffffu, bits);
  n_bad = __reduce_add_sync(0xffffffffu, n_bad);
  if ((threadIdx.x & 31) == 0) { atomicOr(&s_bits, bits); if (n_bad) atomicAdd(&s_bad, n_bad); }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (tot) atomicAdd(total, tot);
    if (s_bits) atomicOr(level_mask, s_bits);
    if (s_bad) atomicAdd(bad, (int)s_bad);
  }
}

__global__ void hash_clear(HashSlot *h, uint32_t cap)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) { h[i].key = ~0ULL; h[i].start = 0xffffffffu; h[i].count = 1; }  // count of a claimed slot starts at 1
}

// Every background cell links itself into the slot of each grid cell it covers.  Entry id = the
// cell id for its first grid cell (the only one for a grid-aligned tree), ids >= n_bg for the rest.
// HashSlot.start is the head of the chain, HashSlot.count its length.
template <int D>
__global__ void index_insert(GridDev G, uint64_t n, const double *__restrict__ box, HashSlot *h, uint32_t *__restrict__ next,
                             uint32_t *__restrict__ extra_bg, unsigned *extra_cursor)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int level; long long L[3], H[3];
  bg_range<D>(G, box + i * 8, level, L, H);
  bool first = true;
  for (long long z = L[2]; z <= H[2]; ++z)
    for (long long y = L[1]; y <= H[1]; ++y)
      for (long long x = L[0]; x <= H[0]; ++x) {
        uint32_t e = (uint32_t)i;
        if (!first) { const unsigned k = atomicAdd(extra_cursor, 1u); extra_bg[k] = (uint32_t)i; e = G.n_bg + k; }
        first = false;
        const unsigned long long key = pack_key(D, level, x, y, z);
        uint32_t s = (uint32_t)mix64(key) & G.hash_mask;
        while (true) {
          const unsigned long long old = atomicCAS(&h[s].key, ~0ULL, key);
          if (old == ~0ULL || old == key) break;
          s = (s + 1) & G.hash_mask;
        }
        const uint32_t prev = atomicExch(&h[s].start, e);
        next[e] = prev;
        if (prev != 0xffffffffu) atomicAdd(&h[s].count, 1u);  // the first entry is already counted
      }
}

// ---------------------------------------------------------------------------------------
// candidate query
// ---------------------------------------------------------------------------------------
template <int D, typename Emit>
__device__ inline void query_box(const GridDev &G, const double *__restrict__ bg_box, const double *qlo, const double *qhi, Emit emit)
{
  const long long lim = (1LL << G.idx_bits) - 1;
  uint32_t lm = G.level_mask;
  while (lm) {
    const int level = __ffs(lm) - 1;
    lm &= lm - 1;
    const double inv = G.inv[level];
    long long L[3] = {0, 0, 0}, H[3] = {0, 0, 0};
    bool empty = false;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      L[k] = g_ceilm1(qlo[k], G.org[k], inv);
      H[k] = g_floor(qhi[k], G.org[k], inv);
      if (L[k] < 0) L[k] = 0;
      if (H[k] > lim) H[k] = lim;
      empty |= L[k] > H[k];
    }
    if (empty) continue;
    for (long long z = L[2]; z <= H[2]; ++z)
      for (long long y = L[1]; y <= H[1]; ++y)
        for (long long x = L[0]; x <= H[0]; ++x) {
          const unsigned long long key = pack_key(D, level, x, y, z);
          uint32_t s = (uint32_t)mix64(key) & G.hash_mask;
          uint32_t e = 0, count = 0;
          while (true) {
            const unsigned long long k = G.hash[s].key;
            if (k == key) { e = G.hash[s].start; count = G.hash[s].count; break; }
            if (k == ~0ULL) break;
            s = (s + 1) & G.hash_mask;
          }
          const long long cell[3] = {x, y, z};
          for (uint32_t c = 0; c < count; ++c) {
            const uint32_t b = e < G.n_bg ? e : G.extra_bg[e - G.n_bg];
            if (c + 1 < count) e = G.next[e];
            const double *bb = bg_box + (uint64_t)b * 8;
            bool hit = true;
#pragma unroll
            for (int k = 0; k < D; ++k) {
              const double blo = bb[k], bhi = bb[4 + k];
              hit &= (blo <= qhi[k]) && (qlo[k] <= bhi);  // the reference's closed-box test, verbatim semantics
              // report the pair from one grid cell only: the per-axis max of the lower ends
              long long lb = g_floor(blo, G.org[k], inv);
              long long first = lb > L[k] ? lb : L[k];
              hit &= cell[k] == first;
            }
            if (hit) emit(b);
          }
        }
  }
}

template <int D, int NVI>
__device__ inline void im_box(const double *__restrict__ v, double *lo, double *hi)
{
#pragma unroll
  for (int k = 0; k < D; ++k) { lo[k] = v[k]; hi[k] = v[k]; }
#pragma unroll
  for (int q = 1; q < NVI; ++q)
#pragma unroll
    for (int k = 0; k < D; ++k) { lo[k] = fmin(lo[k], v[q * D + k]); hi[k] = fmax(hi[k], v[q * D + k]); }
}

// Single probing pass: hits go, insertion-sorted, to a fixed-capacity row per immersed cell;
// the true count is recorded even when it exceeds the capacity.
template <int D, int CAP>
__global__ void cand_probe(GridDev G, uint64_t n_im, const double *__restrict__ im_verts, const double *__restrict__ bg_box,
                           uint32_t *__restrict__ cnt, uint32_t *__restrict__ tmp)
{
  uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_im) return;
  constexpr int NVI = 1 << (D - 1);
  double lo[3], hi[3];
  im_box<D, NVI>(im_verts + m * NVI * D, lo, hi);
  uint32_t c = 0;
  uint32_t *row = tmp + m * CAP;
  query_box<D>(G, bg_box, lo, hi, [&](uint32_t b) {
    if (c < CAP) {
      uint32_t j = c;
      while (j > 0 && row[j - 1] > b) { row[j] = row[j - 1]; --j; }
      row[j] = b;
    }
    ++c;
  });
  cnt[m] = c;
}

// Second pass: copy the rows to their final place (LANES threads per cell); cells that
// overflowed the capacity are probed again (rare).
template <int D, int CAP, int LANES>
__global__ void cand_place(GridDev G, uint64_t n_im, const double *__restrict__ im_verts, const double *__restrict__ bg_box,
                           const uint64_t *__restrict__ ptr, const uint32_t *__restrict__ tmp, uint32_t *__restrict__ cand_im,
                           uint32_t *__restrict__ cand_bg)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t m = t / LANES;
  const unsigned l = t % LANES;
  if (m >= n_im) return;
  const uint64_t p0 = ptr[m];
  const uint32_t n = (uint32_t)(ptr[m + 1] - p0);
  if (n <= CAP) {
    const uint32_t *row = tmp + m * CAP;
    for (uint32_t k = l; k < n; k += LANES) { cand_bg[p0 + k] = row[k]; cand_im[p0 + k] = (uint32_t)m; }
    return;
  }
  if (l != 0) return;
  constexpr int NVI = 1 << (D - 1);
  double lo[3], hi[3];
  im_box<D, NVI>(im_verts + m * NVI * D, lo, hi);
  uint64_t p = p0;
  query_box<D>(G, bg_box, lo, hi, [&](uint32_t b) {
    uint64_t j = p;
    while (j > p0 && cand_bg[j - 1] > b) { cand_bg[j] = cand_bg[j - 1]; --j; }
    cand_bg[j] = b;
    cand_im[p] = (uint32_t)m;
    ++p;
  });
}


// ---------------------------------------------------------------------------------------
// brick index: the fast path for grid-aligned trees (every reference configuration)
// ---------------------------------------------------------------------------------------
// A background cell of a hyper_cube refinement occupies exactly one cell of the uniform grid of its
// own level, and its corners are c(i) = org + i * g_k bit for bit (one origin for all levels:
// nmbrick::aligned_origin).  When that holds for every cell (verified below, cell by cell, on the raw
// doubles) the grid cells are grouped into bricks of 64 (4x4x4, 8x8 in 2D): one 32-byte hash slot per
// occupied brick carries a 64-bit occupancy mask and the offset of the brick's cell ids, which are
// stored in bit order.  A query then costs one slot per brick (<= 8 per level instead of 27..64 per-cell
// slots), the occupied slots of a surface band fit in L2, and the closed-box test needs no box at all:
// it is done once per axis on the reconstructed corners, which ARE the raw doubles, so the candidate
// set is still exactly the reference's (coupling_utilities.cc:53-55).  Any cell that fails the
// verification, two cells in one grid cell, or an overfull table raise `bad`, and the per-cell chain
// index above takes over.
struct BrickSlot {
  unsigned long long key;   // packed (level, bx, by, bz); ~0 = empty
  unsigned long long mask;  // occupied grid cells of the brick
  uint32_t base;            // first id of the brick in BrickDev::ids
  uint32_t pad0;
  unsigned long long pad1;
};

struct BrickHeader {
  unsigned level_mask;        // levels that hold cells
  unsigned bad;               // 1: a cell is not a grid cell, 2: table overfull, 4: two cells in one grid cell
  unsigned long long cursor;  // ids handed out so far
};

struct BrickDev {
  BrickSlot *slots;
  uint32_t slot_mask;   // capacity - 1
  uint32_t *ids;
  BrickHeader *hdr;
  double org[3];
  double g0, inv0;
  int idx_bits;
};

__device__ __forceinline__ double pow2_neg(int k) { return __hiloint2double((1023 - k) << 20, 0); }   // 2^-k, 0 <= k < 1023

__device__ __forceinline__ uint32_t brick_hash(unsigned long long key)
{
  uint32_t h = (uint32_t)key * 0x9E3779B1u ^ ((uint32_t)(key >> 32) * 0x85EBCA77u);
  h ^= h >> 15; h *= 0xC2B2AE3Du; h ^= h >> 13;
  return h;
}

template <int D>
__device__ __forceinline__ unsigned long long brick_key(int level, int bx, int by, int bz)
{
  return pack_key(D, level, bx, by, bz);
}

__global__ void brick_clear(BrickSlot *s, uint32_t cap)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) { s[i].key = ~0ULL; s[i].mask = 0ULL; }
}

// every cell verifies that it IS a cell of its level grid, claims the slot of its brick and sets its bit
template <int D>
#ifndef NM_BRICK_AGG
#define NM_BRICK_AGG 1
#endif
#ifndef NM_BRICK_MINB
#define NM_BRICK_MINB 8   // index build 2.08 -> 2.00 ms
#endif
__global__ void __launch_bounds__(256, NM_BRICK_MINB) brick_insert(BrickDev B, uint64_t n, const double *__restrict__ box, uint32_t *__restrict__ cell_slot,
                                                    uint8_t *__restrict__ cell_bit)
{
  __shared__ unsigned s_levels;
  if (threadIdx.x == 0) s_levels = 0;
  __syncthreads();
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int SH = nmbrick::BrickShape<D>::SH;
  unsigned my_level_bit = 0;
  uint32_t slot = 0xffffffffu;
  unsigned bit = 0;
#if NM_BRICK_AGG
  unsigned long long key = 0;
  bool ins = false;
  int lvl = 0;
#endif
  if (i < n) {
    const double *b = box + i * 8;
    double e = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) e = fmax(e, b[4 + k] - b[k]);
    const int level = nmbrick::level_of_extent(e, B.g0);
    const double g = B.g0 * (double)(1u << level), inv = B.inv0 * pow2_neg(level);
    int idx[3] = {0, 0, 0};
    bool ok = true;
    const int lim = 1 << B.idx_bits;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const double t = rint((b[k] - B.org[k]) * inv);
      ok &= t >= 0.0 && t < (double)(lim - 1);
      const int j = ok ? (int)t : 0;
      ok &= nmbrick::grid_corner(B.org[k], j, g) == b[k] && nmbrick::grid_corner(B.org[k], j + 1, g) == b[4 + k];
      idx[k] = j;
    }
#if NM_BRICK_AGG
    if (!ok) {
      atomicOr(&B.hdr->bad, 1u);
    } else {
      bit = D == 3 ? (unsigned)((idx[0] & 3) | ((idx[1] & 3) << 2) | ((idx[2] & 3) << 4)) : (unsigned)((idx[0] & 7) | ((idx[1] & 7) << 3));
      key = brick_key<D>(level, idx[0] >> SH, idx[1] >> SH, idx[2] >> SH);
      ins = true;
      lvl = level;
    }
  }
  // Consecutive cells of a tree mostly share their brick: the lanes of a warp that carry the same key insert together,
  // one claim and one OR per brick instead of one per cell (64 same-address atomics per full brick otherwise).
  {
    const unsigned lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(0xffffffffu, ins ? key : (0xFFFFFFFFFFFFFF00ULL | lane));   // no key has every level bit set
    const unsigned leader = (unsigned)__ffs(peers) - 1u;
    const unsigned long long mine = ins ? 1ULL << bit : 0ULL;
    const unsigned lo = __reduce_or_sync(peers, (unsigned)mine), hi = __reduce_or_sync(peers, (unsigned)(mine >> 32));
    const unsigned long long together = ((unsigned long long)hi << 32) | lo;
    if (ins && lane == leader) {
      uint32_t s = brick_hash(key) & B.slot_mask;
      int probes = 0;
      while (true) {
        const unsigned long long old = atomicCAS(&B.slots[s].key, ~0ULL, key);
        if (old == ~0ULL || old == key) { slot = s; break; }
        s = (s + 1) & B.slot_mask;
        if (++probes > 256) { atomicOr(&B.hdr->bad, 2u); break; }   // table too full: not a surface band
      }
      if (slot != 0xffffffffu) {
        const unsigned long long prev = atomicOr(&B.slots[slot].mask, together);
        // two cells in one grid cell (overlapping boxes): among the peers, or with a cell inserted earlier
        if ((prev & together) != 0ULL || __popcll(together) != __popc(peers)) atomicOr(&B.hdr->bad, 4u);
      }
    }
    slot = __shfl_sync(peers, slot, leader);
    if (ins && slot != 0xffffffffu) my_level_bit = 1u << lvl;
  }
  if (i < n) {
    cell_slot[i] = slot;
    cell_bit[i] = (uint8_t)bit;
  }
#else
    if (!ok) {
      atomicOr(&B.hdr->bad, 1u);
    } else {
      bit = D == 3 ? (unsigned)((idx[0] & 3) | ((idx[1] & 3) << 2) | ((idx[2] & 3) << 4)) : (unsigned)((idx[0] & 7) | ((idx[1] & 7) << 3));
      const unsigned long long key = brick_key<D>(level, idx[0] >> SH, idx[1] >> SH, idx[2] >> SH);
      uint32_t s = brick_hash(key) & B.slot_mask;
      int probes = 0;
      while (true) {
        const unsigned long long old = atomicCAS(&B.slots[s].key, ~0ULL, key);
        if (old == ~0ULL || old == key) { slot = s; break; }
        s = (s + 1) & B.slot_mask;
        if (++probes > 256) { atomicOr(&B.hdr->bad, 2u); break; }   // table too full: not a surface band
      }
      if (slot != 0xffffffffu) {
        const unsigned long long prev = atomicOr(&B.slots[slot].mask, 1ULL << bit);
        if ((prev >> bit) & 1ULL) atomicOr(&B.hdr->bad, 4u);   // two cells in one grid cell: overlapping boxes
        my_level_bit = 1u << level;
      }
    }
    cell_slot[i] = slot;
    cell_bit[i] = (uint8_t)bit;
  }
#endif
  my_level_bit = __reduce_or_sync(0xffffffffu, my_level_bit);
  if ((threadIdx.x & 31) == 0 && my_level_bit) atomicOr(&s_levels, my_level_bit);
  __syncthreads();
  if (threadIdx.x == 0 && s_levels) atomicOr(&B.hdr->level_mask, s_levels);
}

__global__ void brick_base(BrickDev B, uint32_t cap)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long m = 0;
  if (i < cap) m = B.slots[i].mask;   // empty slots carry an empty mask
  const unsigned c = (unsigned)__popcll(m);
  // one atomic per warp
  unsigned inc = c;
  const unsigned lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (unsigned)o) inc += v; }
  const unsigned total = __shfl_sync(0xffffffffu, inc, 31);
  unsigned long long base = 0;
  if (lane == 31 && total) base = atomicAdd(&B.hdr->cursor, (unsigned long long)total);
  base = __shfl_sync(0xffffffffu, base, 31);
  if (c) B.slots[i].base = (uint32_t)(base + inc - c);
}

__global__ void brick_fill(BrickDev B, uint64_t n, const uint32_t *__restrict__ cell_slot, const uint8_t *__restrict__ cell_bit)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = cell_slot[i];
  if (s == 0xffffffffu) return;
  const unsigned bit = cell_bit[i];
  const unsigned long long mask = B.slots[s].mask;
  B.ids[B.slots[s].base + __popcll(mask & ((1ULL << bit) - 1ULL))] = (uint32_t)i;
}

// the exact range of grid indices of `level` whose closed cell meets [qlo, qhi] along `axis`
__device__ __forceinline__ bool brick_axis_range(const BrickDev &B, int level, int axis, double qlo, double qhi, int &L, int &H)
{
  const double g = B.g0 * (double)(1u << level), inv = B.inv0 * pow2_neg(level);
  const double org = axis == 0 ? B.org[0] : (axis == 1 ? B.org[1] : B.org[2]);
  return nmbrick::axis_range(org, g, inv, (1 << B.idx_bits) - 2, qlo, qhi, L, H);
}

__device__ __forceinline__ bool brick_find(const BrickDev &B, unsigned long long key, unsigned long long &mask, uint32_t &base)
{
  uint32_t s = brick_hash(key) & B.slot_mask;
  while (true) {
    const ulonglong2 km = *reinterpret_cast<const ulonglong2 *>(&B.slots[s]);
    if (km.x == key) { mask = km.y; base = B.slots[s].base; return true; }
    if (km.x == ~0ULL) return false;
    s = (s + 1) & B.slot_mask;
  }
}

// G lanes per immersed cell.  Lane (s * D + a) computes the index range of axis a for the s-th level of a pass (two
// levels per pass), the ranges go round by shuffle, then the bricks of both levels are probed in parallel lanes.  The
// hits are collected in shared memory, ordered by rank counting and stored as one row of at most CAP ids (the true
// count is recorded even when it exceeds CAP; cand_place_brick probes such cells again).
template <int D, int G, int CAP>
__global__ void __launch_bounds__(128) cand_probe_brick(BrickDev B, uint64_t n_im, const double *__restrict__ im_verts,
                                                        uint32_t *__restrict__ cnt, uint32_t *__restrict__ tmp)
{
  constexpr int NVI = 1 << (D - 1), CPB = 128 / G, SH = nmbrick::BrickShape<D>::SH;
  constexpr int NSL = G >= 2 * D ? 2 : 1;   // levels handled per pass: one lane per (level slot, axis) when the group has D lanes or more
  constexpr bool SPREAD = G >= D;           // fewer lanes than axes: every lane works out all axes itself
  static_assert(CAP % 4 == 0, "rows are read four ids at a time");
  __shared__ __align__(16) uint32_t s_ids[CPB][CAP];
  __shared__ unsigned s_n[CPB];
  const unsigned g = threadIdx.x / G, l = threadIdx.x % G;
  const unsigned gbase = (threadIdx.x & 31) - l;
  const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << gbase);
  const uint64_t m = (uint64_t)blockIdx.x * CPB + g;
  if (m >= n_im) return;   // whole groups leave together
  if (l == 0) s_n[g] = 0;
  for (unsigned k = l; k < (unsigned)CAP; k += G) s_ids[g][k] = 0xffffffffu;   // padding compares as "not smaller"
  // the bounding interval of the immersed cell along "my" axis
  const int ax = (int)(l % D), sl = (int)(l / D);
  double qlo, qhi, blo[3] = {0, 0, 0}, bhi[3] = {0, 0, 0};
  {
    const double *v = im_verts + m * NVI * D + ax;
    qlo = qhi = v[0];
#pragma unroll
    for (int q = 1; q < NVI; ++q) { const double x = v[q * D]; qlo = fmin(qlo, x); qhi = fmax(qhi, x); }
  }
  if (!SPREAD) im_box<D, NVI>(im_verts + m * NVI * D, blo, bhi);
  uint32_t lm = B.hdr->level_mask;
  __syncwarp(gmask);
  while (lm) {
    int lev[2] = {-1, -1};
#pragma unroll
    for (int s = 0; s < NSL; ++s) { lev[s] = lm ? __ffs(lm) - 1 : -1; lm &= lm - 1; }
    const int my_level = sl == 0 ? lev[0] : (sl == 1 ? lev[1] : -1);
    int myL = 1, myH = 0;   // empty
    if (SPREAD && my_level >= 0 && !brick_axis_range(B, my_level, ax, qlo, qhi, myL, myH)) { myL = 1; myH = 0; }
    int L[2][3], H[2][3], b0[2][3], nb[2][3], T[2] = {0, 0};
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int a = 0; a < 3; ++a) { L[s][a] = 1; H[s][a] = 0; b0[s][a] = 0; nb[s][a] = 1; }
#pragma unroll
    for (int s = 0; s < NSL; ++s) {
      bool ok = lev[s] >= 0;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (a < D && SPREAD) {
          L[s][a] = __shfl_sync(gmask, myL, gbase + s * D + a);
          H[s][a] = __shfl_sync(gmask, myH, gbase + s * D + a);
        } else if (a < D) {
          if (!(lev[s] >= 0 && brick_axis_range(B, lev[s], a, blo[a], bhi[a], L[s][a], H[s][a]))) { L[s][a] = 1; H[s][a] = 0; }
        } else { L[s][a] = 0; H[s][a] = 0; }
        ok &= L[s][a] <= H[s][a];
        b0[s][a] = L[s][a] >> SH;
        nb[s][a] = (H[s][a] >> SH) - b0[s][a] + 1;
      }
      T[s] = ok ? nb[s][0] * nb[s][1] * nb[s][2] : 0;
    }
    const int Ttot = T[0] + T[1];
    for (int t = (int)l; t < Ttot; t += G) {
      const int s = t < T[0] ? 0 : 1;
      const int tt = s ? t - T[0] : t;
      const int n0 = s ? nb[1][0] : nb[0][0], n1 = s ? nb[1][1] : nb[0][1];
      int tx, ty, tz;
      if (n0 <= 2 && n1 <= 2) {   // the usual case (a query spans at most two bricks per axis): no integer division
        tx = n0 == 2 ? (tt & 1) : 0;
        const int r = n0 == 2 ? (tt >> 1) : tt;
        ty = n1 == 2 ? (r & 1) : 0;
        tz = n1 == 2 ? (r >> 1) : r;
      } else { tx = tt % n0; ty = (tt / n0) % n1; tz = tt / (n0 * n1); }
      const int bx = (s ? b0[1][0] : b0[0][0]) + tx, by = (s ? b0[1][1] : b0[0][1]) + ty, bz = (s ? b0[1][2] : b0[0][2]) + tz;
      unsigned long long mask;
      uint32_t base;
      if (!brick_find(B, brick_key<D>(s ? lev[1] : lev[0], bx, by, bz), mask, base)) continue;
      unsigned long long hits = mask & nmbrick::brick_query_mask<D>(s ? L[1] : L[0], s ? H[1] : H[0], bx, by, bz);
      const unsigned nh = (unsigned)__popcll(hits);
      if (!nh) continue;
      unsigned pos = atomicAdd(&s_n[g], nh);
      while (hits) {
        const int bit = __ffsll((long long)hits) - 1;
        hits &= hits - 1;
        if (pos < (unsigned)CAP) s_ids[g][pos] = B.ids[base + __popcll(mask & ((1ULL << bit) - 1ULL))];
        ++pos;
      }
    }
  }
  __syncwarp(gmask);
  const unsigned total = s_n[g];
  if (l == 0) cnt[m] = total;
  if (total <= (unsigned)CAP) {
    uint32_t *row = tmp + m * CAP;
    const uint4 *ids4 = reinterpret_cast<const uint4 *>(s_ids[g]);
    for (unsigned e = l; e < total; e += G) {
      const uint32_t v = s_ids[g][e];
      unsigned r = 0;
      for (unsigned j = 0; j < (total + 3) / 4; ++j) {
        const uint4 q = ids4[j];
        r += (q.x < v) + (q.y < v) + (q.z < v) + (q.w < v);
      }
      row[r] = v;
    }
  }
}

// all candidates of one immersed cell, one thread (cells whose row overflowed CAP)
template <int D, typename Emit>
__device__ inline void brick_query_serial(const BrickDev &B, const double *lo, const double *hi, Emit emit)
{
  constexpr int SH = nmbrick::BrickShape<D>::SH;
  uint32_t lm = B.hdr->level_mask;
  while (lm) {
    const int level = __ffs(lm) - 1;
    lm &= lm - 1;
    int L[3] = {0, 0, 0}, H[3] = {0, 0, 0};
    bool ok = true;
    for (int k = 0; k < D; ++k) ok &= brick_axis_range(B, level, k, lo[k], hi[k], L[k], H[k]);
    if (!ok) continue;
    for (int bz = L[2] >> SH; bz <= (H[2] >> SH); ++bz)
      for (int by = L[1] >> SH; by <= (H[1] >> SH); ++by)
        for (int bx = L[0] >> SH; bx <= (H[0] >> SH); ++bx) {
          unsigned long long mask;
          uint32_t base;
          if (!brick_find(B, brick_key<D>(level, bx, by, bz), mask, base)) continue;
          unsigned long long hits = mask & nmbrick::brick_query_mask<D>(L, H, bx, by, bz);
          while (hits) {
            const int bit = __ffsll((long long)hits) - 1;
            hits &= hits - 1;
            emit(B.ids[base + __popcll(mask & ((1ULL << bit) - 1ULL))]);
          }
        }
  }
}

// (bounded to 32 registers: the copy is the common case and waits on memory -- it wants every warp the SM can hold; the
// serial re-query of an overflowed row may spill.  Unbounded: 57 registers, 4 blocks per SM.)
#ifndef NM_PLACE_MINB
#define NM_PLACE_MINB 8
#endif
template <int D, int CAP, int LANES>
__global__ void __launch_bounds__(256, NM_PLACE_MINB) cand_place_brick(BrickDev B, uint64_t n_im, const double *__restrict__ im_verts, const uint64_t *__restrict__ ptr,
                                 const uint32_t *__restrict__ tmp, uint32_t *__restrict__ cand_im, uint32_t *__restrict__ cand_bg)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t m = t / LANES;
  const unsigned l = t % LANES;
  if (m >= n_im) return;
  const uint64_t p0 = ptr[m];
  const uint32_t n = (uint32_t)(ptr[m + 1] - p0);
  if (n <= CAP) {
    const uint32_t *row = tmp + m * CAP;
    for (uint32_t k = l; k < n; k += LANES) { cand_bg[p0 + k] = row[k]; cand_im[p0 + k] = (uint32_t)m; }
    return;
  }
  if (l != 0) return;
  constexpr int NVI = 1 << (D - 1);
  double lo[3], hi[3];
  im_box<D, NVI>(im_verts + m * NVI * D, lo, hi);
  uint64_t p = p0;
  brick_query_serial<D>(B, lo, hi, [&](uint32_t b) {
    uint64_t j = p;
    while (j > p0 && cand_bg[j - 1] > b) { cand_bg[j] = cand_bg[j - 1]; --j; }
    cand_bg[j] = b;
    cand_im[p] = (uint32_t)m;
    ++p;
  });
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
int nm_exclusive_scan_u64(nm_ctx *ctx, const uint64_t *in, uint64_t *out, uint64_t n)
{
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, ctx->stream);
  NM_TRY(nm_ensure(ctx, ctx->scan_tmp, bytes));
  NM_CUDA(cub::DeviceScan::ExclusiveSum(ctx->scan_tmp.p, bytes, in, out, n, ctx->stream));
  return NM_OK;
}

namespace {
struct U32ToU64 {
  __host__ __device__ uint64_t operator()(uint32_t v) const { return (uint64_t)v; }
};
}  // namespace

namespace {
struct BothWidths {
  __host__ __device__ thrust::tuple<uint64_t, uint32_t> operator()(uint64_t v) const { return thrust::make_tuple(v, (uint32_t)v); }
};
}  // namespace

// the same scan, every prefix written twice: as 64 bits and truncated to 32 (the caller uses the narrow copy only when the
// total turns out to fit) -- saves a pass over the offsets
int nm_exclusive_scan_u32_to_u64_and_u32(nm_ctx *ctx, const uint32_t *in, uint64_t *out, uint32_t *out32, uint64_t n)
{
  thrust::transform_iterator<U32ToU64, const uint32_t *, uint64_t> it(in, U32ToU64());
  auto both = thrust::make_transform_output_iterator(thrust::make_zip_iterator(thrust::make_tuple(out, out32)), BothWidths());
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, both, n, ctx->stream);
  NM_TRY(nm_ensure(ctx, ctx->scan_tmp, bytes));
  NM_CUDA(cub::DeviceScan::ExclusiveSum(ctx->scan_tmp.p, bytes, it, both, n, ctx->stream));
  return NM_OK;
}

int nm_exclusive_scan_u32_to_u64(nm_ctx *ctx, const uint32_t *in, uint64_t *out, uint64_t n)
{
  thrust::transform_iterator<U32ToU64, const uint32_t *, uint64_t> it(in, U32ToU64());
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, out, n, ctx->stream);
  NM_TRY(nm_ensure(ctx, ctx->scan_tmp, bytes));
  NM_CUDA(cub::DeviceScan::ExclusiveSum(ctx->scan_tmp.p, bytes, it, out, n, ctx->stream));
  return NM_OK;
}

int nm_bg_layout(nm_ctx *ctx, int d, uint64_t n, const double *verts_dev, const double *lo_dev, const double *hi_dev, const double *edge_dev)
{
  NM_TRY(nm_ensure(ctx, ctx->bg_box, n * 8 * sizeof(double)));
  NM_TRY(nm_ensure(ctx, ctx->counters, 64 * sizeof(uint64_t)));
  int *bad = ctx->counters.as<int>();
  NM_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream));
  const unsigned T = 256, B = nm_blocks(n, T);
  ctx->bg_stats_ready = false;
  if (n) {
    if (verts_dev) {
      if (d == 2) bg_boxes_from_verts<2><<<B, T, 0, NM_LS(ctx)>>>(n, verts_dev, ctx->bg_box.as<double>(), bad);
      else bg_boxes_from_verts<3><<<B, T, 0, NM_LS(ctx)>>>(n, verts_dev, ctx->bg_box.as<double>(), bad);
    } else {
      NM_TRY(nm_ensure(ctx, ctx->idx_tmp, NM_STATS_BLOCKS * NM_STATS_W * sizeof(double)));
      if (d == 2) bg_boxes_from_lohi_stats<2><<<NM_STATS_BLOCKS, 256, 0, NM_LS(ctx)>>>(n, lo_dev, hi_dev, edge_dev, ctx->bg_box.as<double>(), bad, ctx->idx_tmp.as<double>());
      else bg_boxes_from_lohi_stats<3><<<NM_STATS_BLOCKS, 256, 0, NM_LS(ctx)>>>(n, lo_dev, hi_dev, edge_dev, ctx->bg_box.as<double>(), bad, ctx->idx_tmp.as<double>());
      ctx->bg_stats_ready = true;
    }
    NM_CUDA(cudaGetLastError());
  }
  int h_bad = 0;
  NM_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  NM_SYNC(ctx);
  if (h_bad)
    return nm_fail(ctx, NM_ERR_UNSUPPORTED,
                   "%d background cell(s) are not axis-aligned boxes in deal.II vertex order; only Cartesian "
                   "background cells (hyper_cube refinements, as in every reference config) are implemented", h_bad);
  return NM_OK;
}

int nm_constraints_layout(nm_ctx *ctx, const uint32_t *c_dof_dev)
{
  NM_TRY(nm_ensure(ctx, ctx->line_of, (ctx->n_space_dofs + 1) * sizeof(int32_t)));
  NM_CUDA(cudaMemsetAsync(ctx->line_of.p, 0xff, ctx->n_space_dofs * sizeof(int32_t), ctx->stream));
  if (ctx->n_constrained) {
    int *bad = ctx->counters.as<int>();
    NM_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream));
    fill_line_of<<<nm_blocks(ctx->n_constrained, 256), 256, 0, NM_LS(ctx)>>>(
        ctx->n_constrained, c_dof_dev, ctx->line_of.as<int32_t>(), ctx->n_space_dofs, bad);
    NM_CUDA(cudaGetLastError());
    int h_bad = 0;
    NM_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    NM_SYNC(ctx);
    if (h_bad) return nm_fail(ctx, NM_ERR_INVALID, "%d constrained DoF index(es) >= n_dofs", h_bad);
  }
  return NM_OK;
}

// the constrained dof of every constraint line, in line order (device array of n_constrained entries)
int nm_constrained_dofs(nm_ctx *ctx, uint32_t *out_dev)
{
  if (ctx->n_space_dofs)
    constrained_dofs_kernel<<<nm_blocks(ctx->n_space_dofs, 256), 256, 0, NM_LS(ctx)>>>(ctx->n_space_dofs, ctx->line_of.as<int32_t>(), out_dev);
  NM_CUDA(cudaGetLastError());
  return NM_OK;
}

static GridDev make_grid_dev(const nm_ctx *ctx)
{
  GridDev G;
  memset(&G, 0, sizeof(G));
  G.d = ctx->gi.d;
  for (int k = 0; k < 3; ++k) G.org[k] = ctx->gi.org[k];
  double g = ctx->gi.g0;
  G.g0 = g;
  for (int k = 0; k < 32; ++k) { G.inv[k] = 1.0 / g; g *= 2.0; }
  G.level_mask = ctx->gi.level_mask;
  G.hash = ctx->idx_hash.as<HashSlot>();
  G.hash_mask = ctx->gi.hash_cap - 1;
  G.next = ctx->idx_next.as<uint32_t>();
  G.extra_bg = ctx->idx_extra_bg.as<uint32_t>();
  G.n_bg = (uint32_t)ctx->n_bg;
  G.idx_bits = ctx->gi.d == 2 ? 29 : 19;
  return G;
}

static int chain_index_build(nm_ctx *ctx);
static int brick_index_build(nm_ctx *ctx, bool roomy);

int nm_index_build(nm_ctx *ctx)
{
  const int d = ctx->spacedim;
  const uint64_t n = ctx->n_bg;
  ctx->gi = GridIndex();
  ctx->gi.d = d;
  ctx->index_valid = false;
  if (n == 0) { ctx->gi.g0 = 1; ctx->gi.hash_cap = 1; ctx->index_valid = true;
    NM_TRY(nm_ensure(ctx, ctx->idx_hash, sizeof(HashSlot)));
    hash_clear<<<1, 32, 0, NM_LS(ctx)>>>(ctx->idx_hash.as<HashSlot>(), 1);
    return NM_OK; }
  // 1. origin and level-0 spacing
  constexpr unsigned SB = NM_STATS_BLOCKS;
  NM_TRY(nm_ensure(ctx, ctx->idx_tmp, SB * NM_STATS_W * sizeof(double)));
  if (!ctx->bg_stats_ready) {   // the lo/hi layout kernel has already taken them
    if (d == 2) bg_stats<2><<<SB, 256, 0, NM_LS(ctx)>>>(n, ctx->bg_box.as<double>(), ctx->idx_tmp.as<double>());
    else bg_stats<3><<<SB, 256, 0, NM_LS(ctx)>>>(n, ctx->bg_box.as<double>(), ctx->idx_tmp.as<double>());
    NM_CUDA(cudaGetLastError());
  }
  ctx->bg_stats_ready = false;  // idx_tmp is reused below
  static thread_local double h_stats[SB * NM_STATS_W];
  NM_CUDA(cudaMemcpyAsync(h_stats, ctx->idx_tmp.p, sizeof(h_stats), cudaMemcpyDeviceToHost, ctx->stream));
  NM_SYNC(ctx);
  double org[3] = {INFINITY, INFINITY, INFINITY}, top[3] = {-INFINITY, -INFINITY, -INFINITY}, g0 = INFINITY, gmax = 0;
  double anchor[3] = {0, 0, 0};
  for (unsigned b = 0; b < SB; ++b) {
    const double *h = h_stats + b * NM_STATS_W;
    for (int k = 0; k < 3; ++k) { org[k] = fmin(org[k], h[k]); top[k] = fmax(top[k], h[3 + k]); }
    g0 = fmin(g0, h[6]);
    if (h[7] > gmax) { gmax = h[7]; for (int k = 0; k < 3; ++k) anchor[k] = h[8 + k]; }
  }
  if (!(g0 < INFINITY)) g0 = 1.0;  // all cells degenerate
  for (int k = 0; k < 3; ++k) ctx->gi.org[k] = k < d ? org[k] : 0.0;
  ctx->gi.g0 = g0;
  // one origin for the grids of all levels (brick index): a coarsest cell's corner, moved down by whole coarsest cells
  for (int k = 0; k < 3; ++k) ctx->gi.brick_org[k] = (k < d && gmax > 0) ? nmbrick::aligned_origin(org[k], anchor[k], gmax) : 0.0;
  if (gmax > g0 * 2147483648.0)
    return nm_fail(ctx, NM_ERR_UNSUPPORTED, "background cell extents span more than 31 octaves (%g .. %g)", g0, gmax);
  const int bits = d == 2 ? 29 : 19;
  for (int k = 0; k < d; ++k)
    if ((top[k] - org[k]) / g0 >= (double)(1LL << bits) - 2)
      return nm_fail(ctx, NM_ERR_UNSUPPORTED,
                     "background spans more than 2^%d finest cells along axis %d (extent %g, finest cell %g)", bits, k,
                     top[k] - org[k], g0);

  if (!ctx->force_chain_index) {
    NM_TRY(brick_index_build(ctx, false));
    ctx->index_valid = true;
    return NM_OK;
  }
  return chain_index_build(ctx);
}

// per-cell chain index: any boxes (a cell may cover several grid cells of its level)
static int chain_index_build(nm_ctx *ctx)
{
  const int d = ctx->spacedim;
  const uint64_t n = ctx->n_bg;
  // 2. size of the table: number of (cell, grid cell) registrations
  GridDev G = make_grid_dev(ctx);
  NM_TRY(nm_ensure(ctx, ctx->counters, 64 * sizeof(uint64_t)));
  int *bad = ctx->counters.as<int>();
  unsigned *lmask = ctx->counters.as<unsigned>() + 2;
  unsigned *extra_cursor = ctx->counters.as<unsigned>() + 3;
  unsigned long long *total = ctx->counters.as<unsigned long long>() + 2;
  NM_CUDA(cudaMemsetAsync(ctx->counters.p, 0, 64, ctx->stream));
  const unsigned T = 256, B = nm_blocks(n, T);
  const unsigned GB = (unsigned)(B < (unsigned)ctx->sm_count * 8u ? B : (unsigned)ctx->sm_count * 8u);
  if (d == 2) entries_count<2><<<GB, T, 0, NM_LS(ctx)>>>(G, n, ctx->bg_box.as<double>(), total, lmask, bad);
  else entries_count<3><<<GB, T, 0, NM_LS(ctx)>>>(G, n, ctx->bg_box.as<double>(), total, lmask, bad);
  NM_CUDA(cudaGetLastError());
  unsigned long long h_hdr[3];
  NM_CUDA(cudaMemcpyAsync(h_hdr, ctx->counters.p, sizeof(h_hdr), cudaMemcpyDeviceToHost, ctx->stream));
  NM_SYNC(ctx);
  if ((unsigned)(h_hdr[0] & 0xffffffffu)) return nm_fail(ctx, NM_ERR_UNSUPPORTED, "background cells outside the indexable grid range");
  const uint64_t ne = h_hdr[2];
  if (ne >= 0x7fffffffULL) return nm_fail(ctx, NM_ERR_UNSUPPORTED, "grid index needs %llu entries (limit 2^31)", (unsigned long long)ne);
  ctx->gi.level_mask = (unsigned)(h_hdr[1] & 0xffffffffu);
  ctx->gi.n_entries = (uint32_t)ne;
  // 3. hash: occupied grid cell -> chain of background cells
  uint32_t cap = 64;
  while (cap < 2 * ne) cap <<= 1;
  ctx->gi.hash_cap = cap;
  NM_TRY(nm_ensure(ctx, ctx->idx_hash, (size_t)cap * sizeof(HashSlot)));
  NM_TRY(nm_ensure(ctx, ctx->idx_next, (ne + 1) * 4));
  NM_TRY(nm_ensure(ctx, ctx->idx_extra_bg, (ne - n + 1) * 4));
  G = make_grid_dev(ctx);
  hash_clear<<<nm_blocks(cap, 256), 256, 0, NM_LS(ctx)>>>(ctx->idx_hash.as<HashSlot>(), cap);
  if (d == 2) index_insert<2><<<B, T, 0, NM_LS(ctx)>>>(G, n, ctx->bg_box.as<double>(), ctx->idx_hash.as<HashSlot>(), ctx->idx_next.as<uint32_t>(),
                                                      ctx->idx_extra_bg.as<uint32_t>(), extra_cursor);
  else index_insert<3><<<B, T, 0, NM_LS(ctx)>>>(G, n, ctx->bg_box.as<double>(), ctx->idx_hash.as<HashSlot>(), ctx->idx_next.as<uint32_t>(),
                                                ctx->idx_extra_bg.as<uint32_t>(), extra_cursor);
  NM_CUDA(cudaGetLastError());
  ctx->index_kind = 1;
  ctx->index_valid = true;
  return NM_OK;
}

static BrickDev make_brick_dev(const nm_ctx *ctx)
{
  BrickDev B;
  B.slots = ctx->idx_hash.as<BrickSlot>();
  B.slot_mask = ctx->gi.hash_cap - 1;
  B.ids = ctx->idx_extra_bg.as<uint32_t>();
  B.hdr = ctx->brk_hdr.as<BrickHeader>();
  for (int k = 0; k < 3; ++k) B.org[k] = ctx->gi.brick_org[k];
  B.g0 = ctx->gi.g0;
  B.inv0 = 1.0 / ctx->gi.g0;
  B.idx_bits = ctx->gi.d == 2 ? 29 : 19;
  return B;
}

// brick index (grid-aligned trees): verification + insertion, id offsets, ids.  No host synchronisation: a failed
// verification is noticed with the candidate count (nm_candidates_run) and the chain index takes over.
// `roomy`: a table that cannot overflow (one slot per cell and more); the default holds a surface band comfortably
// (~25 cells per occupied brick) and is cleared and scanned four times faster.
static int brick_index_build(nm_ctx *ctx, bool roomy)
{
  const int d = ctx->spacedim;
  const uint64_t n = ctx->n_bg;
  const uint64_t want = roomy ? 2 * n + 1 : n / 8 + 1;
  uint32_t cap = 1024;
  while (cap < want) cap <<= 1;
  ctx->gi.hash_cap = cap;
  ctx->gi.n_entries = (uint32_t)n;
  ctx->brick_roomy = roomy;
  NM_TRY(nm_ensure(ctx, ctx->idx_hash, (size_t)cap * sizeof(BrickSlot)));
  NM_TRY(nm_ensure(ctx, ctx->idx_next, (n + 1) * 4));       // slot of every cell
  NM_TRY(nm_ensure(ctx, ctx->brk_bit, n + 1));              // bit of every cell inside its brick
  NM_TRY(nm_ensure(ctx, ctx->idx_extra_bg, (n + 1) * 4));   // cell ids, brick by brick in bit order
  NM_TRY(nm_ensure(ctx, ctx->brk_hdr, sizeof(BrickHeader)));
  BrickHeader *hdr = ctx->brk_hdr.as<BrickHeader>();
  NM_CUDA(cudaMemsetAsync(hdr, 0, sizeof(BrickHeader), ctx->stream));
  const BrickDev B = make_brick_dev(ctx);
  const unsigned T = 256, NB = nm_blocks(n, T);
  const double *box = ctx->bg_box.as<double>();
  brick_clear<<<nm_blocks(cap, 256), 256, 0, NM_LS(ctx)>>>(B.slots, cap);
  if (d == 2) brick_insert<2><<<NB, T, 0, NM_LS(ctx)>>>(B, n, box, ctx->idx_next.as<uint32_t>(), ctx->brk_bit.as<uint8_t>());
  else brick_insert<3><<<NB, T, 0, NM_LS(ctx)>>>(B, n, box, ctx->idx_next.as<uint32_t>(), ctx->brk_bit.as<uint8_t>());
  brick_base<<<nm_blocks(cap, 256), 256, 0, NM_LS(ctx)>>>(B, cap);
  brick_fill<<<NB, T, 0, NM_LS(ctx)>>>(B, n, ctx->idx_next.as<uint32_t>(), ctx->brk_bit.as<uint8_t>());
  NM_CUDA(cudaGetLastError());
  ctx->index_kind = 2;
  return NM_OK;
}

template <int D, int CAP>
static int candidates_run_chains(nm_ctx *ctx);

// candidates through the brick index; *fell_back is set when the verification of the index failed
template <int D, int CAP, int G>
static int candidates_run_bricks(nm_ctx *ctx, bool *fell_back)
{
  const uint64_t n_im = ctx->n_im;
  const BrickDev B = make_brick_dev(ctx);
  *fell_back = false;
  NM_TRY(nm_ensure(ctx, ctx->scratch_a, (n_im + 1) * sizeof(uint32_t)));
  NM_TRY(nm_ensure(ctx, ctx->cand_ptr, (n_im + 1) * sizeof(uint64_t)));
  NM_TRY(nm_ensure(ctx, ctx->cand_tmp, (n_im + 1) * CAP * sizeof(uint32_t)));
  ctx->counts.n_candidates = 0;
  NM_CUDA(cudaMemsetAsync(ctx->scratch_a.as<uint32_t>() + n_im, 0, 4, ctx->stream));
  constexpr int CPB = 128 / G;
  KernelTimer kc(ctx, NM_T_K_CAND_COUNT);
  if (n_im)
    cand_probe_brick<D, G, CAP><<<nm_blocks(n_im, CPB), 128, 0, NM_LS(ctx)>>>(B, n_im, ctx->im_verts.as<double>(), ctx->scratch_a.as<uint32_t>(),
                                                                             ctx->cand_tmp.as<uint32_t>());
  kc.stop();
  NM_CUDA(cudaGetLastError());
  NM_TRY(nm_exclusive_scan_u32_to_u64(ctx, ctx->scratch_a.as<uint32_t>(), ctx->cand_ptr.as<uint64_t>(), n_im + 1));
  uint64_t n_cand = 0;
  unsigned bad = 0;
  NM_CUDA(cudaMemcpyAsync(&n_cand, ctx->cand_ptr.as<uint64_t>() + n_im, 8, cudaMemcpyDeviceToHost, ctx->stream));
  NM_CUDA(cudaMemcpyAsync(&bad, &B.hdr->bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
  NM_SYNC(ctx);
  if (bad == 2 && !ctx->brick_roomy) {   // a tree, but denser in bricks than a surface band: same index, roomy table
    NM_TRY(brick_index_build(ctx, true));
    return candidates_run_bricks<D, CAP, G>(ctx, fell_back);
  }
  if (bad) {   // not a grid-aligned tree: redo with the general index
    NM_TRY(chain_index_build(ctx));
    *fell_back = true;
    return NM_OK;
  }
  if (n_cand >= 0xffffffffULL * 16) return nm_fail(ctx, NM_ERR_UNSUPPORTED, "too many candidate pairs (%llu)", (unsigned long long)n_cand);
  ctx->counts.n_candidates = n_cand;
  NM_TRY(nm_ensure(ctx, ctx->cand_im, (n_cand + 1) * 4));
  NM_TRY(nm_ensure(ctx, ctx->cand_bg, (n_cand + 1) * 4));
  constexpr int LANES = 8;
  KernelTimer kf(ctx, NM_T_K_CAND_FILL);
  if (n_im)
    cand_place_brick<D, CAP, LANES><<<nm_blocks(n_im * LANES, 256), 256, 0, NM_LS(ctx)>>>(B, n_im, ctx->im_verts.as<double>(), ctx->cand_ptr.as<uint64_t>(),
                                                                                        ctx->cand_tmp.as<uint32_t>(), ctx->cand_im.as<uint32_t>(),
                                                                                        ctx->cand_bg.as<uint32_t>());
  kf.stop();
  NM_CUDA(cudaGetLastError());
  return NM_OK;
}

template <int D, int CAP>
static int candidates_run_chains(nm_ctx *ctx)
{
  const uint64_t n_im = ctx->n_im;
  GridDev G = make_grid_dev(ctx);
  NM_TRY(nm_ensure(ctx, ctx->scratch_a, (n_im + 1) * sizeof(uint32_t)));
  NM_TRY(nm_ensure(ctx, ctx->cand_ptr, (n_im + 1) * sizeof(uint64_t)));
  NM_TRY(nm_ensure(ctx, ctx->cand_tmp, (n_im + 1) * CAP * sizeof(uint32_t)));
  ctx->counts.n_candidates = 0;
  if (n_im == 0) { NM_CUDA(cudaMemsetAsync(ctx->cand_ptr.p, 0, 8, ctx->stream)); return NM_OK; }
  const unsigned T = 128, B = nm_blocks(n_im, T);
  NM_CUDA(cudaMemsetAsync(ctx->scratch_a.as<uint32_t>() + n_im, 0, 4, ctx->stream));
  KernelTimer kc(ctx, NM_T_K_CAND_COUNT);
  cand_probe<D, CAP><<<B, T, 0, NM_LS(ctx)>>>(G, n_im, ctx->im_verts.as<double>(), ctx->bg_box.as<double>(), ctx->scratch_a.as<uint32_t>(),
                                             ctx->cand_tmp.as<uint32_t>());
  kc.stop();
  NM_CUDA(cudaGetLastError());
  NM_TRY(nm_exclusive_scan_u32_to_u64(ctx, ctx->scratch_a.as<uint32_t>(), ctx->cand_ptr.as<uint64_t>(), n_im + 1));
  uint64_t n_cand = 0;
  NM_CUDA(cudaMemcpyAsync(&n_cand, ctx->cand_ptr.as<uint64_t>() + n_im, 8, cudaMemcpyDeviceToHost, ctx->stream));
  NM_SYNC(ctx);
  if (n_cand >= 0xffffffffULL * 16) return nm_fail(ctx, NM_ERR_UNSUPPORTED, "too many candidate pairs (%llu)", (unsigned long long)n_cand);
  ctx->counts.n_candidates = n_cand;
  NM_TRY(nm_ensure(ctx, ctx->cand_im, (n_cand + 1) * 4));
  NM_TRY(nm_ensure(ctx, ctx->cand_bg, (n_cand + 1) * 4));
  constexpr int LANES = 8;
  KernelTimer kf(ctx, NM_T_K_CAND_FILL);
  cand_place<D, CAP, LANES><<<nm_blocks(n_im * LANES, 256), 256, 0, NM_LS(ctx)>>>(G, n_im, ctx->im_verts.as<double>(), ctx->bg_box.as<double>(),
                                                                                ctx->cand_ptr.as<uint64_t>(), ctx->cand_tmp.as<uint32_t>(),
                                                                                ctx->cand_im.as<uint32_t>(), ctx->cand_bg.as<uint32_t>());
  kf.stop();
  NM_CUDA(cudaGetLastError());
  return NM_OK;
}

int nm_candidates_run(nm_ctx *ctx)
{
  if (ctx->n_im == 0) { NM_TRY(nm_ensure(ctx, ctx->cand_ptr, 16)); ctx->counts.n_candidates = 0; NM_CUDA(cudaMemsetAsync(ctx->cand_ptr.p, 0, 8, ctx->stream)); return NM_OK; }
  if (ctx->index_kind == 2 && ctx->n_bg) {
    bool fell_back = false;
#ifndef NM_CAND_G3
#define NM_CAND_G3 2    // lanes per immersed quad in cand_probe_brick (3D); measured at sphere cycle 8: 8 lanes 2.57 ms, 4: 1.71, 2: 1.70, 1: 1.89
#endif
    NM_TRY((ctx->spacedim == 2 ? candidates_run_bricks<2, 16, 4>(ctx, &fell_back) : candidates_run_bricks<3, 40, NM_CAND_G3>(ctx, &fell_back)));
    if (!fell_back) return NM_OK;
  }
  return ctx->spacedim == 2 ? candidates_run_chains<2, 16>(ctx) : candidates_run_chains<3, 40>(ctx);
}
