// Subsystems (3b) and (4): scatter through constraint lines into a device-built CSR matrix,
// and the C / C^T sparse matrix-vector products.
//
// Replaces
//   create_coupling_sparsity_pattern_with_exact_intersections   code/include/coupling_utilities.h:100-139
//     (AffineConstraints::add_entries_local_to_global with keep_constrained_entries = true)
//   the scatter of create_coupling_mass_matrix_with_exact_intersections   :276-289
//     (AffineConstraints::distribute_local_to_global, rectangular version, + compress(add))
//   Ct.vmult / C.Tvmult on the Trilinos matrix   code/examples/lagrange_multiplier/2d3d/lagrange_multiplier.cc:627-643
//
// Design: the immersed space is discontinuous (FE_DGQ), so an immersed cell owns its column of
// the stored matrix exclusively.  One warp merges all accepted pairs of one immersed cell:
// every (space dof, value) contribution is expanded through its constraint line and added, in
// the order of a sequential loop over the contributions, into a hash table in shared memory
// (merge_rows_fast; cells it cannot take go through the sort-based general kernels); only the
// distinct dofs are ordered.  No floating-point atomics, no global sort, bit-reproducible
// values.  The rows produced this way are the rows of C (= columns of the stored matrix); the
// stored orientation is obtained by one device transpose (counts from the merge, scan, scatter
// of 16-byte records, per-row sort).
#include <cub/cub.cuh>

#include "nm_common.cuh"

namespace {

constexpr int WARP_CAP = 512;    // entries a warp can merge in shared memory
constexpr int BLOCK_CAP = 8192;  // entries a 256-thread block can merge
constexpr int HASH_CAP = 192;    // contributions the hash-table fast path accepts per immersed cell
constexpr int HASH_SLOTS = 256;
constexpr unsigned HASH_EMPTY = 0xffffffffu;

// ---------------------------------------------------------------------------------------
// capacity of each row of C: accepted pairs x (dofs + masters)
// ---------------------------------------------------------------------------------------
template <int NL>
__global__ void row_capacity_kernel(uint64_t n_im, const uint64_t *__restrict__ cand_ptr, const uint32_t *__restrict__ cand_bg,
                                    const uint8_t *__restrict__ acc, const uint32_t *__restrict__ bg_dofs,
                                    const int32_t *__restrict__ line_of, const uint64_t *__restrict__ c_ptr,
                                    uint32_t *__restrict__ cap, unsigned *max_cap, unsigned *n_big, uint32_t n_dofs, unsigned *bad_dof)
{
  const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_im) return;
  uint32_t c = 0;
  for (uint64_t p = cand_ptr[m]; p < cand_ptr[m + 1]; ++p) {
    if (!acc[p]) continue;
    const uint32_t *dofs = bg_dofs + (uint64_t)cand_bg[p] * NL;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      if (dofs[i] >= n_dofs) { atomicOr(bad_dof, 1u); continue; }   // the caller's table points outside the declared dofs
      const int32_t ln = line_of[dofs[i]];
      c += 1 + (ln >= 0 ? (uint32_t)(c_ptr[ln + 1] - c_ptr[ln]) : 0u);
    }
  }
  cap[m] = c;
  if (c > WARP_CAP) { atomicAdd(n_big, 1u); atomicMax(max_cap, c); }
  if (c > HASH_CAP) atomicAdd(n_big + 2, 1u);
}

__global__ void list_big_rows(uint64_t n_im, const uint32_t *__restrict__ cap, uint32_t *__restrict__ list, unsigned *cursor)
{
  const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m < n_im && cap[m] > WARP_CAP) list[atomicAdd(cursor, 1u)] = (uint32_t)m;
}

// ---------------------------------------------------------------------------------------
// merge: one group (warp or block) per immersed cell
// ---------------------------------------------------------------------------------------
template <int G> __device__ inline void group_sync() { if (G == 32) __syncwarp(); else __syncthreads(); }

// exclusive scan of one value per thread inside the group; returns the offset, `total` to all
template <int G>
__device__ inline unsigned group_scan(unsigned v, unsigned &total, unsigned *s_tmp)
{
  const unsigned lane = threadIdx.x & 31;
  unsigned inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (unsigned)o) inc += t;
  }
  if (G == 32) {
    total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - v;
  }
  const unsigned w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 31) s_tmp[w] = inc;
  __syncthreads();
  unsigned base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < G / 32; ++i) { unsigned t = s_tmp[i]; if ((unsigned)i < w) base += t; tot += t; }
  total = tot;
  return base + inc - v;
}

template <int NL, int G, int CAP>
__device__ inline void merge_cell(uint64_t m, unsigned tid, unsigned long long *s_key, double *s_val, unsigned *s_tmp,
                                  const uint64_t *__restrict__ cand_ptr, const uint32_t *__restrict__ cand_bg,
                                  const uint8_t *__restrict__ acc, const double *__restrict__ local,
                                  const uint32_t *__restrict__ bg_dofs, const int32_t *__restrict__ line_of,
                                  const uint64_t *__restrict__ c_ptr, const uint32_t *__restrict__ c_master,
                                  const double *__restrict__ c_weight, const uint64_t *__restrict__ rowstart,
                                  uint32_t *__restrict__ rowlen, uint32_t *__restrict__ out_col, double *__restrict__ out_val)
{
  const uint64_t p0 = cand_ptr[m], p1 = cand_ptr[m + 1];
  // 1. emit (dof, value) entries, constraint lines expanded, in a fixed order
  unsigned E = 0;
  for (uint64_t base = p0; base < p1; base += G) {
    const uint64_t p = base + tid;
    const bool on = p < p1 && acc[p];
    unsigned cnt = 0;
    uint32_t dofs[NL];
    int32_t ln[NL];
    if (on) {
      const uint32_t *dp = bg_dofs + (uint64_t)cand_bg[p] * NL;
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        dofs[i] = dp[i];
        ln[i] = line_of[dofs[i]];
        cnt += 1 + (ln[i] >= 0 ? (unsigned)(c_ptr[ln[i] + 1] - c_ptr[ln[i]]) : 0u);
      }
    }
    unsigned total;
    unsigned off = E + group_scan<G>(cnt, total, s_tmp);
    if (on) {
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        const double v = local[p * NL + i];
        if (ln[i] < 0) {
          s_key[off] = ((unsigned long long)dofs[i] << 32) | off; s_val[off] = v; ++off;
        } else {
          // keep_constrained_entries: the constrained row is in the pattern, structurally zero
          s_key[off] = ((unsigned long long)dofs[i] << 32) | off; s_val[off] = 0.0; ++off;
          for (uint64_t k = c_ptr[ln[i]]; k < c_ptr[ln[i] + 1]; ++k) {
            s_key[off] = ((unsigned long long)c_master[k] << 32) | off; s_val[off] = c_weight[k] * v; ++off;
          }
        }
      }
    }
    E += total;
  }
  if (E == 0) { if (tid == 0) rowlen[m] = 0; return; }
  // 2. bitonic sort by (dof, emission index)
  unsigned n2 = 32;
  while (n2 < E) n2 <<= 1;
  for (unsigned i = E + tid; i < n2; i += G) { s_key[i] = ~0ULL; s_val[i] = 0.0; }
  group_sync<G>();
  for (unsigned k = 2; k <= n2; k <<= 1)
    for (unsigned j = k >> 1; j > 0; j >>= 1) {
      for (unsigned i = tid; i < n2; i += G) {
        const unsigned l = i ^ j;
        if (l > i) {
          const unsigned long long a = s_key[i], b = s_key[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            s_key[i] = b; s_key[l] = a;
            const double t = s_val[i]; s_val[i] = s_val[l]; s_val[l] = t;
          }
        }
      }
      group_sync<G>();
    }
  // 3. heads of equal-dof runs -> one matrix entry each, summed in emission order
  unsigned written = 0;
  const uint64_t dst = rowstart[m];
  for (unsigned base = 0; base < E; base += G) {
    const unsigned i = base + tid;
    bool head = false;
    uint32_t dof = 0;
    if (i < E) {
      dof = (uint32_t)(s_key[i] >> 32);
      head = i == 0 || (uint32_t)(s_key[i - 1] >> 32) != dof;
    }
    unsigned total;
    const unsigned rank = written + group_scan<G>(head ? 1u : 0u, total, s_tmp);
    if (head) {
      double s = s_val[i];
      for (unsigned j = i + 1; j < E && (uint32_t)(s_key[j] >> 32) == dof; ++j) s += s_val[j];
      out_col[dst + rank] = dof;
      out_val[dst + rank] = s;
    }
    written += total;
  }
  if (tid == 0) rowlen[m] = written;
}

// ---------------------------------------------------------------------------------------
// fast path of the merge: per-warp hash table in shared memory instead of a sort
// ---------------------------------------------------------------------------------------
// A cell with at most HASH_CAP contributions has at most HASH_CAP distinct space dofs: they are
// deduplicated with a 256-slot open-addressing table (atomicCAS on the 32-bit key), values are
// accumulated per slot, and only the distinct keys (~35 for a sphere quad instead of ~80-130
// contributions) are ordered, by rank counting.  Contributions are added vertex position by
// vertex position; distinct background cells never share a vertex at the same local position,
// so on a proper mesh every shared-memory add inside one instruction hits a different slot and
// the summation order is the program order (bit-reproducible).  Through constraint lines several
// lanes may add to the same master in one instruction; that order is the hardware's.

__device__ inline unsigned hash_insert(unsigned *keys, unsigned dof)
{
  unsigned s = (dof * 2654435761u) >> 24;
  while (true) {
    const unsigned old = atomicCAS(&keys[s], HASH_EMPTY, dof);
    if (old == HASH_EMPTY || old == dof) return s;
    s = (s + 1) & (HASH_SLOTS - 1);
  }
}

constexpr unsigned NO_SLOT = 0xffffffffu;

// claim (or find) the slot of `key` in a SLOTS-entry open-addressing table; bounded probing gives up instead of spinning
template <int SLOTS>
__device__ __forceinline__ unsigned hash_claim(unsigned *keys, unsigned key, unsigned &claimed, bool &failed)
{
  unsigned s = ((key * 2654435761u) >> 16) & (SLOTS - 1);
  int probes = 0;
  while (true) {
    const unsigned old = atomicCAS(&keys[s], HASH_EMPTY, key);
    if (old == HASH_EMPTY) { ++claimed; return s; }
    if (old == key) return s;
    s = (s + 1) & (SLOTS - 1);
    if (++probes >= SLOTS) { failed = true; return NO_SLOT; }
  }
}

// vals[slot] += v for every lane of the group that carries a contribution (slot != NO_SLOT), in a FIXED order: lanes
// that hit the same slot in this instruction are found with match.any, the lowest of them adds the values in ascending
// lane order and is the only one that touches the slot.  No floating-point atomics: the sums are bit-reproducible.
__device__ __forceinline__ void group_accumulate(unsigned gmask, unsigned wlane, double *vals, unsigned slot, double v)
{
  const bool has = slot != NO_SLOT;
  const unsigned grp = __match_any_sync(gmask, has ? slot : (0x80000000u | wlane));   // a lane without a contribution matches nobody
  const bool leader = has && (unsigned)(__ffs(grp) - 1) == wlane;
  const unsigned mx = __reduce_max_sync(gmask, has ? (unsigned)__popc(grp) : 1u);
  double sum = leader ? vals[slot] + v : 0.0;   // ((slot + v_a) + v_b) + ...: the order a sequential loop over the contributions takes
  unsigned rest = grp & (grp - 1u);   // the other lanes of my slot, ascending
  for (unsigned j = 1; j < mx; ++j) {
    const bool more = rest != 0u;
    const int src = more ? __ffs(rest) - 1 : (int)wlane;
    rest &= rest - 1u;
    const double o = __shfl_sync(gmask, v, src);
    if (leader && more) sum += o;
  }
  if (leader) vals[slot] = sum;
  __syncwarp(gmask);
}

template <int NL>
__global__ void __launch_bounds__(256) merge_rows_hash(uint64_t n_im, const uint32_t *__restrict__ cap, const uint64_t *__restrict__ cand_ptr,
                                                       const uint32_t *__restrict__ cand_bg, const uint8_t *__restrict__ acc,
                                                       const double *__restrict__ local, const uint32_t *__restrict__ bg_dofs,
                                                       const int32_t *__restrict__ line_of, const uint64_t *__restrict__ c_ptr,
                                                       const uint32_t *__restrict__ c_master, const double *__restrict__ c_weight,
                                                       const uint64_t *__restrict__ rowstart, uint32_t *__restrict__ rowlen,
                                                       uint32_t *__restrict__ out_col, double *__restrict__ out_val)
{
  __shared__ unsigned s_keys[8][HASH_SLOTS];
  __shared__ double s_vals[8][HASH_SLOTS];
  __shared__ unsigned s_list[8][HASH_SLOTS];
  const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t m = (uint64_t)blockIdx.x * 8 + w;
  if (m >= n_im) return;
  const uint32_t cp = cap[m];
  if (cp > HASH_CAP) return;  // sort-based kernels take these
  if (cp == 0) { if (lane == 0) rowlen[m] = 0; return; }
  unsigned *keys = s_keys[w];
  double *vals = s_vals[w];
  unsigned *list = s_list[w];
#pragma unroll
  for (int k = 0; k < HASH_SLOTS / 32; ++k) { keys[lane + 32 * k] = HASH_EMPTY; vals[lane + 32 * k] = 0.0; }
  __syncwarp();
  const uint64_t p0 = cand_ptr[m], p1 = cand_ptr[m + 1];
  for (uint64_t base = p0; base < p1; base += 32) {
    const uint64_t p = base + lane;
    const bool on = p < p1 && acc[p];
    uint32_t dofs[NL];
    double v[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) { dofs[i] = 0; v[i] = 0.0; }
    if (on) {
      const uint32_t *dp = bg_dofs + (uint64_t)cand_bg[p] * NL;
      const double *lp = local + p * NL;
#pragma unroll
      for (int i = 0; i < NL; ++i) { dofs[i] = dp[i]; v[i] = lp[i]; }
    }
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      // one (pair, vertex) contribution per lane, added in lane order (group_accumulate): no floating-point atomics
      int32_t ln = -1;
      unsigned s = NO_SLOT;
      if (on) {
        ln = line_of[dofs[i]];
        s = hash_insert(keys, dofs[i]);   // keep_constrained_entries: a constrained row stays in the pattern, structurally zero
      }
      group_accumulate(0xffffffffu, lane, vals, (on && ln < 0) ? s : NO_SLOT, v[i]);
      const unsigned nm = (on && ln >= 0) ? (unsigned)(c_ptr[ln + 1] - c_ptr[ln]) : 0u;   // masters of the line (none for a Dirichlet row)
      const unsigned nm_max = __reduce_max_sync(0xffffffffu, nm);
      for (unsigned k = 0; k < nm_max; ++k) {
        unsigned t = NO_SLOT;
        double wv = 0.0;
        if (k < nm) { const uint64_t kk = c_ptr[ln] + k; t = hash_insert(keys, c_master[kk]); wv = c_weight[kk] * v[i]; }
        group_accumulate(0xffffffffu, lane, vals, t, wv);
      }
    }
  }
  // distinct keys -> list
  unsigned D = 0;
#pragma unroll
  for (int k = 0; k < HASH_SLOTS / 32; ++k) {
    const unsigned s = lane + 32 * k;
    const bool occ = keys[s] != HASH_EMPTY;
    const unsigned bal = __ballot_sync(0xffffffffu, occ);
    if (occ) list[D + __popc(bal & ((1u << lane) - 1u))] = s;
    D += __popc(bal);
  }
  __syncwarp();
  // rank of every key among the distinct keys = its position in the row
  const uint64_t dst = rowstart[m];
  for (unsigned e = lane; e < D; e += 32) {
    const unsigned s = list[e];
    const unsigned key = keys[s];
    unsigned rank = 0;
    for (unsigned f = 0; f < D; ++f) rank += keys[list[f]] < key ? 1u : 0u;
    out_col[dst + rank] = key;
    out_val[dst + rank] = vals[s];
  }
  if (lane == 0) rowlen[m] = D;
}

// Fast path used first: a warp merges FAST_ROWS consecutive immersed cells and carves their rows out
// of chunks it takes from a global cursor, so C comes out compact without a capacity pass, and the
// per-row counts of the transposed matrix are accumulated on the way.  Anything it cannot handle
// (more than HASH_CAP distinct dofs in a cell, storage overflow) raises a flag and the general
// two-pass path below redoes the whole matrix.
constexpr int FAST_ROWS = 64;   // consecutive immersed cells per group

// GROUP lanes cooperate on one immersed cell: a full warp in 3D (~10 accepted pairs x 8 vertices), 8 lanes
// in 2D (~3.4 pairs x 4 vertices), so that a warp merges four 2D cells at a time.
// LINES = false: the caller declared no constraint line at all (band-only backgrounds, unconstrained spaces) -- the
// per-contribution line lookup, a dependent scattered load in the middle of the chain, and the master loop drop out.
template <int NL, int GROUP, int SLOTS, bool LINES>
// resident blocks per SM the register allocation is held to; measured at sphere cycle 8: no bound (47 registers, 5
// blocks) 7.06 ms, 6 blocks 6.91 ms, 7 or 8 blocks 7.01 ms, 1 block (the compiler takes more registers) 10.2 ms
#ifndef NM_MERGE_MINB
#define NM_MERGE_MINB 6
#endif
#ifndef NM_MERGE_CBG
#define NM_MERGE_CBG 2
#endif
#ifndef NM_MERGE_PAIRSTEPS
#define NM_MERGE_PAIRSTEPS 1
#endif
#ifndef NM_MERGE_NOLINES_KERNEL
#define NM_MERGE_NOLINES_KERNEL 1
#endif
__global__ void __launch_bounds__(256, NM_MERGE_MINB) merge_rows_fast(uint64_t n_im, const uint64_t *__restrict__ cand_ptr,
                                                       const uint32_t *__restrict__ cand_bg, const uint8_t *__restrict__ acc,
                                                       const double *__restrict__ local, const uint32_t *__restrict__ bg_dofs,
                                                       const int32_t *__restrict__ line_of, const uint64_t *__restrict__ c_ptr,
                                                       const uint32_t *__restrict__ c_master, const double *__restrict__ c_weight,
                                                       uint64_t *__restrict__ rowstart, uint32_t *__restrict__ rowlen,
                                                       uint32_t *__restrict__ out_col, double *__restrict__ out_val, uint64_t capacity,
                                                       unsigned long long *cursor, unsigned *problem, uint32_t *__restrict__ a_cnt, int mark_bits,
                                                       uint32_t n_dofs, unsigned rows_per_group, unsigned chunk)
{
  constexpr int NG = 256 / GROUP;                 // groups per block
  constexpr unsigned CAP = SLOTS * 3 / 4;         // distinct dofs a group accepts
  __shared__ unsigned s_keys[NG][SLOTS];
  __shared__ double s_vals[NG][SLOTS];
  __shared__ unsigned short s_list[NG][SLOTS];   // slot of every distinct dof (SLOTS <= 256)
  __shared__ __align__(16) unsigned s_dkey[NG][SLOTS];
  __shared__ unsigned s_cand[NG][GROUP];
#if NM_MERGE_CBG == 1
  __shared__ unsigned s_cbg[NG][GROUP];   // background cell of every accepted pair of the chunk (loaded coalesced, once)
#elif NM_MERGE_CBG == 2
  // DoF rows of the accepted pairs of the chunk: every lane fetches the row of its own pair (one 32-byte sector) while
  // the chunk is compacted, so the contribution rounds below carry no dependent global load
  __shared__ __align__(16) unsigned s_dofs[NG][GROUP * NL];
#endif
  const unsigned g = threadIdx.x / GROUP, lane = threadIdx.x % GROUP;
  const unsigned gbase = (threadIdx.x & 31) - lane;                      // first lane of the group inside its warp
  const unsigned gmask = GROUP == 32 ? 0xffffffffu : (((1u << GROUP) - 1u) << gbase);
  unsigned *keys = s_keys[g];
  double *vals = s_vals[g];
  unsigned short *list = s_list[g];
  unsigned *dkey = s_dkey[g];
  unsigned *cands = s_cand[g];
#if NM_MERGE_CBG == 1
  unsigned *cbg = s_cbg[g];
#elif NM_MERGE_CBG == 2
  unsigned *sdofs = s_dofs[g];
  const bool vec_ok = (reinterpret_cast<uintptr_t>(bg_dofs) & 15u) == 0u;   // a table adopted in place may sit at any 4-byte address
#endif
  const uint64_t m0 = ((uint64_t)blockIdx.x * NG + g) * rows_per_group;
  uint64_t chunk_pos = 0;
  unsigned chunk_left = 0;
  for (uint64_t m = m0; m < m0 + rows_per_group && m < n_im; ++m) {
#pragma unroll
    for (int k = 0; k < SLOTS / GROUP; ++k) { keys[lane + GROUP * k] = HASH_EMPTY; vals[lane + GROUP * k] = 0.0; }
    __syncwarp(gmask);
    const uint64_t p0 = cand_ptr[m], p1 = cand_ptr[m + 1];
    unsigned claimed = 0;
    bool failed = false;   // table full: bounded probing gives up instead of spinning
    for (uint64_t base = p0; base < p1; base += GROUP) {
      // accepted pairs of this chunk, compacted; then one lane per (pair, vertex) contribution
      const uint64_t p = base + lane;
      const bool on = p < p1 && acc[p];
      const unsigned bal = (__ballot_sync(gmask, on) >> gbase) & (GROUP == 32 ? 0xffffffffu : ((1u << GROUP) - 1u));
#if NM_MERGE_CBG == 1
      if (on) { const unsigned k = __popc(bal & ((1u << lane) - 1u)); cands[k] = lane; cbg[k] = cand_bg[p]; }
#elif NM_MERGE_CBG == 2
      if (on) {
        const unsigned k = __popc(bal & ((1u << lane) - 1u));
        cands[k] = lane;
        const uint32_t *row = bg_dofs + (uint64_t)cand_bg[p] * NL;
        uint32_t rv[NL];
        if (vec_ok) {
#pragma unroll
          for (int j = 0; j < NL / 4; ++j) {
            const uint4 q4 = reinterpret_cast<const uint4 *>(row)[j];
            rv[4 * j] = q4.x; rv[4 * j + 1] = q4.y; rv[4 * j + 2] = q4.z; rv[4 * j + 3] = q4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < NL; ++j) rv[j] = row[j];
        }
        if (NM_MERGE_PAIRSTEPS) {
          // the pair-by-pair accumulation below needs the dofs of a cell to be distinct (they are, for any finite
          // element); a table that repeats one sends the whole matrix to the general path, which sums any multiset
          bool dup = false;
#pragma unroll
          for (int a = 0; a < NL; ++a)
#pragma unroll
            for (int b = a + 1; b < NL; ++b) dup |= rv[a] == rv[b];
          if (dup) atomicOr(problem, 8u);
        }
#pragma unroll
        for (int j = 0; j < NL / 4; ++j)
          reinterpret_cast<uint4 *>(sdofs + k * NL)[j] = make_uint4(rv[4 * j], rv[4 * j + 1], rv[4 * j + 2], rv[4 * j + 3]);
      }
#else
      if (on) cands[__popc(bal & ((1u << lane) - 1u))] = lane;
#endif
      __syncwarp(gmask);
      const unsigned n_entries = (unsigned)__popc(bal) * NL;
      for (unsigned e0 = 0; e0 < n_entries; e0 += GROUP) {
        const unsigned e = e0 + lane;
        const bool on_e = e < n_entries && !failed;
        double v = 0.0;
        int32_t ln = -1;
        unsigned s = NO_SLOT;
        if (on_e) {
          const uint64_t q = base + cands[e / NL];
          const unsigned i = e % NL;
#if NM_MERGE_CBG == 1
          const uint32_t dof = bg_dofs[(uint64_t)cbg[e / NL] * NL + i];
#elif NM_MERGE_CBG == 2
          const uint32_t dof = sdofs[e];
#else
          const uint32_t dof = bg_dofs[(uint64_t)cand_bg[q] * NL + i];
#endif
          v = local[q * NL + i];
          if (dof >= n_dofs) {   // the caller's table points outside the declared dofs: reported as NM_ERR_INVALID
            atomicOr(problem, 4u);
            failed = true;
          } else {
            if (LINES) ln = line_of[dof];
            s = hash_claim<SLOTS>(keys, dof, claimed, failed);
          }
        }
        // keep_constrained_entries: a constrained row stays in the pattern as a structural zero (its slot is claimed,
        // nothing is added); its value goes to the masters of the line (none for a Dirichlet row)
        // A round none of whose dofs sits on a line (all of them, without lines): the NL lanes of one pair hit NL
        // different slots -- the dofs of a cell are distinct, checked above -- so the pairs of the round add one after
        // the other, in the order of the contributions, with plain shared-memory read-modify-writes.
        bool plain_round = false;
        if constexpr (NM_MERGE_CBG == 2 && NM_MERGE_PAIRSTEPS != 0) plain_round = LINES ? !__any_sync(gmask, ln >= 0) : true;
        if (plain_round) {
          const bool has = on_e && !failed;
#pragma unroll
          for (int j = 0; j < GROUP / NL; ++j) {
            if (has && lane / NL == (unsigned)j) vals[s] += v;
            __syncwarp(gmask);
          }
        } else {
          group_accumulate(gmask, gbase + lane, vals, (on_e && !failed && ln < 0) ? s : NO_SLOT, v);
        }
        if (LINES && !plain_round) {
          const unsigned nm = (on_e && !failed && ln >= 0) ? (unsigned)(c_ptr[ln + 1] - c_ptr[ln]) : 0u;
          const unsigned nm_max = __reduce_max_sync(gmask, nm);
          for (unsigned k = 0; k < nm_max; ++k) {
            unsigned t = NO_SLOT;
            double wv = 0.0;
            if (k < nm && !failed) {
              const uint64_t kk = c_ptr[ln] + k;
              t = hash_claim<SLOTS>(keys, c_master[kk], claimed, failed);
              wv = c_weight[kk] * v;
            }
            group_accumulate(gmask, gbase + lane, vals, failed ? NO_SLOT : t, wv);
          }
        }
      }
    }
    const unsigned total_claimed = __reduce_add_sync(gmask, claimed);
    if (__any_sync(gmask, failed) || total_claimed > CAP) {
      if (lane == 0) { atomicOr(problem, 1u); rowlen[m] = 0; rowstart[m] = 0; }
      __syncwarp(gmask);
      continue;
    }
    unsigned D = 0;
#pragma unroll
    for (int k = 0; k < SLOTS / GROUP; ++k) {
      const unsigned s = lane + GROUP * k;
      const bool occ = keys[s] != HASH_EMPTY;
      const unsigned bal = (__ballot_sync(gmask, occ) >> gbase) & (GROUP == 32 ? 0xffffffffu : ((1u << GROUP) - 1u));
      if (occ) { const unsigned q = D + __popc(bal & ((1u << lane) - 1u)); list[q] = (unsigned short)s; dkey[q] = keys[s]; }  // dense list of the distinct dofs
      D += __popc(bal);
    }
    if (lane < 4) dkey[D + lane] = HASH_EMPTY;   // sentinel padding for the vector loads below (D <= 3/4 SLOTS)
    __syncwarp(gmask);
    if (D > chunk_left) {   // next chunk (the tail of the old one stays unused)
      unsigned long long base = 0;
      const unsigned need = D > chunk ? D : chunk;
      if (lane == 0) base = atomicAdd(cursor, (unsigned long long)need);
      base = __shfl_sync(gmask, base, gbase);
      if (base + need > capacity) {
        if (lane == 0) { atomicOr(problem, 2u); rowlen[m] = 0; rowstart[m] = 0; }
        chunk_left = 0;
        __syncwarp(gmask);
        continue;
      }
      chunk_pos = base;
      chunk_left = need;
    }
    const uint64_t dst = chunk_pos;
    // rank of every distinct key = its position in the row.  One key per lane and pass: the lane counts the smaller keys
    // of the dense list (vector loads, broadcast).  Measured at sphere cycle 8: same time as two keys per lane sharing
    // the loads (most cells hold 33..64 distinct dofs), and ranking a short second pass by ballots changes nothing.
    for (unsigned e0 = 0; e0 < D; e0 += GROUP) {
      const unsigned ea = e0 + lane;
      const unsigned ka = ea < D ? dkey[ea] : HASH_EMPTY;
      unsigned ra = 0;
      const uint4 *dk4 = reinterpret_cast<const uint4 *>(dkey);
      for (unsigned f = 0; f < (D + 3) / 4; ++f) {
        const uint4 k4 = dk4[f];
        ra += (k4.x < ka) + (k4.y < ka) + (k4.z < ka) + (k4.w < ka);
      }
      // per-dof entry counts for the transpose -- or, in compact-row mode, only the bitmap of touched dofs
      if (ea < D) {
        out_col[dst + ra] = ka; out_val[dst + ra] = vals[list[ea]];
        if (mark_bits) atomicOr(&a_cnt[ka >> 5], 1u << (ka & 31)); else atomicAdd(&a_cnt[ka], 1u);
      }
    }
    if (lane == 0) { rowlen[m] = D; rowstart[m] = dst; }
    chunk_pos += D;
    chunk_left -= D;
    __syncwarp(gmask);
  }
}

template <int NL>
__global__ void __launch_bounds__(256) merge_rows_warp(uint64_t n_im, const uint32_t *__restrict__ cap, const uint64_t *cand_ptr,
                                                       const uint32_t *cand_bg, const uint8_t *acc, const double *local,
                                                       const uint32_t *bg_dofs, const int32_t *line_of, const uint64_t *c_ptr,
                                                       const uint32_t *c_master, const double *c_weight, const uint64_t *rowstart,
                                                       uint32_t *rowlen, uint32_t *out_col, double *out_val)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long *s_key = reinterpret_cast<unsigned long long *>(smem) + (size_t)w * WARP_CAP;
  double *s_val = reinterpret_cast<double *>(smem + (size_t)(blockDim.x >> 5) * WARP_CAP * 8) + (size_t)w * WARP_CAP;
  const uint64_t m = (uint64_t)blockIdx.x * (blockDim.x >> 5) + w;
  if (m >= n_im) return;
  if (cap[m] > WARP_CAP || cap[m] <= HASH_CAP) return;  // block kernel / hash kernel take those
  merge_cell<NL, 32, WARP_CAP>(m, lane, s_key, s_val, nullptr, cand_ptr, cand_bg, acc, local, bg_dofs, line_of, c_ptr, c_master,
                               c_weight, rowstart, rowlen, out_col, out_val);
}

template <int NL>
__global__ void __launch_bounds__(256) merge_rows_block(unsigned n_big, const uint32_t *__restrict__ big, const uint64_t *cand_ptr,
                                                        const uint32_t *cand_bg, const uint8_t *acc, const double *local,
                                                        const uint32_t *bg_dofs, const int32_t *line_of, const uint64_t *c_ptr,
                                                        const uint32_t *c_master, const double *c_weight, const uint64_t *rowstart,
                                                        uint32_t *rowlen, uint32_t *out_col, double *out_val)
{
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ unsigned s_tmp[8];
  unsigned long long *s_key = reinterpret_cast<unsigned long long *>(smem);
  double *s_val = reinterpret_cast<double *>(smem + (size_t)BLOCK_CAP * 8);
  if (blockIdx.x >= n_big) return;
  merge_cell<NL, 256, BLOCK_CAP>(big[blockIdx.x], threadIdx.x, s_key, s_val, s_tmp, cand_ptr, cand_bg, acc, local, bg_dofs, line_of, c_ptr,
                                 c_master, c_weight, rowstart, rowlen, out_col, out_val);
}

// ---------------------------------------------------------------------------------------
// transpose C (rows = immersed cells) -> stored orientation (rows = space dofs)
// ---------------------------------------------------------------------------------------
// global dof -> row of the stored matrix: identity, or the rank of the dof among the touched dofs
struct RowMap {
  const uint32_t *bits;     // nullptr: identity
  const uint64_t *prefix;
  __device__ __forceinline__ uint32_t operator()(uint32_t dof) const
  {
    if (!bits) return dof;
    const uint32_t w = dof >> 5;
    return (uint32_t)prefix[w] + __popc(bits[w] & ((1u << (dof & 31)) - 1u));
  }
};

template <int LANES>
__global__ void t_mark(uint64_t n_im, const uint64_t *__restrict__ rowstart, const uint32_t *__restrict__ rowlen,
                       const uint32_t *__restrict__ col, uint32_t *__restrict__ bits)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t m = t / LANES;
  const unsigned l = t % LANES;
  if (m >= n_im) return;
  const uint64_t s = rowstart[m];
  const uint32_t n = rowlen[m];
  for (uint32_t k = l; k < n; k += LANES) { const uint32_t d = col[s + k]; atomicOr(&bits[d >> 5], 1u << (d & 31)); }
}

__global__ void bits_popc(uint64_t nw, const uint32_t *__restrict__ bits, uint32_t *__restrict__ out)
{
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < nw) out[w] = __popc(bits[w]);
  if (w == nw) out[w] = 0;
}

__global__ void row_ids_kernel(uint64_t nw, const uint32_t *__restrict__ bits, const uint64_t *__restrict__ prefix, uint32_t *__restrict__ ids)
{
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  uint32_t b = bits[w];
  uint64_t o = prefix[w];
  while (b) { const int k = __ffs(b) - 1; ids[o++] = (uint32_t)(w * 32 + k); b &= b - 1; }
}

// row offsets per GLOBAL dof from the compact ones (nm_get_csr): an untouched dof gets the offset of the next touched one
__global__ void expand_rowptr(uint64_t n_rows, RowMap map, const uint64_t *__restrict__ rowptr_local, uint64_t *__restrict__ out)
{
  const uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d <= n_rows) out[d] = rowptr_local[map((uint32_t)d)];
}

template <int LANES>
__global__ void t_count(uint64_t n_im, const uint64_t *__restrict__ rowstart, const uint32_t *__restrict__ rowlen,
                        const uint32_t *__restrict__ col, RowMap map, uint32_t *__restrict__ cnt)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t m = t / LANES;
  const unsigned l = t % LANES;
  if (m >= n_im) return;
  const uint64_t s = rowstart[m];
  const uint32_t n = rowlen[m];
  for (uint32_t k = l; k < n; k += LANES) atomicAdd(&cnt[map(col[s + k])], 1u);
}

// The chain per entry is col -> (row offset, cursor atomic) -> two stores, all to scattered addresses: the kernel is
// bound by memory latency, so every lane keeps T_UNROLL independent chains in flight.
constexpr int T_UNROLL = 4;
template <int LANES>
#ifndef NM_TSCATTER_MINB
#define NM_TSCATTER_MINB 6   // 6 and the unbounded default measure the same (5.40 ms for the transpose), 8 spills: 5.79 ms
#endif
#ifndef NM_TSORT_MINB
#define NM_TSORT_MINB 6
#endif
__global__ void __launch_bounds__(256, NM_TSCATTER_MINB) t_scatter(uint64_t n_im, const uint64_t *__restrict__ rowstart, const uint32_t *__restrict__ rowlen,
                          const uint32_t *__restrict__ col, const double *__restrict__ val, const uint32_t *__restrict__ im_dofs,
                          RowMap map, const uint64_t *__restrict__ a_rowptr, uint32_t *__restrict__ cursor, uint32_t *__restrict__ a_col,
                          double *__restrict__ a_val)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t m = t / LANES;
  const unsigned l = t % LANES;
  if (m >= n_im) return;
  const uint64_t s = rowstart[m];
  const uint32_t n = rowlen[m];
  const uint32_t c = im_dofs[m];
  for (uint32_t k0 = l; k0 < n; k0 += LANES * T_UNROLL) {
    uint32_t r[T_UNROLL];
    double v[T_UNROLL];
    uint64_t pos[T_UNROLL];
#pragma unroll
    for (int u = 0; u < T_UNROLL; ++u) {
      const uint32_t k = k0 + u * LANES;
      const bool ok = k < n;
      r[u] = ok ? col[s + k] : 0xffffffffu;
      v[u] = ok ? val[s + k] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < T_UNROLL; ++u)
      if (r[u] != 0xffffffffu) { r[u] = map(r[u]); }
#pragma unroll
    for (int u = 0; u < T_UNROLL; ++u)
      if (k0 + u * LANES < n) pos[u] = a_rowptr[r[u]] + atomicAdd(&cursor[r[u]], 1u);
#pragma unroll
    for (int u = 0; u < T_UNROLL; ++u)
      if (k0 + u * LANES < n) { a_col[pos[u]] = c; a_val[pos[u]] = v[u]; }
  }
}

// columns ascending in every row.  Rows are short (a space dof meets a handful of immersed
// cells): up to 8 entries are sorted in registers with a fixed network, longer rows in place.
__device__ inline void cswap(uint32_t &ka, double &va, uint32_t &kb, double &vb)
{
  const bool sw = ka > kb;
  const uint32_t k = sw ? kb : ka, K = sw ? ka : kb;
  const double v = sw ? vb : va, V = sw ? va : vb;
  ka = k; kb = K; va = v; vb = V;
}

// rowptr32 (optional): 32-bit copy of the row offsets for the SpMV, written on the way
__global__ void __launch_bounds__(256, NM_TSORT_MINB) t_sort_rows(uint64_t n_rows, const uint64_t *__restrict__ rowptr, uint32_t *__restrict__ col, double *__restrict__ val,
                            uint32_t *__restrict__ rowptr32)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const uint64_t s = rowptr[r], e = rowptr[r + 1];
  if (rowptr32) {
    rowptr32[r] = (uint32_t)s;
    if (r + 1 == n_rows) rowptr32[n_rows] = (uint32_t)e;
  }
  const uint64_t n = e - s;
  if (n <= 1) return;
  if (n <= 8) {
    uint32_t k[8];
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { const bool ok = (uint64_t)i < n; k[i] = ok ? col[s + i] : 0xffffffffu; v[i] = ok ? val[s + i] : 0.0; }
    // 19-comparator network for 8 keys
    cswap(k[0], v[0], k[1], v[1]); cswap(k[2], v[2], k[3], v[3]); cswap(k[4], v[4], k[5], v[5]); cswap(k[6], v[6], k[7], v[7]);
    cswap(k[0], v[0], k[2], v[2]); cswap(k[1], v[1], k[3], v[3]); cswap(k[4], v[4], k[6], v[6]); cswap(k[5], v[5], k[7], v[7]);
    cswap(k[1], v[1], k[2], v[2]); cswap(k[5], v[5], k[6], v[6]); cswap(k[0], v[0], k[4], v[4]); cswap(k[3], v[3], k[7], v[7]);
    cswap(k[1], v[1], k[5], v[5]); cswap(k[2], v[2], k[6], v[6]);
    cswap(k[1], v[1], k[4], v[4]); cswap(k[3], v[3], k[6], v[6]);
    cswap(k[2], v[2], k[4], v[4]); cswap(k[3], v[3], k[5], v[5]);
    cswap(k[3], v[3], k[4], v[4]);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if ((uint64_t)i < n) { col[s + i] = k[i]; val[s + i] = v[i]; }
    return;
  }
  for (uint64_t i = s + 1; i < e; ++i) {
    const uint32_t c = col[i];
    const double v = val[i];
    uint64_t j = i;
    while (j > s && col[j - 1] > c) { col[j] = col[j - 1]; val[j] = val[j - 1]; --j; }
    col[j] = c; val[j] = v;
  }
}

// ---- transpose with 32-bit offsets (nnz < 2^32): one scattered access less per entry on each side ----
// `ends[r + 1]` starts as the first slot of row r (the scan of the row counts writes its prefixes a second time, as 32 bits,
// one entry further) and is the cursor the scatter advances, so a position costs ONE atomic (no separate row-offset
// load), and when the scatter is done `ends` is the 32-bit row-offset array of the matrix.
#ifndef NM_T_ABS
#define NM_T_ABS 1
#endif
// NM_T_REC: the scatter writes one 16-byte (value, column) record per entry -- one sector instead of two -- and the row
// sort, which rewrites every entry anyway, splits the records into the column and value arrays.
#ifndef NM_T_REC
#define NM_T_REC 1
#endif
struct __align__(16) TRec { double v; uint32_t c; uint32_t pad; };

template <int LANES, bool REC>
__global__ void __launch_bounds__(256, NM_TSCATTER_MINB) t_scatter32(uint64_t n_im, const uint64_t *__restrict__ rowstart, const uint32_t *__restrict__ rowlen,
                          const uint32_t *__restrict__ col, const double *__restrict__ val, const uint32_t *__restrict__ im_dofs,
                          RowMap map, uint32_t *__restrict__ ends, uint32_t *__restrict__ a_col, double *__restrict__ a_val, TRec *__restrict__ rec)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t m = t / LANES;
  const unsigned l = t % LANES;
  if (m >= n_im) return;
  const uint64_t s = rowstart[m];
  const uint32_t n = rowlen[m];
  const uint32_t c = im_dofs[m];
  for (uint32_t k0 = l; k0 < n; k0 += LANES * T_UNROLL) {
    uint32_t r[T_UNROLL];
    double v[T_UNROLL];
    uint32_t pos[T_UNROLL];
#pragma unroll
    for (int u = 0; u < T_UNROLL; ++u) {
      const uint32_t k = k0 + u * LANES;
      const bool ok = k < n;
      r[u] = ok ? col[s + k] : 0xffffffffu;
      v[u] = ok ? val[s + k] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < T_UNROLL; ++u)
      if (r[u] != 0xffffffffu) { r[u] = map(r[u]); }
#pragma unroll
    for (int u = 0; u < T_UNROLL; ++u)
      if (k0 + u * LANES < n) pos[u] = atomicAdd(&ends[r[u] + 1], 1u);
#pragma unroll
    for (int u = 0; u < T_UNROLL; ++u)
      if (k0 + u * LANES < n) {
        if (REC) {
          uint4 w;
          w.x = (uint32_t)__double2loint(v[u]); w.y = (uint32_t)__double2hiint(v[u]); w.z = c; w.w = 0u;
          *reinterpret_cast<uint4 *>(rec + pos[u]) = w;
        } else { a_col[pos[u]] = c; a_val[pos[u]] = v[u]; }
      }
  }
}

template <bool REC>
__global__ void __launch_bounds__(256, NM_TSORT_MINB) t_sort_rows32(uint64_t n_rows, const uint32_t *__restrict__ rowptr, uint32_t *__restrict__ col,
                                                                    double *__restrict__ val, const TRec *__restrict__ rec)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const uint32_t s = rowptr[r], e = rowptr[r + 1];
  const uint32_t n = e - s;
  if (n == 0 || (!REC && n == 1)) return;
  if (n <= 8) {
    uint32_t k[8];
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool ok = (uint32_t)i < n;
      if (REC) {
        uint4 w = make_uint4(0u, 0u, 0xffffffffu, 0u);
        if (ok) w = *reinterpret_cast<const uint4 *>(rec + s + i);
        k[i] = w.z; v[i] = __hiloint2double((int)w.y, (int)w.x);
      } else { k[i] = ok ? col[s + i] : 0xffffffffu; v[i] = ok ? val[s + i] : 0.0; }
    }
    cswap(k[0], v[0], k[1], v[1]); cswap(k[2], v[2], k[3], v[3]); cswap(k[4], v[4], k[5], v[5]); cswap(k[6], v[6], k[7], v[7]);
    cswap(k[0], v[0], k[2], v[2]); cswap(k[1], v[1], k[3], v[3]); cswap(k[4], v[4], k[6], v[6]); cswap(k[5], v[5], k[7], v[7]);
    cswap(k[1], v[1], k[2], v[2]); cswap(k[5], v[5], k[6], v[6]); cswap(k[0], v[0], k[4], v[4]); cswap(k[3], v[3], k[7], v[7]);
    cswap(k[1], v[1], k[5], v[5]); cswap(k[2], v[2], k[6], v[6]);
    cswap(k[1], v[1], k[4], v[4]); cswap(k[3], v[3], k[6], v[6]);
    cswap(k[2], v[2], k[4], v[4]); cswap(k[3], v[3], k[5], v[5]);
    cswap(k[3], v[3], k[4], v[4]);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if ((uint32_t)i < n) { col[s + i] = k[i]; val[s + i] = v[i]; }
    return;
  }
  if (REC)
    for (uint32_t i = s; i < e; ++i) { const TRec q = rec[i]; col[i] = q.c; val[i] = q.v; }
  for (uint32_t i = s + 1; i < e; ++i) {
    const uint32_t c = col[i];
    const double v = val[i];
    uint32_t j = i;
    while (j > s && col[j - 1] > c) { col[j] = col[j - 1]; val[j] = val[j - 1]; --j; }
    col[j] = c; val[j] = v;
  }
}

// ---------------------------------------------------------------------------------------
// SpMV: LANES threads per row, FP64, shuffle reduction.  Each lane issues UNROLL independent
// (col, val, x[col]) load chains before the first FMA: with rows of 3..40 entries the kernels are
// bound by memory-level parallelism, not by arithmetic.
// ---------------------------------------------------------------------------------------
// Slots past the row's end are masked by predicates on the remaining length: they load entry 0 of x (an address that
// always exists) but never enter the sum -- x may be a scratch vector in local numbering of which only the addressed
// entries are written, or hold a non-finite entry of the caller, and 0 * NaN is NaN.
#ifndef NM_SPMV_PAD
#define NM_SPMV_PAD 1
#endif
template <int LANES, int UNROLL>
__device__ inline double row_dot(const uint32_t *__restrict__ col, const double *__restrict__ val, const double *__restrict__ x,
                                 uint64_t b, uint64_t e, unsigned l)
{
  double s = 0.0;
  for (uint64_t k0 = b + l; k0 < e; k0 += (uint64_t)LANES * UNROLL) {
    const uint64_t left = e - k0;
    const unsigned rem = left > 0xffffffffULL ? 0xffffffffu : (unsigned)left;   // entries from k0 to the row's end
    uint32_t c[UNROLL];
    double v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const bool ok = (unsigned)(u * LANES) < rem;
      c[u] = ok ? col[k0 + (uint64_t)(u * LANES)] : 0u;
      v[u] = ok ? val[k0 + (uint64_t)(u * LANES)] : 0.0;
    }
    double xv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) xv[u] = __ldg(x + c[u]);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
#if NM_SPMV_PAD
      if ((unsigned)(u * LANES) < rem) s = fma(v[u], xv[u], s);
#else
      s = fma(v[u], xv[u], s);
#endif
    }
  }
  return s;
}

template <int LANES, int UNROLL, typename PTR>
__global__ void __launch_bounds__(256) spmv_csr(uint64_t n_rows, const PTR *__restrict__ rowptr, const uint32_t *__restrict__ col,
                                                const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y,
                                                const uint32_t *__restrict__ row_ids)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t r = t / LANES;
  const unsigned l = threadIdx.x % LANES;
  double s = 0.0;
  if (r < n_rows) s = row_dot<LANES, UNROLL>(col, val, x, rowptr[r], rowptr[r + 1], l);
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, LANES);
  if (r < n_rows && l == 0) y[row_ids ? row_ids[r] : r] = s;   // compact-row mode: y was zeroed, row r is dof row_ids[r]
}

// Ct.vmult fused with its reduction over NVLink: every non-empty row result is added straight into
// the result vectors of all ranks (peer-mapped pointers, or ONE multicast pointer on which the
// NVSwitch performs the add in every rank's copy: multimem.red).  The vectors must be zero before any
// rank starts and are complete once every rank's kernel has finished -- no NCCL all-reduce, and only
// the rows a rank actually owns cross the links instead of the whole background vector twice.
struct PushTargets {
  double *p[16];
  int n;
};

template <int UNROLL, typename PTR, bool MULTICAST>
__global__ void __launch_bounds__(128) spmv_csr_push(uint64_t n_rows, const PTR *__restrict__ rowptr, const uint32_t *__restrict__ col,
                                                     const double *__restrict__ val, const double *__restrict__ x, PushTargets T,
                                                     const uint32_t *__restrict__ row_ids)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const uint64_t b = rowptr[r], e = rowptr[r + 1];
  if (b == e) return;
  const double s = row_dot<1, UNROLL>(col, val, x, b, e, 0u);
  const uint64_t g = row_ids ? row_ids[r] : r;
  if (MULTICAST) {
    asm volatile("multimem.red.relaxed.sys.global.add.f64 [%0], %1;" ::"l"(T.p[0] + g), "d"(s) : "memory");
  } else {
    for (int t = 0; t < T.n; ++t) asm volatile("red.relaxed.sys.global.add.f64 [%0], %1;" ::"l"(T.p[t] + g), "d"(s) : "memory");
  }
}

// Same product, reduce-scatter semantics: the background dofs are split into contiguous ranges of `rows_per_rank`
// (the last rank takes the remainder) and a row result is added only into the vector of the rank that owns the dof --
// what a distributed consumer needs (the reference's Ct*lambda is a Trilinos vector distributed by dof ownership).
// One 8-byte reduction per touched row crosses NVLink instead of one per row and rank.
template <int UNROLL, typename PTR>
__global__ void __launch_bounds__(128) spmv_csr_push_owned(uint64_t n_rows, const PTR *__restrict__ rowptr, const uint32_t *__restrict__ col,
                                                           const double *__restrict__ val, const double *__restrict__ x, PushTargets T,
                                                           const uint32_t *__restrict__ row_ids, uint64_t rows_per_rank)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const uint64_t b = rowptr[r], e = rowptr[r + 1];
  if (b == e) return;
  const double s = row_dot<1, UNROLL>(col, val, x, b, e, 0u);
  const uint64_t g = row_ids ? row_ids[r] : r;
  uint64_t owner = g / rows_per_rank;
  if (owner >= (uint64_t)T.n) owner = T.n - 1;
  asm volatile("red.relaxed.sys.global.add.f64 [%0], %1;" ::"l"(T.p[owner] + g), "d"(s) : "memory");
}

// rows of C: (start, len) storage, result scattered to the immersed dof of the cell
template <int LANES, int UNROLL>
__global__ void __launch_bounds__(256) spmv_crows(uint64_t n_im, const uint64_t *__restrict__ rowstart, const uint32_t *__restrict__ rowlen,
                                                  const uint32_t *__restrict__ col, const double *__restrict__ val,
                                                  const uint32_t *__restrict__ im_dofs, const double *__restrict__ x, double *__restrict__ y)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t r = t / LANES;
  const unsigned l = threadIdx.x % LANES;
  double s = 0.0;
  if (r < n_im) {
    const uint64_t b = rowstart[r];
    s = row_dot<LANES, UNROLL>(col, val, x, b, b + rowlen[r], l);
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, LANES);
  if (r < n_im && l == 0) y[im_dofs ? im_dofs[r] : r] = s;   // im_dofs == nullptr: result in local (cell) numbering
}

// compact copy of C's rows (for nm_get_csr(transpose = 1)): row index = immersed dof
template <int LANES>
__global__ void c_compact(uint64_t n_im, const uint64_t *__restrict__ rowstart, const uint32_t *__restrict__ rowlen,
                          const uint32_t *__restrict__ col, const double *__restrict__ val, const uint32_t *__restrict__ im_dofs,
                          const uint64_t *__restrict__ out_ptr, uint32_t *__restrict__ out_col, double *__restrict__ out_val)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t m = t / LANES;
  const unsigned l = t % LANES;
  if (m >= n_im) return;
  const uint64_t s = rowstart[m], d = out_ptr[im_dofs[m]];
  const uint32_t n = rowlen[m];
  for (uint32_t k = l; k < n; k += LANES) { out_col[d + k] = col[s + k]; if (out_val) out_val[d + k] = val[s + k]; }
}

// rows of the stored orientation that hold entries (full-row mode: most dofs of a band-only background have none)
__global__ void nonempty_flags(uint64_t n_rows, const uint64_t *__restrict__ rowptr, uint32_t *__restrict__ flag)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_rows) flag[r] = rowptr[r + 1] > rowptr[r] ? 1u : 0u;
  if (r == n_rows) flag[r] = 0u;
}

// row_ids / row_len of the non-empty rows; ids == nullptr: row r is dof r (full-row mode), pos = compaction offsets
__global__ void compact_rows_out(uint64_t n_rows, const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ ids,
                                 const uint64_t *__restrict__ pos, uint32_t *__restrict__ out_ids, uint32_t *__restrict__ out_len)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const uint64_t b = rowptr[r], e = rowptr[r + 1];
  if (pos) {
    if (e == b) return;
    const uint64_t o = pos[r];
    if (out_ids) out_ids[o] = (uint32_t)r;
    if (out_len) out_len[o] = (uint32_t)(e - b);
  } else {
    if (out_ids) out_ids[r] = ids[r];
    if (out_len) out_len[r] = (uint32_t)(e - b);
  }
}

// vectors in local numbering <-> the global numbering the kernels address: idx[i] = global position of local entry i
__global__ void scatter_by_index(uint64_t n, const uint32_t *__restrict__ idx, const double *__restrict__ local, double *__restrict__ global)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) global[idx[i]] = local[i];
}
__global__ void gather_by_index(uint64_t n, const uint32_t *__restrict__ idx, const double *__restrict__ global, double *__restrict__ local)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) local[i] = global[idx[i]];
}

// dof tables come from the caller: an index beyond the declared dof count would send the merge out of bounds
__global__ void check_index_range(uint64_t n, const uint32_t *__restrict__ idx, uint32_t limit, unsigned *bad)
{
  unsigned b = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) b |= idx[i] >= limit ? 1u : 0u;
  if (__any_sync(0xffffffffu, b) && (threadIdx.x & 31) == 0) atomicOr(bad, 1u);
}

__global__ void scatter_len_by_dof(uint64_t n_im, const uint32_t *__restrict__ rowlen, const uint32_t *__restrict__ im_dofs, uint32_t *__restrict__ out)
{
  const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m < n_im) out[im_dofs[m]] = rowlen[m];
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
template <int NL>
static int transpose_rows(nm_ctx *ctx, bool counts_ready);

// fast path: returns NM_OK and sets *done when the matrix was completed
template <int NL>
static int matrix_build_fast(nm_ctx *ctx, bool *done)
{
  const uint64_t n_im = ctx->n_im, n_rows = ctx->n_space_dofs;
  *done = false;
  if (n_im == 0 || ctx->force_general_merge) return NM_OK;
  PhaseTimer tm(ctx, NM_T_MERGE);
  // distinct dofs <= contributions; +25% and the chunk tails for constraint expansion
  constexpr int GROUP = NL == 8 ? 32 : 8, NG = 256 / GROUP;
  // consecutive immersed cells per group: FAST_ROWS when the grid is large, fewer on small grids so that every SM still
  // gets a few blocks (a 2 K-cell curve on one block of 64-row groups took as long as a 500 K-cell one)
  uint64_t rpg = (n_im + (uint64_t)NG * 4 * ctx->sm_count - 1) / ((uint64_t)NG * 4 * ctx->sm_count);
  rpg = rpg < 1 ? 1 : (rpg > FAST_ROWS ? FAST_ROWS : rpg);
  const unsigned chunk = (unsigned)(rpg * (NL == 8 ? 16 : 4));   // entries a group takes from the cursor at a time
  const uint64_t blocks = (n_im + NG * rpg - 1) / (NG * rpg);
  // distinct dofs <= contributions; +25% and the chunk tails for constraint expansion
  const uint64_t capacity = ctx->counts.n_accepted * NL + ctx->counts.n_accepted * NL / 4 + blocks * NG * chunk + 1024;
  NM_TRY(nm_ensure(ctx, ctx->c_rowstart, (n_im + 2) * sizeof(uint64_t)));
  NM_TRY(nm_ensure(ctx, ctx->c_rowlen, (n_im + 1) * sizeof(uint32_t)));
  NM_TRY(nm_ensure(ctx, ctx->c_col, (capacity + 1) * sizeof(uint32_t)));
  NM_TRY(nm_ensure(ctx, ctx->c_val, (capacity + 1) * sizeof(double)));
  NM_TRY(nm_ensure(ctx, ctx->counters, 64 * sizeof(uint64_t)));
  unsigned long long *cursor = ctx->counters.as<unsigned long long>() + 20;
  unsigned *problem = ctx->counters.as<unsigned>() + 44;
  // Hash slots per cell: 3/4 of them is the number of distinct dofs the fast path accepts.  The small table (half the
  // clearing and compaction work: 7.6 -> 7.1 ms at sphere cycle 8) is tried first when the cells average few accepted
  // pairs; a cell that overflows it raises the flag and the large table redoes the merge before the general path would.
  const bool small_first = NL == 8 && !ctx->merge_large_table && ctx->counts.n_accepted <= 12 * n_im;
  unsigned h_problem = 0;
  for (int attempt = small_first ? 0 : 1; attempt < 2; ++attempt) {
    uint32_t *dof_marks = nullptr;   // per-dof counts, or the bitmap of touched dofs in compact-row mode
    if (ctx->compact_rows) {
      const uint64_t nw = n_rows / 32 + 2;
      NM_TRY(nm_ensure(ctx, ctx->row_bits, nw * sizeof(uint32_t)));
      NM_CUDA(cudaMemsetAsync(ctx->row_bits.p, 0, nw * sizeof(uint32_t), ctx->stream));
      dof_marks = ctx->row_bits.as<uint32_t>();
    } else {
      NM_TRY(nm_ensure(ctx, ctx->a_cursor, (n_rows + 2) * sizeof(uint32_t) * 2));
      NM_CUDA(cudaMemsetAsync(ctx->a_cursor.p, 0, (n_rows + 2) * sizeof(uint32_t) * 2, ctx->stream));
      dof_marks = ctx->a_cursor.as<uint32_t>();
    }
    NM_CUDA(cudaMemsetAsync(cursor, 0, 32, ctx->stream));   // bytes 160..191: the chunk cursor AND the problem flag (byte 176)
    KernelTimer km(ctx, NM_T_K_MERGE);
    auto launch = [&](auto kernel) {
      kernel<<<(unsigned)blocks, 256, 0, NM_LS(ctx)>>>(
          n_im, ctx->cand_ptr.as<uint64_t>(), ctx->cand_bg.as<uint32_t>(), ctx->cand_acc.as<uint8_t>(), ctx->cand_local.as<double>(),
          ctx->bg_dofs.as<uint32_t>(), ctx->line_of.as<int32_t>(), ctx->c_ptr.as<uint64_t>(), ctx->c_master.as<uint32_t>(),
          ctx->c_weight.as<double>(), ctx->c_rowstart.as<uint64_t>(), ctx->c_rowlen.as<uint32_t>(), ctx->c_col.as<uint32_t>(),
          ctx->c_val.as<double>(), capacity, cursor, problem, dof_marks, ctx->compact_rows ? 1 : 0, (uint32_t)n_rows, (unsigned)rpg, chunk);
    };
    bool launched = false;
    const bool lines = NM_MERGE_NOLINES_KERNEL == 0 || ctx->n_constrained != 0;
    if constexpr (NL == 8) {
      if (attempt == 0) {
        if (lines) launch(merge_rows_fast<NL, GROUP, 128, true>); else launch(merge_rows_fast<NL, GROUP, 128, false>);
        ctx->merge_slots = 128; launched = true;
      }
    }
    if (!launched) {
      constexpr int SL = NL == 8 ? 256 : 64;
      if (lines) launch(merge_rows_fast<NL, GROUP, SL, true>); else launch(merge_rows_fast<NL, GROUP, SL, false>);
      ctx->merge_slots = SL;
    }
    km.stop();
    NM_CUDA(cudaGetLastError());
    NM_CUDA(cudaMemcpyAsync(&h_problem, problem, 4, cudaMemcpyDeviceToHost, ctx->stream));
    NM_SYNC(ctx);
    if (attempt == 0 && h_problem == 1u) { ctx->merge_large_table = true; continue; }   // only "too many distinct dofs": the large table may take it
    break;
  }
  tm.stop();
  if (h_problem & 4u) return nm_fail(ctx, NM_ERR_INVALID, "background DoF table holds indices >= n_dofs (%llu)", (unsigned long long)n_rows);
  if (h_problem) return NM_OK;  // the general path redoes everything
  NM_TRY(transpose_rows<NL>(ctx, true));
  ctx->merge_path = 0;
  *done = true;
  return NM_OK;
}

template <int NL>
static int matrix_build_t(nm_ctx *ctx)
{
  const uint64_t n_im = ctx->n_im, n_rows = ctx->n_space_dofs;
  ctx->matrix_valid = false;
  ctx->out_rows_valid = false;
  ctx->nnz = 0;
  // sharded grids carry the global dof numbering: keep per-dof arrays out of the step when this rank can touch only a
  // small part of the dofs (each cell has 2^d of them)
  ctx->n_crows = n_im;                 // FE_DGQ(0): one row of C per immersed cell, its column id = the cell's dof
  ctx->crow_dofs = &ctx->im_dofs;
  ctx->compact_rows = ctx->compact_rows_env >= 0 ? ctx->compact_rows_env == 1 : n_rows > 4 * ctx->n_bg;
  ctx->n_local_rows = n_rows;
  {
    bool done = false;
    NM_TRY(matrix_build_fast<NL>(ctx, &done));
    if (done) return NM_OK;
  }
  ctx->merge_path = 1;
  PhaseTimer tm(ctx, NM_T_MERGE);
  NM_TRY(nm_ensure(ctx, ctx->scratch_a, (n_im + 1) * sizeof(uint32_t)));
  NM_TRY(nm_ensure(ctx, ctx->c_rowstart, (n_im + 2) * sizeof(uint64_t)));
  NM_TRY(nm_ensure(ctx, ctx->c_rowlen, (n_im + 1) * sizeof(uint32_t)));
  NM_TRY(nm_ensure(ctx, ctx->counters, 64 * sizeof(uint64_t)));
  unsigned *hdr = ctx->counters.as<unsigned>() + 32;  // [0] max cap, [1] n_big, [2] cursor, [3] rows above HASH_CAP
  NM_CUDA(cudaMemsetAsync(hdr, 0, 16, ctx->stream));
  NM_CUDA(cudaMemsetAsync(ctx->scratch_a.as<uint32_t>() + n_im, 0, 4, ctx->stream));
  if (n_im) {
    row_capacity_kernel<NL><<<nm_blocks(n_im, 128), 128, 0, NM_LS(ctx)>>>(
        n_im, ctx->cand_ptr.as<uint64_t>(), ctx->cand_bg.as<uint32_t>(), ctx->cand_acc.as<uint8_t>(), ctx->bg_dofs.as<uint32_t>(),
        ctx->line_of.as<int32_t>(), ctx->c_ptr.as<uint64_t>(), ctx->scratch_a.as<uint32_t>(), hdr, hdr + 1, (uint32_t)n_rows,
        ctx->counters.as<unsigned>() + 48);
    NM_CUDA(cudaGetLastError());
  }
  NM_TRY(nm_exclusive_scan_u32_to_u64(ctx, ctx->scratch_a.as<uint32_t>(), ctx->c_rowstart.as<uint64_t>(), n_im + 1));
  uint64_t total_cap = 0;
  unsigned h_hdr[4];
  NM_CUDA(cudaMemcpyAsync(&total_cap, ctx->c_rowstart.as<uint64_t>() + n_im, 8, cudaMemcpyDeviceToHost, ctx->stream));
  NM_CUDA(cudaMemcpyAsync(h_hdr, hdr, 16, cudaMemcpyDeviceToHost, ctx->stream));
  unsigned h_bad_dof = 0;
  NM_CUDA(cudaMemcpyAsync(&h_bad_dof, ctx->counters.as<unsigned>() + 48, 4, cudaMemcpyDeviceToHost, ctx->stream));
  NM_SYNC(ctx);
  if (h_bad_dof) return nm_fail(ctx, NM_ERR_INVALID, "background DoF table holds indices >= n_dofs (%llu)", (unsigned long long)n_rows);
  if (h_hdr[0] > (unsigned)BLOCK_CAP)
    return nm_fail(ctx, NM_ERR_UNSUPPORTED,
                   "an immersed cell couples to %u (dof, value) contributions; at most %d per immersed cell are implemented "
                   "(immersed cells far larger than background cells)", h_hdr[0], BLOCK_CAP);
  NM_TRY(nm_ensure(ctx, ctx->c_col, (total_cap + 1) * sizeof(uint32_t)));
  NM_TRY(nm_ensure(ctx, ctx->c_val, (total_cap + 1) * sizeof(double)));
  if (n_im) {
    const int warps = 8;
    const size_t smem = (size_t)warps * WARP_CAP * 16;
    NM_CUDA(cudaFuncSetAttribute(merge_rows_warp<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KernelTimer km(ctx, NM_T_K_MERGE);
    merge_rows_hash<NL><<<nm_blocks(n_im, 8), 256, 0, NM_LS(ctx)>>>(
        n_im, ctx->scratch_a.as<uint32_t>(), ctx->cand_ptr.as<uint64_t>(), ctx->cand_bg.as<uint32_t>(), ctx->cand_acc.as<uint8_t>(),
        ctx->cand_local.as<double>(), ctx->bg_dofs.as<uint32_t>(), ctx->line_of.as<int32_t>(), ctx->c_ptr.as<uint64_t>(),
        ctx->c_master.as<uint32_t>(), ctx->c_weight.as<double>(), ctx->c_rowstart.as<uint64_t>(), ctx->c_rowlen.as<uint32_t>(),
        ctx->c_col.as<uint32_t>(), ctx->c_val.as<double>());
    km.stop();
    if (h_hdr[3])
    merge_rows_warp<NL><<<nm_blocks(n_im, warps), warps * 32, smem, NM_LS(ctx)>>>(
        n_im, ctx->scratch_a.as<uint32_t>(), ctx->cand_ptr.as<uint64_t>(), ctx->cand_bg.as<uint32_t>(), ctx->cand_acc.as<uint8_t>(),
        ctx->cand_local.as<double>(), ctx->bg_dofs.as<uint32_t>(), ctx->line_of.as<int32_t>(), ctx->c_ptr.as<uint64_t>(),
        ctx->c_master.as<uint32_t>(), ctx->c_weight.as<double>(), ctx->c_rowstart.as<uint64_t>(), ctx->c_rowlen.as<uint32_t>(),
        ctx->c_col.as<uint32_t>(), ctx->c_val.as<double>());
    NM_CUDA(cudaGetLastError());
    if (h_hdr[1]) {
      NM_TRY(nm_ensure(ctx, ctx->scratch_b, (size_t)h_hdr[1] * 4 + 16));
      list_big_rows<<<nm_blocks(n_im, 256), 256, 0, NM_LS(ctx)>>>(n_im, ctx->scratch_a.as<uint32_t>(), ctx->scratch_b.as<uint32_t>(), hdr + 2);
      const size_t smem_b = (size_t)BLOCK_CAP * 16;
      NM_CUDA(cudaFuncSetAttribute(merge_rows_block<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
      merge_rows_block<NL><<<h_hdr[1], 256, smem_b, NM_LS(ctx)>>>(
          h_hdr[1], ctx->scratch_b.as<uint32_t>(), ctx->cand_ptr.as<uint64_t>(), ctx->cand_bg.as<uint32_t>(), ctx->cand_acc.as<uint8_t>(),
          ctx->cand_local.as<double>(), ctx->bg_dofs.as<uint32_t>(), ctx->line_of.as<int32_t>(), ctx->c_ptr.as<uint64_t>(),
          ctx->c_master.as<uint32_t>(), ctx->c_weight.as<double>(), ctx->c_rowstart.as<uint64_t>(), ctx->c_rowlen.as<uint32_t>(),
          ctx->c_col.as<uint32_t>(), ctx->c_val.as<double>());
      NM_CUDA(cudaGetLastError());
    }
  }
  tm.stop();

  return transpose_rows<NL>(ctx, false);
}

template <int NL>
static int transpose_rows(nm_ctx *ctx, bool counts_ready)
{
  const uint64_t n_im = ctx->n_crows, n_dofs = ctx->n_space_dofs;   // rows of C: immersed cells (FE_DGQ(0)) or (cell, dof) pairs
  PhaseTimer tt(ctx, NM_T_TRANSPOSE);
  constexpr int TL = NL == 8 ? 8 : 4;
  RowMap map{nullptr, nullptr};
  uint64_t n_rows = n_dofs;   // rows of the stored matrix: all dofs, or only the touched ones
  if (ctx->compact_rows) {
    const uint64_t nw = n_dofs / 32 + 1;   // the bitmap holds one more (zero) word so that rank(n_dofs) is defined
    if (!counts_ready) {                   // general merge path: the bitmap has not been filled yet
      NM_TRY(nm_ensure(ctx, ctx->row_bits, (nw + 1) * sizeof(uint32_t)));
      NM_CUDA(cudaMemsetAsync(ctx->row_bits.p, 0, (nw + 1) * sizeof(uint32_t), ctx->stream));
      if (n_im) {
        t_mark<TL><<<nm_blocks(n_im * TL, 256), 256, 0, NM_LS(ctx)>>>(n_im, ctx->c_rowstart.as<uint64_t>(), ctx->c_rowlen.as<uint32_t>(),
                                                                      ctx->c_col.as<uint32_t>(), ctx->row_bits.as<uint32_t>());
        NM_CUDA(cudaGetLastError());
      }
    }
    NM_TRY(nm_ensure(ctx, ctx->scratch_b, (nw + 2) * sizeof(uint32_t)));
    NM_TRY(nm_ensure(ctx, ctx->row_prefix, (nw + 2) * sizeof(uint64_t)));
    bits_popc<<<nm_blocks(nw + 1, 256), 256, 0, NM_LS(ctx)>>>(nw, ctx->row_bits.as<uint32_t>(), ctx->scratch_b.as<uint32_t>());
    NM_CUDA(cudaGetLastError());
    NM_TRY(nm_exclusive_scan_u32_to_u64(ctx, ctx->scratch_b.as<uint32_t>(), ctx->row_prefix.as<uint64_t>(), nw + 1));
    NM_CUDA(cudaMemcpyAsync(&n_rows, ctx->row_prefix.as<uint64_t>() + nw, 8, cudaMemcpyDeviceToHost, ctx->stream));
    NM_SYNC(ctx);
    NM_TRY(nm_ensure(ctx, ctx->row_ids, (n_rows + 1) * sizeof(uint32_t)));
    row_ids_kernel<<<nm_blocks(nw, 256), 256, 0, NM_LS(ctx)>>>(nw, ctx->row_bits.as<uint32_t>(), ctx->row_prefix.as<uint64_t>(),
                                                               ctx->row_ids.as<uint32_t>());
    NM_CUDA(cudaGetLastError());
    map = RowMap{ctx->row_bits.as<uint32_t>(), ctx->row_prefix.as<uint64_t>()};
    counts_ready = false;   // the merge marked bits, the counts per compact row are taken below
  }
  ctx->n_local_rows = n_rows;
  NM_TRY(nm_ensure(ctx, ctx->a_cursor, (n_rows + 2) * sizeof(uint32_t) * 2));
  NM_TRY(nm_ensure(ctx, ctx->a_rowptr, (n_rows + 2) * sizeof(uint64_t)));
  uint32_t *cnt = ctx->a_cursor.as<uint32_t>(), *cursor = cnt + (n_rows + 2);
  if (!counts_ready) NM_CUDA(cudaMemsetAsync(cnt, 0, (n_rows + 2) * sizeof(uint32_t) * 2, ctx->stream));
  if (n_im && !counts_ready) {
    t_count<TL><<<nm_blocks(n_im * TL, 256), 256, 0, NM_LS(ctx)>>>(n_im, ctx->c_rowstart.as<uint64_t>(), ctx->c_rowlen.as<uint32_t>(),
                                                                   ctx->c_col.as<uint32_t>(), map, cnt);
    NM_CUDA(cudaGetLastError());
  }
  if (NM_T_ABS) {
    // the offsets also as 32 bits, shifted by one entry: `ends` (see t_scatter32) comes out of the scan itself
    NM_TRY(nm_ensure(ctx, ctx->a_rowptr32, (n_rows + 3) * sizeof(uint32_t)));
    NM_CUDA(cudaMemsetAsync(ctx->a_rowptr32.p, 0, sizeof(uint32_t), ctx->stream));
    NM_TRY(nm_exclusive_scan_u32_to_u64_and_u32(ctx, cnt, ctx->a_rowptr.as<uint64_t>(), ctx->a_rowptr32.as<uint32_t>() + 1, n_rows + 1));
  } else
  NM_TRY(nm_exclusive_scan_u32_to_u64(ctx, cnt, ctx->a_rowptr.as<uint64_t>(), n_rows + 1));
  uint64_t nnz = 0;
  NM_CUDA(cudaMemcpyAsync(&nnz, ctx->a_rowptr.as<uint64_t>() + n_rows, 8, cudaMemcpyDeviceToHost, ctx->stream));
  NM_SYNC(ctx);
  ctx->nnz = nnz;
  NM_TRY(nm_ensure(ctx, ctx->a_col, (nnz + 1) * sizeof(uint32_t)));
  NM_TRY(nm_ensure(ctx, ctx->a_val, (nnz + 1) * sizeof(double)));
  if (NM_T_ABS && nnz < 0xffffffffULL) {
    // 32-bit offsets: the offset array doubles as the scatter cursor (see t_scatter32)
    ctx->a_rowptr32_valid = false;
    if (NM_T_REC) NM_TRY(nm_ensure(ctx, ctx->t_rec, (nnz + 1) * sizeof(TRec)));
    uint32_t *ends = ctx->a_rowptr32.as<uint32_t>();   // filled by the scan above
    TRec *rec = NM_T_REC ? ctx->t_rec.as<TRec>() : nullptr;
    if (n_im) {
      t_scatter32<TL, NM_T_REC != 0><<<nm_blocks(n_im * TL, 256), 256, 0, NM_LS(ctx)>>>(
          n_im, ctx->c_rowstart.as<uint64_t>(), ctx->c_rowlen.as<uint32_t>(), ctx->c_col.as<uint32_t>(), ctx->c_val.as<double>(),
          ctx->crow_dofs->as<uint32_t>(), map, ends, ctx->a_col.as<uint32_t>(), ctx->a_val.as<double>(), rec);
      NM_CUDA(cudaGetLastError());
    }
    if (n_rows) {
      // (a warp-cooperative version -- the 32 rows of a warp staged in shared memory with coalesced loads and stores --
      // measured 5.10 ms for the transpose against 4.97 ms: L1 already captures the reuse of the thread-per-row loads)
      t_sort_rows32<NM_T_REC != 0><<<nm_blocks(n_rows, 256), 256, 0, NM_LS(ctx)>>>(n_rows, ends, ctx->a_col.as<uint32_t>(), ctx->a_val.as<double>(), rec);
      NM_CUDA(cudaGetLastError());
    }
    ctx->a_rowptr32_valid = true;
    tt.stop();
    ctx->matrix_valid = true;
    return NM_OK;
  }
  if (n_im) {
    t_scatter<TL><<<nm_blocks(n_im * TL, 256), 256, 0, NM_LS(ctx)>>>(
        n_im, ctx->c_rowstart.as<uint64_t>(), ctx->c_rowlen.as<uint32_t>(), ctx->c_col.as<uint32_t>(), ctx->c_val.as<double>(),
        ctx->crow_dofs->as<uint32_t>(), map, ctx->a_rowptr.as<uint64_t>(), cursor, ctx->a_col.as<uint32_t>(), ctx->a_val.as<double>());
    NM_CUDA(cudaGetLastError());
  }
  // 32-bit row offsets for the SpMV when they fit (halves the row-pointer traffic), written by the sort kernel
  ctx->a_rowptr32_valid = false;
  uint32_t *rp32 = nullptr;
  if (nnz < 0xffffffffULL) {
    NM_TRY(nm_ensure(ctx, ctx->a_rowptr32, (n_rows + 2) * sizeof(uint32_t)));
    rp32 = ctx->a_rowptr32.as<uint32_t>();
    if (n_rows == 0) NM_CUDA(cudaMemsetAsync(rp32, 0, 8, ctx->stream));
    ctx->a_rowptr32_valid = true;
  }
  if (n_rows) {
    t_sort_rows<<<nm_blocks(n_rows, 256), 256, 0, NM_LS(ctx)>>>(n_rows, ctx->a_rowptr.as<uint64_t>(), ctx->a_col.as<uint32_t>(), ctx->a_val.as<double>(),
                                                                rp32);
    NM_CUDA(cudaGetLastError());
  }
  tt.stop();
  ctx->matrix_valid = true;
  return NM_OK;
}

// row offsets of the stored orientation, one per global dof (+1), whatever the internal row mode
int nm_stored_rowptr_global(nm_ctx *ctx, uint64_t *rowptr_dev)
{
  const uint64_t n_dofs = ctx->n_space_dofs;
  if (!ctx->compact_rows) {
    NM_CUDA(cudaMemcpyAsync(rowptr_dev, ctx->a_rowptr.p, (n_dofs + 1) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return NM_OK;
  }
  const RowMap map{ctx->row_bits.as<uint32_t>(), ctx->row_prefix.as<uint64_t>()};
  expand_rowptr<<<nm_blocks(n_dofs + 1, 256), 256, 0, NM_LS(ctx)>>>(n_dofs, map, ctx->a_rowptr.as<uint64_t>(), rowptr_dev);
  NM_CUDA(cudaGetLastError());
  return NM_OK;
}

// The immersed DoF table becomes column indices the SpMV dereferences: verified here (one small pass per upload).
// The background table is verified where it is read, by the merge kernels.
static int check_dof_tables(nm_ctx *ctx)
{
  unsigned *bad = ctx->counters.as<unsigned>() + 48;   // [48] background table (merge kernels), [49] immersed table
  NM_CUDA(cudaMemsetAsync(bad, 0, 8, ctx->stream));
  if (ctx->dofs_checked) return NM_OK;
  const uint64_t ni = ctx->n_im;
  if (ni) check_index_range<<<nm_blocks(ni, 256 * 8), 256, 0, NM_LS(ctx)>>>(ni, ctx->im_dofs.as<uint32_t>(), (uint32_t)ctx->n_im_dofs, bad + 1);
  NM_CUDA(cudaGetLastError());
  unsigned h = 0;
  NM_CUDA(cudaMemcpyAsync(&h, bad + 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
  NM_SYNC(ctx);
  if (h) return nm_fail(ctx, NM_ERR_INVALID, "immersed DoF table holds indices >= n_dofs (%llu)", (unsigned long long)ctx->n_im_dofs);
  ctx->dofs_checked = true;
  return NM_OK;
}

// C (rows described by ctx->n_crows / crow_dofs, storage c_rowstart / c_rowlen / c_col / c_val) -> stored orientation
int nm_transpose_build(nm_ctx *ctx) { return transpose_rows<8>(ctx, false); }

int nm_matrix_build(nm_ctx *ctx)
{
  NM_TRY(check_dof_tables(ctx));
  if (ctx->download_pending) {   // an asynchronous read-back of the previous matrix may still be reading the buffers rebuilt here
    NM_TRY(nm_wait_download_issued(ctx));
    NM_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_down_done, 0));
    ctx->download_pending = false;
  }
  return ctx->spacedim == 3 ? matrix_build_t<8>(ctx) : matrix_build_t<4>(ctx);
}

// flags: NM_SPMV_LOCAL_OUT -- y in local numbering (transpose = 0: y[r] for stored row r, compact-row mode only;
// transpose = 1: y[m] for immersed cell m) and nothing outside the computed entries is written.
int nm_spmv_run(nm_ctx *ctx, int transpose, const double *x, double *y, int flags)
{
  const uint64_t n_rows = ctx->n_local_rows, n_im = ctx->n_crows;
  const bool local_out = (flags & NM_SPMV_LOCAL_OUT) != 0;
  PhaseTimer tm(ctx, transpose ? NM_T_SPMV_T : NM_T_SPMV);
  if (!transpose) {
    const uint32_t *ids = nullptr;
    if (ctx->compact_rows && !local_out) {   // only the touched rows are computed; the rest of y is zero
      NM_CUDA(cudaMemsetAsync(y, 0, ctx->n_space_dofs * sizeof(double), ctx->stream));
      ids = ctx->row_ids.as<uint32_t>();
    }
    const double avg = n_rows ? (double)ctx->nnz / (double)n_rows : 0;
    const uint64_t *rp = ctx->a_rowptr.as<uint64_t>();
    const uint32_t *cl = ctx->a_col.as<uint32_t>();
    const double *vl = ctx->a_val.as<double>();
    if (n_rows) {
      // rows of the coupling matrix are short (a space dof meets 3-8 immersed cells): one thread per row with eight
      // independent load chains is the fastest shape measured; longer rows get 4 or 8 lanes
      if (avg > 24) spmv_csr<8, 4, uint64_t><<<nm_blocks(n_rows * 8, 256), 256, 0, NM_LS(ctx)>>>(n_rows, rp, cl, vl, x, y, ids);
      else if (avg > 8) spmv_csr<4, 4, uint64_t><<<nm_blocks(n_rows * 4, 256), 256, 0, NM_LS(ctx)>>>(n_rows, rp, cl, vl, x, y, ids);
      else if (ctx->a_rowptr32_valid) spmv_csr<1, 8, uint32_t><<<nm_blocks(n_rows, 128), 128, 0, NM_LS(ctx)>>>(n_rows, ctx->a_rowptr32.as<uint32_t>(), cl, vl, x, y, ids);
      else spmv_csr<1, 8, uint64_t><<<nm_blocks(n_rows, 128), 128, 0, NM_LS(ctx)>>>(n_rows, rp, cl, vl, x, y, ids);
    }
  } else {
    if (!local_out) NM_CUDA(cudaMemsetAsync(y, 0, ctx->n_im_dofs * sizeof(double), ctx->stream));
    const double avg = n_im ? (double)ctx->nnz / (double)n_im : 0;
    const uint64_t *rs = ctx->c_rowstart.as<uint64_t>();
    const uint32_t *rl = ctx->c_rowlen.as<uint32_t>(), *cl = ctx->c_col.as<uint32_t>(), *idf = local_out ? nullptr : ctx->crow_dofs->as<uint32_t>();
    const double *vl = ctx->c_val.as<double>();
    if (n_im) {
      if (avg > 20) spmv_crows<4, 8><<<nm_blocks(n_im * 4, 256), 256, 0, NM_LS(ctx)>>>(n_im, rs, rl, cl, vl, idf, x, y);
      else if (avg > 6) spmv_crows<4, 2><<<nm_blocks(n_im * 4, 256), 256, 0, NM_LS(ctx)>>>(n_im, rs, rl, cl, vl, idf, x, y);
      else spmv_crows<2, 2><<<nm_blocks(n_im * 2, 256), 256, 0, NM_LS(ctx)>>>(n_im, rs, rl, cl, vl, idf, x, y);
    }
  }
  NM_CUDA(cudaGetLastError());
  tm.stop();
  return NM_OK;
}

int nm_spmv_push_run(nm_ctx *ctx, const double *x, int n_targets, double *const *targets, int multicast)
{
  const uint64_t n_rows = ctx->n_local_rows;
  const uint32_t *ids = ctx->compact_rows ? ctx->row_ids.as<uint32_t>() : nullptr;
  PushTargets T;
  T.n = n_targets;
  for (int i = 0; i < 16; ++i) T.p[i] = i < n_targets ? targets[i] : nullptr;
  PhaseTimer tm(ctx, NM_T_SPMV);
  const uint32_t *cl = ctx->a_col.as<uint32_t>();
  const double *vl = ctx->a_val.as<double>();
  if (n_rows) {
    const unsigned B = nm_blocks(n_rows, 128);
    if (ctx->a_rowptr32_valid) {
      if (multicast) spmv_csr_push<8, uint32_t, true><<<B, 128, 0, NM_LS(ctx)>>>(n_rows, ctx->a_rowptr32.as<uint32_t>(), cl, vl, x, T, ids);
      else spmv_csr_push<8, uint32_t, false><<<B, 128, 0, NM_LS(ctx)>>>(n_rows, ctx->a_rowptr32.as<uint32_t>(), cl, vl, x, T, ids);
    } else {
      if (multicast) spmv_csr_push<8, uint64_t, true><<<B, 128, 0, NM_LS(ctx)>>>(n_rows, ctx->a_rowptr.as<uint64_t>(), cl, vl, x, T, ids);
      else spmv_csr_push<8, uint64_t, false><<<B, 128, 0, NM_LS(ctx)>>>(n_rows, ctx->a_rowptr.as<uint64_t>(), cl, vl, x, T, ids);
    }
  }
  NM_CUDA(cudaGetLastError());
  tm.stop();
  return NM_OK;
}

int nm_spmv_push_owned_run(nm_ctx *ctx, const double *x, int n_ranks, double *const *targets, uint64_t rows_per_rank)
{
  const uint64_t n_rows = ctx->n_local_rows;
  const uint32_t *ids = ctx->compact_rows ? ctx->row_ids.as<uint32_t>() : nullptr;
  PushTargets T;
  T.n = n_ranks;
  for (int i = 0; i < 16; ++i) T.p[i] = i < n_ranks ? targets[i] : nullptr;
  PhaseTimer tm(ctx, NM_T_SPMV);
  if (n_rows) {
    const unsigned B = nm_blocks(n_rows, 128);
    if (ctx->a_rowptr32_valid)
      spmv_csr_push_owned<8, uint32_t><<<B, 128, 0, NM_LS(ctx)>>>(n_rows, ctx->a_rowptr32.as<uint32_t>(), ctx->a_col.as<uint32_t>(), ctx->a_val.as<double>(), x, T, ids, rows_per_rank);
    else
      spmv_csr_push_owned<8, uint64_t><<<B, 128, 0, NM_LS(ctx)>>>(n_rows, ctx->a_rowptr.as<uint64_t>(), ctx->a_col.as<uint32_t>(), ctx->a_val.as<double>(), x, T, ids, rows_per_rank);
  }
  NM_CUDA(cudaGetLastError());
  tm.stop();
  return NM_OK;
}

// compact CSR with rows = immersed dofs (device buffers owned by the caller of this helper)
int nm_c_rows_compact(nm_ctx *ctx, uint64_t *rowptr_dev, uint32_t *col_dev, double *val_dev)
{
  const uint64_t n_im = ctx->n_crows, n_cols = ctx->n_im_dofs;
  NM_TRY(nm_ensure(ctx, ctx->scratch_a, (n_cols + 2) * sizeof(uint32_t)));
  NM_CUDA(cudaMemsetAsync(ctx->scratch_a.p, 0, (n_cols + 2) * sizeof(uint32_t), ctx->stream));
  if (n_im) scatter_len_by_dof<<<nm_blocks(n_im, 256), 256, 0, NM_LS(ctx)>>>(n_im, ctx->c_rowlen.as<uint32_t>(), ctx->crow_dofs->as<uint32_t>(), ctx->scratch_a.as<uint32_t>());
  NM_TRY(nm_exclusive_scan_u32_to_u64(ctx, ctx->scratch_a.as<uint32_t>(), rowptr_dev, n_cols + 1));
  if (n_im && col_dev)
    c_compact<8><<<nm_blocks(n_im * 8, 256), 256, 0, NM_LS(ctx)>>>(n_im, ctx->c_rowstart.as<uint64_t>(), ctx->c_rowlen.as<uint32_t>(), ctx->c_col.as<uint32_t>(),
                                                                 ctx->c_val.as<double>(), ctx->crow_dofs->as<uint32_t>(), rowptr_dev, col_dev, val_dev);
  NM_CUDA(cudaGetLastError());
  return NM_OK;
}

// ---------------------------------------------------------------------------------------
// compact read-back of the stored orientation and vectors in local numbering
// ---------------------------------------------------------------------------------------
// List of the rows that hold entries: ctx->row_ids (compact-row mode: already there) or built here from the row
// offsets (full-row mode) into ctx->out_ids / ctx->out_len.  Returns device pointers valid until the next matrix build.
int nm_nonempty_rows(nm_ctx *ctx, uint64_t *n_rows_out, const uint32_t **ids_dev, const uint32_t **len_dev)
{
  const uint64_t n_rows = ctx->n_local_rows;
  if (!ctx->out_rows_valid) {
    uint64_t n_out = n_rows;
    const uint64_t *pos = nullptr;
    if (!ctx->compact_rows) {
      NM_TRY(nm_ensure(ctx, ctx->scratch_a, (n_rows + 2) * sizeof(uint32_t)));
      NM_TRY(nm_ensure(ctx, ctx->scratch_b, (n_rows + 2) * sizeof(uint64_t)));
      nonempty_flags<<<nm_blocks(n_rows + 1, 256), 256, 0, NM_LS(ctx)>>>(n_rows, ctx->a_rowptr.as<uint64_t>(), ctx->scratch_a.as<uint32_t>());
      NM_CUDA(cudaGetLastError());
      NM_TRY(nm_exclusive_scan_u32_to_u64(ctx, ctx->scratch_a.as<uint32_t>(), ctx->scratch_b.as<uint64_t>(), n_rows + 1));
      NM_CUDA(cudaMemcpyAsync(&n_out, ctx->scratch_b.as<uint64_t>() + n_rows, 8, cudaMemcpyDeviceToHost, ctx->stream));
      NM_SYNC(ctx);
      pos = ctx->scratch_b.as<uint64_t>();
    }
    NM_TRY(nm_ensure(ctx, ctx->out_ids, (n_out + 1) * sizeof(uint32_t)));
    NM_TRY(nm_ensure(ctx, ctx->out_len, (n_out + 1) * sizeof(uint32_t)));
    if (n_rows) {
      compact_rows_out<<<nm_blocks(n_rows, 256), 256, 0, NM_LS(ctx)>>>(n_rows, ctx->a_rowptr.as<uint64_t>(), ctx->row_ids.as<uint32_t>(), pos,
                                                                     ctx->out_ids.as<uint32_t>(), ctx->out_len.as<uint32_t>());
      NM_CUDA(cudaGetLastError());
    }
    ctx->n_out_rows = n_out;
    ctx->out_rows_valid = true;
  }
  if (n_rows_out) *n_rows_out = ctx->n_out_rows;
  if (ids_dev) *ids_dev = ctx->out_ids.as<uint32_t>();
  if (len_dev) *len_dev = ctx->out_len.as<uint32_t>();
  return NM_OK;
}

// which = NM_LOCAL_IMMERSED: local entry m <-> global immersed dof im_dofs[m];  NM_LOCAL_ROWS: local entry r <-> dof of the
// r-th non-empty row.  Device pointers.
int nm_local_map(nm_ctx *ctx, int which, int to_global, const double *src, double *dst)
{
  uint64_t n = 0;
  const uint32_t *idx = nullptr;
  if (which == NM_LOCAL_IMMERSED) { n = ctx->n_crows; idx = ctx->crow_dofs->as<uint32_t>(); }
  else NM_TRY(nm_nonempty_rows(ctx, &n, &idx, nullptr));
  if (!n) return NM_OK;
  if (to_global) scatter_by_index<<<nm_blocks(n, 256), 256, 0, NM_LS(ctx)>>>(n, idx, src, dst);
  else gather_by_index<<<nm_blocks(n, 256), 256, 0, NM_LS(ctx)>>>(n, idx, src, dst);
  NM_CUDA(cudaGetLastError());
  return NM_OK;
}
