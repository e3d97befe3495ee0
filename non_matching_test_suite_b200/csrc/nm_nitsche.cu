// Interface-penalisation (Nitsche) terms on the quadrature stream of the intersection path.
//
// Replaces, on the caller's quadrature, the arithmetic of
//   assemble_nitsche_with_exact_intersections     code/include/coupling_utilities.h:306-398
//   create_nitsche_rhs_with_exact_intersections   code/include/coupling_utilities.h:400-481
// of the reference:  for every (background cell, -, quadrature) tuple
//   local(i,j) += c(x_q) * (penalty / h) * phi_i(xi_q) * phi_j(xi_q) * w_q          (:366-370)
//   rhs(i)     += c(x_q) * (penalty / h) * phi_i(xi_q) * f(x_q) * w_q               (:447-451)
// with h = cell->diameter() (:357,:429), xi_q = F^-1(x_q) on the axis-aligned Q1 cell, followed by
// AffineConstraints::distribute_local_to_global (:387-388, :458-459).  The Function objects of the reference are
// evaluated by the host wrapper (value_list, :349,:438-441) and arrive here as per-point values.
//
// Data flow: the tuples are grouped by background cell (stable radix sort of the cell ids); one thread per CUT CELL
// walks its tuples in input order, accumulates the symmetric local block in registers, resolves rows and columns
// through the constraint lines and writes (row<<32|col, value) records into a slot range obtained from a scan of exact
// counts; a stable radix sort on the key followed by a sequential sum over each run gives a CSR whose values do not
// depend on scheduling.  Summing the blocks of one cell before distributing them is the same bilinear operation as the
// reference's per-tuple distribute_local_to_global; only the diagonal entry of a constrained dof is not linear
// (|local(i,i)| per tuple, or the tuple's mean |diagonal| when that is zero) and is therefore tracked per tuple.
#include <cub/device/device_radix_sort.cuh>

#include "nm_common.cuh"

namespace {

constexpr int NIT_THREADS = 128;

// number of rows a local dof resolves to: itself, its masters, or nothing (constrained without masters)
__device__ __forceinline__ uint32_t n_resolved(uint32_t dof, const int32_t *__restrict__ line_of, const uint64_t *__restrict__ c_ptr,
                                               bool &constrained)
{
  const int32_t ln = line_of ? line_of[dof] : -1;
  constrained = ln >= 0;
  return ln < 0 ? 1u : (uint32_t)(c_ptr[ln + 1] - c_ptr[ln]);
}

__global__ void iota_kernel(uint64_t n, uint32_t *__restrict__ out)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)i;
}

// start[u] = first position of run u in the sorted cell ids; start[n_runs] = n
__global__ void run_starts_kernel(uint64_t n, const uint64_t *__restrict__ head, const uint64_t *__restrict__ upos, uint64_t *__restrict__ start)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && head[i]) start[upos[i]] = i;
  if (i == n) start[upos[n]] = n;
}

template <int D>
__global__ void nitsche_count_kernel(uint64_t n_cut, uint64_t n_bg, const uint32_t *__restrict__ sorted_bg, const uint64_t *__restrict__ start,
                                     const uint32_t *__restrict__ bg_dofs, const int32_t *__restrict__ line_of,
                                     const uint64_t *__restrict__ c_ptr, uint64_t *__restrict__ cnt_m, uint64_t *__restrict__ cnt_r,
                                     unsigned long long *n_bad)
{
  constexpr int NL = 1 << D;
  const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u > n_cut) return;
  if (u == n_cut) { cnt_m[u] = 0; cnt_r[u] = 0; return; }
  const uint32_t b = sorted_bg[start[u]];
  if (b >= n_bg) { atomicAdd(n_bad, (unsigned long long)(start[u + 1] - start[u])); cnt_m[u] = 0; cnt_r[u] = 0; return; }
  uint32_t nres = 0, ncon = 0;
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    bool con;
    nres += n_resolved(bg_dofs[(uint64_t)b * NL + i], line_of, c_ptr, con);
    ncon += con;
  }
  cnt_m[u] = (uint64_t)nres * nres + ncon;
  cnt_r[u] = nres;
}

// index of (i,j), i <= j, in the packed upper triangle
__host__ __device__ constexpr int sym_idx(int i, int j, int NL) { return i * NL - i * (i - 1) / 2 + (j - i); }

template <int D>
__global__ void __launch_bounds__(NIT_THREADS)
nitsche_fill_kernel(uint64_t n_cut, uint64_t n_bg, const uint32_t *__restrict__ sorted_bg, const uint32_t *__restrict__ perm,
                    const uint64_t *__restrict__ start, const uint64_t *__restrict__ q_off,
                    const double *__restrict__ pts, const double *__restrict__ wts, const double *__restrict__ coeff,
                    const double *__restrict__ fval, double penalty, const double *__restrict__ bg_box,
                    const uint32_t *__restrict__ bg_dofs, const int32_t *__restrict__ line_of, const uint64_t *__restrict__ c_ptr,
                    const uint32_t *__restrict__ c_master, const double *__restrict__ c_weight,
                    const uint64_t *__restrict__ off_m, const uint64_t *__restrict__ off_r, int want_matrix, int cb,
                    unsigned long long *__restrict__ keys, double *__restrict__ vals, uint32_t *__restrict__ rkeys,
                    double *__restrict__ rvals)
{
  constexpr int NL = 1 << D, NS = NL * (NL + 1) / 2;
  __shared__ double s_loc[NS + NL][NIT_THREADS];
  const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_cut) return;
  const uint32_t b = sorted_bg[start[u]];
  if (b >= n_bg) return;
  const double *box = bg_box + (uint64_t)b * 8;
  double ih[D], h2 = 0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const double H = box[4 + k] - box[k];
    ih[k] = 1.0 / H;
    h2 += H * H;
  }
  const double scale = penalty / sqrt(h2);  // penalty / cell->diameter()
  uint32_t dof[NL];
  int32_t line[NL];
  bool any_con = false;
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    dof[i] = bg_dofs[(uint64_t)b * NL + i];
    line[i] = line_of ? line_of[dof[i]] : -1;
    any_con |= line[i] >= 0;
  }
  double acc[NS], racc[NL], fix[NL];  // fix: diagonal additions of constrained dofs, summed over the tuples
#pragma unroll
  for (int i = 0; i < NS; ++i) acc[i] = 0;
#pragma unroll
  for (int i = 0; i < NL; ++i) { racc[i] = 0; fix[i] = 0; }
  for (uint64_t t = start[u]; t < start[u + 1]; ++t) {
   const uint32_t p = perm[t];
   double dg[NL];  // this tuple's diagonal
#pragma unroll
   for (int i = 0; i < NL; ++i) dg[i] = 0;
   for (uint64_t q = q_off[p]; q < q_off[p + 1]; ++q) {
    double xi[D], phi[NL];
#pragma unroll
    for (int k = 0; k < D; ++k) xi[k] = (pts[q * D + k] - box[k]) * ih[k];
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      double v = 1.0;
#pragma unroll
      for (int k = 0; k < D; ++k) v *= ((i >> k) & 1) ? xi[k] : 1.0 - xi[k];
      phi[i] = v;
    }
    const double cq = (coeff ? coeff[q] : 1.0) * scale, w = wts[q];
    if (want_matrix) {
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        dg[i] += cq * phi[i] * phi[i] * w;
#pragma unroll
        for (int j = i + 1; j < NL; ++j) acc[sym_idx(i, j, NL)] += cq * phi[i] * phi[j] * w;
      }
    }
    if (fval) {
      const double f = fval[q];
#pragma unroll
      for (int i = 0; i < NL; ++i) racc[i] += cq * phi[i] * f * w;
    }
   }
   if (want_matrix) {
     if (any_con) {  // set_matrix_diagonals of this tuple: |local(i,i)|, or the mean |diagonal| of the block if zero
       double avg = 0;
#pragma unroll
       for (int i = 0; i < NL; ++i) avg += fabs(dg[i]);
       avg /= NL;
#pragma unroll
       for (int i = 0; i < NL; ++i) fix[i] += fabs(dg[i]) != 0.0 ? fabs(dg[i]) : avg;
     }
#pragma unroll
     for (int i = 0; i < NL; ++i) acc[sym_idx(i, i, NL)] += dg[i];
   }
  }
#pragma unroll
  for (int i = 0; i < NS; ++i) s_loc[i][threadIdx.x] = acc[i];
#pragma unroll
  for (int i = 0; i < NL; ++i) s_loc[NS + i][threadIdx.x] = racc[i];

  if (want_matrix) {
    uint64_t pos = off_m[u];
    for (int i = 0; i < NL; ++i) {
      const uint64_t ib = line[i] < 0 ? 0 : c_ptr[line[i]], ie = line[i] < 0 ? 1 : c_ptr[line[i] + 1];
      for (uint64_t a = ib; a < ie; ++a) {
        const uint32_t r = line[i] < 0 ? dof[i] : c_master[a];
        const double wr = line[i] < 0 ? 1.0 : c_weight[a];
        for (int j = 0; j < NL; ++j) {
          const double lij = s_loc[i <= j ? sym_idx(i, j, NL) : sym_idx(j, i, NL)][threadIdx.x];
          const uint64_t jb = line[j] < 0 ? 0 : c_ptr[line[j]], je = line[j] < 0 ? 1 : c_ptr[line[j] + 1];
          for (uint64_t c = jb; c < je; ++c) {
            const uint32_t cc = line[j] < 0 ? dof[j] : c_master[c];
            const double wc = line[j] < 0 ? 1.0 : c_weight[c];
            keys[pos] = ((unsigned long long)r << cb) | cc;
            vals[pos] = wr * wc * lij;
            ++pos;
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NL; ++i)
      if (line[i] >= 0) {
        keys[pos] = ((unsigned long long)dof[i] << cb) | dof[i];
        vals[pos] = fix[i];
        ++pos;
      }
  }
  if (fval) {
    uint64_t pos = off_r[u];
    for (int i = 0; i < NL; ++i) {
      const double v = s_loc[NS + i][threadIdx.x];
      const uint64_t ib = line[i] < 0 ? 0 : c_ptr[line[i]], ie = line[i] < 0 ? 1 : c_ptr[line[i] + 1];
      for (uint64_t a = ib; a < ie; ++a) {
        rkeys[pos] = line[i] < 0 ? dof[i] : c_master[a];
        rvals[pos] = (line[i] < 0 ? 1.0 : c_weight[a]) * v;
        ++pos;
      }
    }
  }
}

template <typename K>
__global__ void run_heads_kernel(uint64_t n, const K *__restrict__ k, uint64_t *__restrict__ head)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || k[i] != k[i - 1]) ? 1 : 0;
  if (i == n) head[i] = 0;
}

// one thread per run of equal keys: sequential sum in record order
__global__ void run_sum_matrix_kernel(uint64_t n, const unsigned long long *__restrict__ k, const double *__restrict__ v,
                                      const uint64_t *__restrict__ upos, int cb, unsigned long long *__restrict__ ukey,
                                      uint32_t *__restrict__ col, double *__restrict__ val)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = k[i];
  if (i != 0 && k[i - 1] == key) return;
  double s = 0;
  for (uint64_t j = i; j < n && k[j] == key; ++j) s += v[j];
  const uint64_t u = upos[i];
  ukey[u] = key;
  col[u] = (uint32_t)(key & ((1ull << cb) - 1));
  val[u] = s;
}

__global__ void run_sum_rhs_kernel(uint64_t n, const uint32_t *__restrict__ k, const double *__restrict__ v, double *__restrict__ rhs)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t key = k[i];
  if (i != 0 && k[i - 1] == key) return;
  double s = 0;
  for (uint64_t j = i; j < n && k[j] == key; ++j) s += v[j];
  rhs[key] = s;
}

__global__ void rowptr_from_keys_kernel(uint64_t n_rows, uint64_t n_unique, int cb, const unsigned long long *__restrict__ ukey,
                                        uint64_t *__restrict__ rowptr)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  const unsigned long long target = (unsigned long long)r << cb;
  uint64_t lo = 0, hi = n_unique;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (ukey[mid] < target) lo = mid + 1; else hi = mid;
  }
  rowptr[r] = lo;
}

int bits_for(uint64_t n)
{
  int b = 1;
  while (b < 64 && (n >> b)) ++b;
  return b;
}

}  // namespace

int nm_nitsche_run(nm_ctx *ctx, uint64_t n_pairs, const uint32_t *bg, const uint64_t *q_offset, const double *points,
                   const double *weights, const double *coefficient, const double *rhs_values, double penalty, int what, int where)
{
  const int d = ctx->spacedim;
  const bool want_m = what & NM_NITSCHE_MATRIX, want_r = (what & NM_NITSCHE_RHS) != 0;
  const uint32_t *bg_d = bg;
  const uint64_t *qo_d = q_offset;
  const double *p_d = points, *w_d = weights, *c_d = coefficient, *f_d = rhs_values;
  if (where == NM_HOST && n_pairs) {
    const uint64_t nq = q_offset[n_pairs];
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t b0 = n_pairs * 4, b1 = (n_pairs + 1) * 8, b2 = nq * d * 8, b3 = nq * 8, b4 = coefficient ? nq * 8 : 0,
                 b5 = (want_r && rhs_values) ? nq * 8 : 0;
    NM_TRY(nm_ensure(ctx, ctx->h2d_stage, al(b0) + al(b1) + al(b2) + al(b3) + al(b4) + al(b5) + 256));
    char *p0 = ctx->h2d_stage.as<char>(), *p1 = p0 + al(b0), *p2 = p1 + al(b1), *p3 = p2 + al(b2), *p4 = p3 + al(b3), *p5 = p4 + al(b4);
    NM_CUDA(cudaMemcpyAsync(p0, bg, b0, cudaMemcpyHostToDevice, ctx->stream));
    NM_CUDA(cudaMemcpyAsync(p1, q_offset, b1, cudaMemcpyHostToDevice, ctx->stream));
    NM_CUDA(cudaMemcpyAsync(p2, points, b2, cudaMemcpyHostToDevice, ctx->stream));
    NM_CUDA(cudaMemcpyAsync(p3, weights, b3, cudaMemcpyHostToDevice, ctx->stream));
    if (b4) NM_CUDA(cudaMemcpyAsync(p4, coefficient, b4, cudaMemcpyHostToDevice, ctx->stream));
    if (b5) NM_CUDA(cudaMemcpyAsync(p5, rhs_values, b5, cudaMemcpyHostToDevice, ctx->stream));
    bg_d = (const uint32_t *)p0; qo_d = (const uint64_t *)p1; p_d = (const double *)p2; w_d = (const double *)p3;
    c_d = b4 ? (const double *)p4 : nullptr;
    f_d = b5 ? (const double *)p5 : nullptr;
  }
  if (!want_r) f_d = nullptr;
  const uint64_t n_rows = ctx->n_space_dofs;
  const int cb = bits_for(n_rows);
  const int32_t *line_of = ctx->n_constrained ? ctx->line_of.as<int32_t>() : nullptr;

  // ---- group the tuples by background cell: stable sort of (cell id, tuple index) ----
  const uint64_t n = n_pairs;
  NM_TRY(nm_ensure(ctx, ctx->nit_cnt, 4 * (n + 2) * 4 + 7 * (n + 2) * 8 + 256));
  uint32_t *id_in = ctx->nit_cnt.as<uint32_t>(), *id_sorted = id_in + (n + 2), *perm_in = id_sorted + (n + 2), *perm = perm_in + (n + 2);
  uint64_t *head = reinterpret_cast<uint64_t *>(perm + (n + 2)), *upos = head + (n + 2), *start = upos + (n + 2),
           *cnt_m = start + (n + 2), *cnt_r = cnt_m + (n + 2), *off_m = cnt_r + (n + 2), *off_r = off_m + (n + 2);
  unsigned long long *n_bad = ctx->counters.as<unsigned long long>() + 9;
  NM_CUDA(cudaMemsetAsync(n_bad, 0, 8, ctx->stream));
  uint64_t n_cut = 0;
  if (n) {
    if (n >= (1ull << 32)) return nm_fail(ctx, NM_ERR_UNSUPPORTED, "more than 2^32 tuples in one call");
    iota_kernel<<<nm_blocks(n, 256), 256, 0, NM_LS(ctx)>>>(n, perm_in);
    NM_CUDA(cudaGetLastError());
    // only the bits of a valid id take part; an out-of-range id may land anywhere, runs are cut on the full id and
    // any run of an invalid id fails the call below
    size_t tb = 0;
    const int id_bits = bits_for(ctx->n_bg);
    cub::DeviceRadixSort::SortPairs(nullptr, tb, bg_d, id_sorted, perm_in, perm, n, 0, id_bits, ctx->stream);
    NM_TRY(nm_ensure(ctx, ctx->scan_tmp, tb + 256));
    NM_CUDA(cub::DeviceRadixSort::SortPairs(ctx->scan_tmp.p, tb, bg_d, id_sorted, perm_in, perm, n, 0, id_bits, ctx->stream));
    run_heads_kernel<uint32_t><<<nm_blocks(n + 1, 256), 256, 0, NM_LS(ctx)>>>(n, id_sorted, head);
    NM_CUDA(cudaGetLastError());
    NM_TRY(nm_exclusive_scan_u64(ctx, head, upos, n + 1));
    run_starts_kernel<<<nm_blocks(n + 1, 256), 256, 0, NM_LS(ctx)>>>(n, head, upos, start);
    NM_CUDA(cudaGetLastError());
    NM_CUDA(cudaMemcpyAsync(&n_cut, upos + n, 8, cudaMemcpyDeviceToHost, ctx->stream));
    NM_SYNC(ctx);
  }
  (void)id_in;

  // ---- exact record counts per cut cell -> slot ranges ----
  if (d == 3)
    nitsche_count_kernel<3><<<nm_blocks(n_cut + 1, 256), 256, 0, NM_LS(ctx)>>>(n_cut, ctx->n_bg, id_sorted, start, ctx->bg_dofs.as<uint32_t>(),
                                                                               line_of, ctx->c_ptr.as<uint64_t>(), cnt_m, cnt_r, n_bad);
  else
    nitsche_count_kernel<2><<<nm_blocks(n_cut + 1, 256), 256, 0, NM_LS(ctx)>>>(n_cut, ctx->n_bg, id_sorted, start, ctx->bg_dofs.as<uint32_t>(),
                                                                               line_of, ctx->c_ptr.as<uint64_t>(), cnt_m, cnt_r, n_bad);
  NM_CUDA(cudaGetLastError());
  NM_TRY(nm_exclusive_scan_u64(ctx, cnt_m, off_m, n_cut + 1));
  NM_TRY(nm_exclusive_scan_u64(ctx, cnt_r, off_r, n_cut + 1));
  uint64_t tot[2] = {0, 0};
  unsigned long long bad = 0;
  NM_CUDA(cudaMemcpyAsync(&tot[0], off_m + n_cut, 8, cudaMemcpyDeviceToHost, ctx->stream));
  NM_CUDA(cudaMemcpyAsync(&tot[1], off_r + n_cut, 8, cudaMemcpyDeviceToHost, ctx->stream));
  NM_CUDA(cudaMemcpyAsync(&bad, n_bad, 8, cudaMemcpyDeviceToHost, ctx->stream));
  NM_SYNC(ctx);
  if (bad) return nm_fail(ctx, NM_ERR_INVALID, "%llu background cell ids are out of range", bad);
  const uint64_t n_m = want_m ? tot[0] : 0, n_r = f_d ? tot[1] : 0;

  NM_TRY(nm_ensure(ctx, ctx->nit_keys, 2 * (n_m + 1) * 8));
  NM_TRY(nm_ensure(ctx, ctx->nit_vals, 2 * (n_m + 1) * 8));
  NM_TRY(nm_ensure(ctx, ctx->nit_rkeys, 2 * (n_r + 1) * 4));
  NM_TRY(nm_ensure(ctx, ctx->nit_rvals, 2 * (n_r + 1) * 8));
  unsigned long long *k0 = ctx->nit_keys.as<unsigned long long>(), *k1 = k0 + (n_m + 1);
  double *v0 = ctx->nit_vals.as<double>(), *v1 = v0 + (n_m + 1);
  uint32_t *rk0 = ctx->nit_rkeys.as<uint32_t>(), *rk1 = rk0 + (n_r + 1);
  double *rv0 = ctx->nit_rvals.as<double>(), *rv1 = rv0 + (n_r + 1);
  if (n_cut) {
    if (d == 3)
      nitsche_fill_kernel<3><<<nm_blocks(n_cut, NIT_THREADS), NIT_THREADS, 0, NM_LS(ctx)>>>(
          n_cut, ctx->n_bg, id_sorted, perm, start, qo_d, p_d, w_d, c_d, f_d, penalty, ctx->bg_box.as<double>(), ctx->bg_dofs.as<uint32_t>(),
          line_of, ctx->c_ptr.as<uint64_t>(), ctx->c_master.as<uint32_t>(), ctx->c_weight.as<double>(), off_m, off_r, want_m ? 1 : 0, cb, k0, v0,
          rk0, rv0);
    else
      nitsche_fill_kernel<2><<<nm_blocks(n_cut, NIT_THREADS), NIT_THREADS, 0, NM_LS(ctx)>>>(
          n_cut, ctx->n_bg, id_sorted, perm, start, qo_d, p_d, w_d, c_d, f_d, penalty, ctx->bg_box.as<double>(), ctx->bg_dofs.as<uint32_t>(),
          line_of, ctx->c_ptr.as<uint64_t>(), ctx->c_master.as<uint32_t>(), ctx->c_weight.as<double>(), off_m, off_r, want_m ? 1 : 0, cb, k0, v0,
          rk0, rv0);
    NM_CUDA(cudaGetLastError());
  }

  // ---- matrix: stable sort by (row, col), sum the runs, CSR ----
  ctx->nit_nnz = 0;
  NM_TRY(nm_ensure(ctx, ctx->nit_rowptr, (n_rows + 2) * 8));
  if (want_m) {
    if (n_m) {
      size_t tb = 0;
      const int end_bit = 2 * cb;  // keys are (row << cb) | col with cb = bits of the largest dof index
      cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, v0, v1, n_m, 0, end_bit, ctx->stream);
      NM_TRY(nm_ensure(ctx, ctx->scan_tmp, tb + 256));
      NM_CUDA(cub::DeviceRadixSort::SortPairs(ctx->scan_tmp.p, tb, k0, k1, v0, v1, n_m, 0, end_bit, ctx->stream));
      // run heads -> position of every distinct key (k0 is free now and serves as scratch)
      NM_TRY(nm_ensure(ctx, ctx->scratch_a, 2 * (n_m + 2) * 8));
      uint64_t *mhead = ctx->scratch_a.as<uint64_t>(), *mupos = mhead + (n_m + 2);
      run_heads_kernel<unsigned long long><<<nm_blocks(n_m + 1, 256), 256, 0, NM_LS(ctx)>>>(n_m, k1, mhead);
      NM_CUDA(cudaGetLastError());
      NM_TRY(nm_exclusive_scan_u64(ctx, mhead, mupos, n_m + 1));
      uint64_t nu = 0;
      NM_CUDA(cudaMemcpyAsync(&nu, mupos + n_m, 8, cudaMemcpyDeviceToHost, ctx->stream));
      NM_SYNC(ctx);
      ctx->nit_nnz = nu;
      NM_TRY(nm_ensure(ctx, ctx->nit_col, (nu + 1) * 4));
      NM_TRY(nm_ensure(ctx, ctx->nit_val, (nu + 1) * 8));
      run_sum_matrix_kernel<<<nm_blocks(n_m, 256), 256, 0, NM_LS(ctx)>>>(n_m, k1, v1, mupos, cb, k0, ctx->nit_col.as<uint32_t>(), ctx->nit_val.as<double>());
      NM_CUDA(cudaGetLastError());
      rowptr_from_keys_kernel<<<nm_blocks(n_rows + 1, 256), 256, 0, NM_LS(ctx)>>>(n_rows, nu, cb, k0, ctx->nit_rowptr.as<uint64_t>());
      NM_CUDA(cudaGetLastError());
    } else {
      NM_CUDA(cudaMemsetAsync(ctx->nit_rowptr.p, 0, (n_rows + 1) * 8, ctx->stream));
    }
  }
  // ---- right-hand side ----
  if (want_r) {
    NM_TRY(nm_ensure(ctx, ctx->nit_rhs, (n_rows + 1) * 8));
    NM_CUDA(cudaMemsetAsync(ctx->nit_rhs.p, 0, (n_rows + 1) * 8, ctx->stream));
    if (n_r) {
      size_t tb = 0;
      const int end_bit = bits_for(n_rows) > 32 ? 32 : bits_for(n_rows);
      cub::DeviceRadixSort::SortPairs(nullptr, tb, rk0, rk1, rv0, rv1, n_r, 0, end_bit, ctx->stream);
      NM_TRY(nm_ensure(ctx, ctx->scan_tmp, tb + 256));
      NM_CUDA(cub::DeviceRadixSort::SortPairs(ctx->scan_tmp.p, tb, rk0, rk1, rv0, rv1, n_r, 0, end_bit, ctx->stream));
      run_sum_rhs_kernel<<<nm_blocks(n_r, 256), 256, 0, NM_LS(ctx)>>>(n_r, rk1, rv1, ctx->nit_rhs.as<double>());
      NM_CUDA(cudaGetLastError());
    }
  }
  NM_SYNC(ctx);
  ctx->nit_has_matrix = want_m;
  ctx->nit_has_rhs = want_r;
  return NM_OK;
}
