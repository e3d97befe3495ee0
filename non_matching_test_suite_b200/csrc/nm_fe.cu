// General tensor-product Lagrange elements on the coupling path: FE_Q(p) on the background,
// FE_DGQ(q) on the immersed grid.
//
// Replaces, for any (p, q), the per-pair loop of create_coupling_mass_matrix_with_exact_intersections
//   xi_q  = F_bg^-1(x_q)                         code/include/coupling_utilities.h:248-249
//   eta_q = F_im^-1(x_q)   (codimension one)     :251-254
//   local(i, j) += phi_i(xi_q) psi_j(eta_q) w_q  :258-274
//   distribute_local_to_global through the space constraints, compress(add)   :276-289
// and the pattern of create_coupling_sparsity_pattern_with_exact_intersections (:100-139, keep_constrained_entries).
// The shape functions are the Lagrange polynomials of the element's own unit support points, which the caller hands
// over in the element's dof order (fe.get_unit_support_points()): no numbering convention of deal.II is assumed.
//
// Device pipeline: one thread per accepted pair evaluates its quadrature points and writes (row of C, space dof, value)
// records with the constraint lines expanded; a stable radix sort groups them, runs of equal keys are summed in order
// (bit-reproducible), the rows of C come out sorted, and the transpose of nm_matrix.cu gives the stored orientation.
#include <cub/cub.cuh>

#include <vector>

#include "nm_common.cuh"

int nm_accepted_index(nm_ctx *ctx, uint64_t **idx_dev);
int nm_constraints_set(nm_ctx *ctx, uint64_t n_dofs, uint64_t n_constrained, const uint32_t *c_dof, const uint64_t *c_ptr, const uint32_t *c_master,
                       const double *c_weight, int where);

namespace {

constexpr int FE_MAXN = 6;     // nodes per direction (degree <= 5)
constexpr int FE_MAX_BG = 64;  // background dofs per cell: Q3 in 3D, Q5 in 2D (36)
constexpr int FE_MAX_IM = 16;  // immersed dofs per cell: Q3 quads, Q5 segments

struct FeTable {
  int n1;
  double node[FE_MAXN];
  double inv_den[FE_MAXN];  // 1 / prod_{b != a} (x_a - x_b)
};
struct FeConst {
  FeTable bg, im;
  int n_bg, n_im;
  unsigned char bg_idx[FE_MAX_BG][3];  // dof -> node index per direction
  unsigned char im_idx[FE_MAX_IM][2];
};
__constant__ FeConst c_fe;

__device__ inline void lagrange_all(const FeTable &T, double x, double *L)
{
  for (int a = 0; a < T.n1; ++a) {
    double p = T.inv_den[a];
    for (int b = 0; b < T.n1; ++b)
      if (b != a) p *= x - T.node[b];
    L[a] = p;
  }
}

// eta = F_im^-1(x): the point of the immersed cell closest to x in unit coordinates (what Mapping::transform_real_to_unit_cell
// returns for a codimension-one cell, coupling_utilities.h:251-254).  Segment: orthogonal projection.  Bilinear quad: Newton
// on grad |F(eta) - x|^2 = 0 with the full Hessian (the only second derivative of a bilinear map is the mixed one).
template <int D>
__device__ inline void immersed_unit_coords(const double *v, const double *x, double *eta)
{
  if (D == 2) {
    const double ex = v[2] - v[0], ey = v[3] - v[1];
    eta[0] = ((x[0] - v[0]) * ex + (x[1] - v[1]) * ey) / (ex * ex + ey * ey);
    eta[1] = 0.0;
    return;
  }
  const double *v0 = v, *v1 = v + 3, *v2 = v + 6, *v3 = v + 9;
  double a[3], b[3], c[3];   // F = v0 + a s + b t + c s t
#pragma unroll
  for (int k = 0; k < 3; ++k) { a[k] = v1[k] - v0[k]; b[k] = v2[k] - v0[k]; c[k] = v0[k] - v1[k] - v2[k] + v3[k]; }
  double s = 0.5, t = 0.5;
  for (int it = 0; it < 30; ++it) {
    double r[3], Js[3], Jt[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      r[k] = v0[k] + a[k] * s + b[k] * t + c[k] * s * t - x[k];
      Js[k] = a[k] + c[k] * t;
      Jt[k] = b[k] + c[k] * s;
    }
    const double gs = Js[0] * r[0] + Js[1] * r[1] + Js[2] * r[2], gt = Jt[0] * r[0] + Jt[1] * r[1] + Jt[2] * r[2];
    const double rc = r[0] * c[0] + r[1] * c[1] + r[2] * c[2];
    const double hss = Js[0] * Js[0] + Js[1] * Js[1] + Js[2] * Js[2], htt = Jt[0] * Jt[0] + Jt[1] * Jt[1] + Jt[2] * Jt[2];
    const double hst = Js[0] * Jt[0] + Js[1] * Jt[1] + Js[2] * Jt[2] + rc;
    const double det = hss * htt - hst * hst;
    if (!(fabs(det) > 0.0)) break;
    const double ds = (htt * gs - hst * gt) / det, dt = (hss * gt - hst * gs) / det;
    s -= ds;
    t -= dt;
    if (fabs(ds) + fabs(dt) < 1e-15) break;
  }
  eta[0] = s;
  eta[1] = t;
}

// records a pair produces: every (i, j) gives one record for its own row (a structural zero when the row is constrained)
// plus one per master of the row's constraint line
__global__ void fe_count_kernel(uint64_t n_pairs, const uint32_t *__restrict__ pair_bg, const uint32_t *__restrict__ bg_dofs,
                                const int32_t *__restrict__ line_of, const uint64_t *__restrict__ c_ptr, uint32_t n_space_dofs,
                                uint32_t *__restrict__ cnt, unsigned *bad)
{
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n_pairs) return;
  if (p == n_pairs) { cnt[p] = 0; return; }
  const int nb = c_fe.n_bg, ni = c_fe.n_im;
  const uint32_t *dofs = bg_dofs + (uint64_t)pair_bg[p] * nb;
  uint32_t rows = 0;
  for (int i = 0; i < nb; ++i) {
    const uint32_t d = dofs[i];
    if (d >= n_space_dofs) { atomicOr(bad, 1u); continue; }
    const int32_t ln = line_of[d];
    rows += 1u + (ln >= 0 ? (uint32_t)(c_ptr[ln + 1] - c_ptr[ln]) : 0u);
  }
  cnt[p] = rows * (uint32_t)ni;
}

template <int D, int NBMAX, int NIMAX>
__global__ void __launch_bounds__(64) fe_fill_kernel(uint64_t n_pairs, const uint32_t *__restrict__ pair_im, const uint32_t *__restrict__ pair_bg,
                                                     const uint64_t *__restrict__ q_off, const double *__restrict__ pts,
                                                     const double *__restrict__ wts, const double *__restrict__ bg_box,
                                                     const double *__restrict__ im_verts, const uint32_t *__restrict__ bg_dofs,
                                                     const int32_t *__restrict__ line_of, const uint64_t *__restrict__ c_ptr,
                                                     const uint32_t *__restrict__ c_master, const double *__restrict__ c_weight,
                                                     uint32_t n_space_dofs, const uint64_t *__restrict__ rec_off,
                                                     unsigned long long *__restrict__ keys, double *__restrict__ vals)
{
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  constexpr int NVI = 1 << (D - 1);
  const int nb = c_fe.n_bg, ni = c_fe.n_im;
  const uint32_t m = pair_im[p], b = pair_bg[p];
  const double *box = bg_box + (uint64_t)b * 8;
  const double *iv = im_verts + (uint64_t)m * NVI * D;
  double ih[3];
#pragma unroll
  for (int k = 0; k < D; ++k) ih[k] = 1.0 / (box[4 + k] - box[k]);
  double blk[NBMAX * NIMAX];
  for (int e = 0; e < nb * ni; ++e) blk[e] = 0.0;
  for (uint64_t q = q_off[p]; q < q_off[p + 1]; ++q) {
    const double *x = pts + q * D;
    double Lb[3][FE_MAXN], Li[2][FE_MAXN], eta[2];
#pragma unroll
    for (int k = 0; k < D; ++k) lagrange_all(c_fe.bg, (x[k] - box[k]) * ih[k], Lb[k]);
    immersed_unit_coords<D>(iv, x, eta);
#pragma unroll
    for (int k = 0; k < D - 1; ++k) lagrange_all(c_fe.im, eta[k], Li[k]);
    const double w = wts[q];
    double psi[NIMAX];
    for (int j = 0; j < ni; ++j) psi[j] = D == 2 ? Li[0][c_fe.im_idx[j][0]] : Li[0][c_fe.im_idx[j][0]] * Li[1][c_fe.im_idx[j][1]];
    for (int i = 0; i < nb; ++i) {
      double phi = Lb[0][c_fe.bg_idx[i][0]] * Lb[1][c_fe.bg_idx[i][1]];
      if (D == 3) phi *= Lb[2][c_fe.bg_idx[i][2]];
      const double pw = phi * w;
      for (int j = 0; j < ni; ++j) blk[i * ni + j] += pw * psi[j];
    }
  }
  // records, constraint lines expanded (reference: distribute_local_to_global, coupling_utilities.h:285-287)
  uint64_t o = rec_off[p];
  const uint32_t *dofs = bg_dofs + (uint64_t)b * nb;
  for (int i = 0; i < nb; ++i) {
    const uint32_t d = dofs[i];
    if (d >= n_space_dofs) continue;
    const int32_t ln = line_of[d];
    for (int j = 0; j < ni; ++j) {
      const unsigned long long row = ((unsigned long long)((uint64_t)m * ni + j)) << 32;
      const double v = blk[i * ni + j];
      keys[o] = row | d;
      vals[o] = ln < 0 ? v : 0.0;   // keep_constrained_entries: the constrained row stays in the pattern, structurally zero
      ++o;
      if (ln >= 0)
        for (uint64_t k = c_ptr[ln]; k < c_ptr[ln + 1]; ++k) { keys[o] = row | c_master[k]; vals[o] = c_weight[k] * v; ++o; }
    }
  }
}

__global__ void fe_heads_kernel(uint64_t n, const unsigned long long *__restrict__ keys, uint32_t *__restrict__ head)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
  if (i == n) head[i] = 0u;
}

// one matrix entry per run of equal keys, summed in record order; row starts and lengths of C on the way
__global__ void fe_reduce_kernel(uint64_t n, const unsigned long long *__restrict__ keys, const double *__restrict__ vals,
                                 const uint32_t *__restrict__ head, const uint64_t *__restrict__ pos, uint32_t *__restrict__ c_col,
                                 double *__restrict__ c_val, uint64_t *__restrict__ rowstart, uint32_t *__restrict__ rowlen)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !head[i]) return;
  const unsigned long long k = keys[i];
  double s = vals[i];
  for (uint64_t j = i + 1; j < n && keys[j] == k; ++j) s += vals[j];
  const uint64_t e = pos[i];
  c_col[e] = (uint32_t)k;
  c_val[e] = s;
  const uint32_t row = (uint32_t)(k >> 32);
  if (i == 0 || (uint32_t)(keys[i - 1] >> 32) != row) rowstart[row] = e;
  atomicAdd(&rowlen[row], 1u);
}

template <typename T>
__global__ void fe_gather(uint64_t n, const uint64_t *__restrict__ idx, const T *__restrict__ src, T *__restrict__ dst)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

// dof -> node index per direction from the unit support points; fills `T` with the distinct 1D nodes
bool fe_table_from_support_points(int dim, unsigned degree, unsigned dpc, const double *sp, FeTable &T, unsigned char (*idx)[3], unsigned char (*idx2)[2])
{
  const unsigned n1 = degree + 1;
  unsigned expect = 1;
  for (int k = 0; k < dim; ++k) expect *= n1;
  if (n1 > (unsigned)FE_MAXN || dpc != expect) return false;
  std::vector<double> nodes;
  const double eps = 1e-10;
  for (unsigned i = 0; i < dpc; ++i) {
    const double x = dim > 0 ? sp[i * dim] : 0.5;
    bool seen = false;
    for (double y : nodes) seen |= fabs(x - y) < eps;
    if (!seen) nodes.push_back(x);
  }
  if (dim == 0) nodes.assign(1, 0.5);
  if (nodes.size() != n1) return false;
  for (size_t a = 0; a < nodes.size(); ++a)
    for (size_t b = a + 1; b < nodes.size(); ++b)
      if (nodes[b] < nodes[a]) { const double t = nodes[a]; nodes[a] = nodes[b]; nodes[b] = t; }
  T.n1 = (int)n1;
  for (unsigned a = 0; a < n1; ++a) {
    T.node[a] = nodes[a];
    double den = 1.0;
    for (unsigned b = 0; b < n1; ++b)
      if (b != a) den *= nodes[a] - nodes[b];
    T.inv_den[a] = 1.0 / den;
  }
  std::vector<char> used(dpc, 0);
  for (unsigned i = 0; i < dpc; ++i) {
    unsigned flat = 0, stride = 1;
    for (int k = 0; k < dim; ++k) {
      int found = -1;
      for (unsigned a = 0; a < n1; ++a)
        if (fabs(sp[i * dim + k] - nodes[a]) < eps) found = (int)a;
      if (found < 0) return false;
      if (idx) idx[i][k] = (unsigned char)found;
      if (idx2 && k < 2) idx2[i][k] = (unsigned char)found;
      flat += (unsigned)found * stride;
      stride *= n1;
    }
    if (used[flat]) return false;   // two dofs on one support point: not a Lagrange basis on a tensor grid
    used[flat] = 1;
  }
  return true;
}

}  // namespace

extern "C" {

int nm_set_finite_elements(nm_ctx *ctx, const nm_fe *space_fe, const uint32_t *space_dofs, uint64_t n_space_dofs, uint64_t n_constrained,
                           const uint32_t *c_dof, const uint64_t *c_ptr, const uint32_t *c_master, const double *c_weight,
                           const nm_fe *immersed_fe, const uint32_t *immersed_dofs, uint64_t n_immersed_dofs, int where)
{
  if (!ctx) return NM_ERR_INVALID;
  NM_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->bg_set || !ctx->im_set) return nm_fail(ctx, NM_ERR_STATE, "nm_set_background and nm_set_immersed must precede nm_set_finite_elements");
  if (ctx->codim0)
    return nm_fail(ctx, NM_ERR_UNSUPPORTED, "general elements are implemented for codimension-1 immersed grids; the codimension-0 case <2,2,2> takes FE_Q(1) x FE_DGQ(0)");
  if (!space_fe || !immersed_fe || !space_fe->unit_support_points || (ctx->n_bg && !space_dofs) || (ctx->n_im && !immersed_dofs))
    return nm_fail(ctx, NM_ERR_INVALID, "null finite element description or DoF table");
  if (immersed_fe->dofs_per_cell > 1 && !immersed_fe->unit_support_points) return nm_fail(ctx, NM_ERR_INVALID, "null immersed support points");
  if (n_space_dofs >= 0xffffffffULL || n_immersed_dofs >= 0xffffffffULL) return nm_fail(ctx, NM_ERR_UNSUPPORTED, "DoF counts must fit 32 bits");
  const int d = ctx->spacedim;
  if (space_fe->dofs_per_cell > (unsigned)FE_MAX_BG || immersed_fe->dofs_per_cell > (unsigned)FE_MAX_IM)
    return nm_fail(ctx, NM_ERR_UNSUPPORTED, "at most %d background and %d immersed DoFs per cell are implemented (got %u, %u)", FE_MAX_BG, FE_MAX_IM,
                   space_fe->dofs_per_cell, immersed_fe->dofs_per_cell);
  FeConst F;
  memset(&F, 0, sizeof(F));
  if (!fe_table_from_support_points(d, space_fe->degree, space_fe->dofs_per_cell, space_fe->unit_support_points, F.bg, F.bg_idx, nullptr))
    return nm_fail(ctx, NM_ERR_UNSUPPORTED,
                   "the background element is not a tensor-product Lagrange element of degree %u with %u DoFs per cell on distinct support points "
                   "(FE_Q(p), p <= %d)", space_fe->degree, space_fe->dofs_per_cell, FE_MAXN - 1);
  const double mid[2] = {0.5, 0.5};
  const double *isp = immersed_fe->unit_support_points ? immersed_fe->unit_support_points : mid;
  if (!fe_table_from_support_points(d - 1, immersed_fe->degree, immersed_fe->dofs_per_cell, isp, F.im, nullptr, F.im_idx))
    return nm_fail(ctx, NM_ERR_UNSUPPORTED,
                   "the immersed element is not a tensor-product Lagrange element of degree %u with %u DoFs per cell (FE_DGQ(q), q <= %d)",
                   immersed_fe->degree, immersed_fe->dofs_per_cell, FE_MAXN - 1);
  F.n_bg = (int)space_fe->dofs_per_cell;
  F.n_im = (int)immersed_fe->dofs_per_cell;
  ctx->fe = std::vector<unsigned char>(reinterpret_cast<unsigned char *>(&F), reinterpret_cast<unsigned char *>(&F) + sizeof(F));
  ctx->fe_nb = F.n_bg;
  ctx->fe_ni = F.n_im;
  ctx->fe_general = true;
  ctx->matrix_valid = false;
  ctx->n_im_dofs = n_immersed_dofs;
  NM_TRY(nm_upload(ctx, ctx->fe_bg_dofs, space_dofs, ctx->n_bg * F.n_bg * sizeof(uint32_t), where));
  NM_TRY(nm_upload(ctx, ctx->fe_im_dofs, immersed_dofs, ctx->n_im * F.n_im * sizeof(uint32_t), where));
  NM_TRY(nm_constraints_set(ctx, n_space_dofs, n_constrained, c_dof, c_ptr, c_master, c_weight, where));
  NM_SYNC(ctx);
  return NM_OK;
}

int nm_assemble_fe(nm_ctx *ctx, uint64_t n_pairs, const uint32_t *immersed, const uint32_t *background, const uint64_t *q_offset,
                   const double *points, const double *weights, int where, uint64_t *nnz_out)
{
  if (!ctx) return NM_ERR_INVALID;
  NM_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->fe_general) return nm_fail(ctx, NM_ERR_STATE, "nm_set_finite_elements has not been called");
  if (!ctx->bg_set || !ctx->im_set) return nm_fail(ctx, NM_ERR_STATE, "grids have not been set");
  const bool own = immersed == nullptr;
  if (!own && n_pairs && (!background || !q_offset || !points || !weights)) return nm_fail(ctx, NM_ERR_INVALID, "null quadrature arrays");
  if (own && !ctx->intersect_valid) return nm_fail(ctx, NM_ERR_STATE, "nm_intersect has not been run (or pass the quadrature explicitly)");
  const int d = ctx->spacedim, nb = ctx->fe_nb, ni = ctx->fe_ni;
  PhaseTimer tm(ctx, NM_T_MERGE);
  NM_CUDA(cudaMemcpyToSymbolAsync(c_fe, ctx->fe.data(), sizeof(FeConst), 0, cudaMemcpyHostToDevice, ctx->stream));
  // ---- the pair list and its quadrature on the device ----
  const uint32_t *p_im = immersed, *p_bg = background;
  const uint64_t *qo = q_offset;
  const double *pt = points, *wt = weights;
  uint64_t nq = 0;
  if (own) {
    n_pairs = ctx->counts.n_accepted;
    nq = ctx->counts.n_qpoints;
    uint64_t *idx = nullptr;
    NM_TRY(nm_accepted_index(ctx, &idx));
    NM_TRY(nm_ensure(ctx, ctx->fe_pairs, (2 * n_pairs + 2) * sizeof(uint32_t)));
    uint32_t *pi = ctx->fe_pairs.as<uint32_t>(), *pb = pi + n_pairs;
    if (n_pairs) {
      fe_gather<uint32_t><<<nm_blocks(n_pairs, 256), 256, 0, NM_LS(ctx)>>>(n_pairs, idx, ctx->cand_im.as<uint32_t>(), pi);
      fe_gather<uint32_t><<<nm_blocks(n_pairs, 256), 256, 0, NM_LS(ctx)>>>(n_pairs, idx, ctx->cand_bg.as<uint32_t>(), pb);
      NM_CUDA(cudaGetLastError());
    }
    NM_TRY(nm_ensure(ctx, ctx->fe_qoff, (n_pairs + 2) * sizeof(uint64_t)));
    NM_TRY(nm_ensure(ctx, ctx->fe_pts, (nq + 1) * d * sizeof(double)));
    NM_TRY(nm_ensure(ctx, ctx->fe_w, (nq + 1) * sizeof(double)));
    NM_TRY(nm_quadrature_emit(ctx, ctx->fe_qoff.as<uint64_t>(), ctx->fe_pts.as<double>(), ctx->fe_w.as<double>(), NM_DEVICE));
    p_im = pi; p_bg = pb; qo = ctx->fe_qoff.as<uint64_t>(); pt = ctx->fe_pts.as<double>(); wt = ctx->fe_w.as<double>();
  } else if (where == NM_HOST) {
    nq = n_pairs ? q_offset[n_pairs] : 0;
    NM_TRY(nm_ensure(ctx, ctx->fe_pairs, (2 * n_pairs + 2) * sizeof(uint32_t)));
    uint32_t *pi = ctx->fe_pairs.as<uint32_t>(), *pb = pi + n_pairs;
    if (n_pairs) {
      NM_CUDA(cudaMemcpyAsync(pi, immersed, n_pairs * 4, cudaMemcpyHostToDevice, ctx->stream));
      NM_CUDA(cudaMemcpyAsync(pb, background, n_pairs * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    NM_TRY(nm_upload(ctx, ctx->fe_qoff, q_offset, (n_pairs + 1) * sizeof(uint64_t), NM_HOST));
    NM_TRY(nm_upload(ctx, ctx->fe_pts, points, nq * d * sizeof(double), NM_HOST));
    NM_TRY(nm_upload(ctx, ctx->fe_w, weights, nq * sizeof(double), NM_HOST));
    p_im = pi; p_bg = pb; qo = ctx->fe_qoff.as<uint64_t>(); pt = ctx->fe_pts.as<double>(); wt = ctx->fe_w.as<double>();
  }
  // ---- records ----
  const uint64_t n_crows = ctx->n_im * (uint64_t)ni;
  if (n_crows >= 0xffffffffULL) return nm_fail(ctx, NM_ERR_UNSUPPORTED, "immersed cells x DoFs per cell must fit 32 bits");
  NM_TRY(nm_ensure(ctx, ctx->scratch_a, (n_pairs + 2) * sizeof(uint32_t)));
  NM_TRY(nm_ensure(ctx, ctx->fe_recoff, (n_pairs + 2) * sizeof(uint64_t)));
  unsigned *bad = ctx->counters.as<unsigned>() + 50;
  NM_CUDA(cudaMemsetAsync(bad, 0, 4, ctx->stream));
  fe_count_kernel<<<nm_blocks(n_pairs + 1, 128), 128, 0, NM_LS(ctx)>>>(n_pairs, p_bg, ctx->fe_bg_dofs.as<uint32_t>(), ctx->line_of.as<int32_t>(),
                                                                       ctx->c_ptr.as<uint64_t>(), (uint32_t)ctx->n_space_dofs,
                                                                       ctx->scratch_a.as<uint32_t>(), bad);
  NM_CUDA(cudaGetLastError());
  NM_TRY(nm_exclusive_scan_u32_to_u64(ctx, ctx->scratch_a.as<uint32_t>(), ctx->fe_recoff.as<uint64_t>(), n_pairs + 1));
  uint64_t n_rec = 0;
  unsigned h_bad = 0;
  NM_CUDA(cudaMemcpyAsync(&n_rec, ctx->fe_recoff.as<uint64_t>() + n_pairs, 8, cudaMemcpyDeviceToHost, ctx->stream));
  NM_CUDA(cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
  NM_SYNC(ctx);
  if (h_bad) return nm_fail(ctx, NM_ERR_INVALID, "background DoF table holds indices >= n_dofs (%llu)", (unsigned long long)ctx->n_space_dofs);
  NM_TRY(nm_ensure(ctx, ctx->fe_keys, 2 * (n_rec + 1) * sizeof(unsigned long long)));
  NM_TRY(nm_ensure(ctx, ctx->fe_vals, 2 * (n_rec + 1) * sizeof(double)));
  unsigned long long *k_in = ctx->fe_keys.as<unsigned long long>(), *k_out = k_in + n_rec + 1;
  double *v_in = ctx->fe_vals.as<double>(), *v_out = v_in + n_rec + 1;
  if (n_pairs) {
    const bool small = nb <= 27 && ni <= 4;
#define FE_LAUNCH(D, NB, NI)                                                                                                             \
  fe_fill_kernel<D, NB, NI><<<nm_blocks(n_pairs, 64), 64, 0, NM_LS(ctx)>>>(                                                               \
      n_pairs, p_im, p_bg, qo, pt, wt, ctx->bg_box.as<double>(), ctx->im_verts.as<double>(), ctx->fe_bg_dofs.as<uint32_t>(),               \
      ctx->line_of.as<int32_t>(), ctx->c_ptr.as<uint64_t>(), ctx->c_master.as<uint32_t>(), ctx->c_weight.as<double>(),                     \
      (uint32_t)ctx->n_space_dofs, ctx->fe_recoff.as<uint64_t>(), k_in, v_in)
    if (d == 2) { if (small) FE_LAUNCH(2, 27, 4); else FE_LAUNCH(2, FE_MAX_BG, FE_MAX_IM); }
    else { if (small) FE_LAUNCH(3, 27, 4); else FE_LAUNCH(3, FE_MAX_BG, FE_MAX_IM); }
#undef FE_LAUNCH
    NM_CUDA(cudaGetLastError());
  }
  // ---- group, sum, rows of C ----
  int row_bits = 1;
  while ((1ULL << row_bits) < n_crows + 1) ++row_bits;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, v_out, n_rec, 0, 32 + row_bits, ctx->stream);
  NM_TRY(nm_ensure(ctx, ctx->scan_tmp, tmp_bytes));
  if (n_rec) NM_CUDA(cub::DeviceRadixSort::SortPairs(ctx->scan_tmp.p, tmp_bytes, k_in, k_out, v_in, v_out, n_rec, 0, 32 + row_bits, ctx->stream));
  NM_TRY(nm_ensure(ctx, ctx->scratch_a, (n_rec + 2) * sizeof(uint32_t)));
  NM_TRY(nm_ensure(ctx, ctx->scratch_b, (n_rec + 2) * sizeof(uint64_t)));
  fe_heads_kernel<<<nm_blocks(n_rec + 1, 256), 256, 0, NM_LS(ctx)>>>(n_rec, k_out, ctx->scratch_a.as<uint32_t>());
  NM_CUDA(cudaGetLastError());
  NM_TRY(nm_exclusive_scan_u32_to_u64(ctx, ctx->scratch_a.as<uint32_t>(), ctx->scratch_b.as<uint64_t>(), n_rec + 1));
  uint64_t n_unique = 0;
  NM_CUDA(cudaMemcpyAsync(&n_unique, ctx->scratch_b.as<uint64_t>() + n_rec, 8, cudaMemcpyDeviceToHost, ctx->stream));
  NM_SYNC(ctx);
  NM_TRY(nm_ensure(ctx, ctx->c_rowstart, (n_crows + 2) * sizeof(uint64_t)));
  NM_TRY(nm_ensure(ctx, ctx->c_rowlen, (n_crows + 1) * sizeof(uint32_t)));
  NM_TRY(nm_ensure(ctx, ctx->c_col, (n_unique + 1) * sizeof(uint32_t)));
  NM_TRY(nm_ensure(ctx, ctx->c_val, (n_unique + 1) * sizeof(double)));
  NM_CUDA(cudaMemsetAsync(ctx->c_rowstart.p, 0, (n_crows + 2) * sizeof(uint64_t), ctx->stream));
  NM_CUDA(cudaMemsetAsync(ctx->c_rowlen.p, 0, (n_crows + 1) * sizeof(uint32_t), ctx->stream));
  if (n_rec) {
    fe_reduce_kernel<<<nm_blocks(n_rec, 256), 256, 0, NM_LS(ctx)>>>(n_rec, k_out, v_out, ctx->scratch_a.as<uint32_t>(), ctx->scratch_b.as<uint64_t>(),
                                                                    ctx->c_col.as<uint32_t>(), ctx->c_val.as<double>(), ctx->c_rowstart.as<uint64_t>(),
                                                                    ctx->c_rowlen.as<uint32_t>());
    NM_CUDA(cudaGetLastError());
  }
  tm.stop();
  // ---- stored orientation ----
  ctx->n_crows = n_crows;
  ctx->crow_dofs = &ctx->fe_im_dofs;
  ctx->compact_rows = false;
  ctx->n_local_rows = ctx->n_space_dofs;
  ctx->out_rows_valid = false;
  ctx->merge_path = 2;
  if (ctx->download_pending) { NM_TRY(nm_wait_download_issued(ctx)); NM_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_down_done, 0)); ctx->download_pending = false; }
  NM_TRY(nm_transpose_build(ctx));
  if (nnz_out) *nnz_out = ctx->nnz;
  return NM_OK;
}

}  // extern "C"
