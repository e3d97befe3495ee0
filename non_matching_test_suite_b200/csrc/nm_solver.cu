// The Schur-complement solve of the Lagrange-multiplier drivers around the GPU C / C^T products.
//
// Replaces PoissonLM::solve()   code/examples/lagrange_multiplier/2d3d/lagrange_multiplier.cc:613-659
//                                 code/examples/lagrange_multiplier/1d2d/lagrange_multiplier.cc:675-730
//   K_inv          = inverse_operator(K, SolverCG(ReductionControl), preconditioner)            :618-624
//   preconditioner = C K Ct + M   (applied by multiplication, as the reference does)             :630
//   S              = C K_inv Ct                                                                 :632
//   lambda         = S_inv (C K_inv space_rhs - embedded_rhs),  S_inv = GMRES(ReductionControl) :636-642
//   solution       = K_inv (space_rhs - Ct lambda);  space_constraints.distribute(solution)     :643, :658
// with Ct = the context's assembled coupling matrix (stored orientation) and C its transpose.  K and M are the caller's
// matrices (CSR).  Everything runs on the context's stream: CSR products, vector updates and reductions are kernels of this
// file; only the (basis x basis) Hessenberg / Givens bookkeeping of GMRES is host arithmetic, as in deal.II.
// [UPSTREAM] restated from deal.II 9.4: ReductionControl stops at |r| <= tol or |r| <= reduce |r_0|; SolverGMRES is
// restarted, left-preconditioned (the preconditioned residual is what is checked), modified Gram-Schmidt, basis size
// max_n_tmp_vectors - 2 = 28.  Deviation: the reference preconditions the CG on K with Trilinos ML (AMG); here the
// preconditioner is the diagonal of K -- with the inner tolerance tight both give the same K_inv, only the inner
// iteration count differs.
#include <math.h>

#include <vector>

#include "nm_common.cuh"

namespace {

__global__ void csr_spmv_kernel(uint64_t n, const uint64_t *__restrict__ rp, const uint32_t *__restrict__ col, const double *__restrict__ val,
                                const double *__restrict__ x, double *__restrict__ y)
{
  // 8 lanes per row
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t r = t >> 3;
  const unsigned l = threadIdx.x & 7;
  double s = 0.0;
  if (r < n)
    for (uint64_t k = rp[r] + l; k < rp[r + 1]; k += 8) s = fma(val[k], x[col[k]], s);
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, 8);
  if (r < n && l == 0) y[r] = s;
}

__global__ void csr_diag_inv_kernel(uint64_t n, const uint64_t *__restrict__ rp, const uint32_t *__restrict__ col, const double *__restrict__ val,
                                    double *__restrict__ dinv)
{
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double d = 0.0;
  for (uint64_t k = rp[r]; k < rp[r + 1]; ++k)
    if (col[k] == r) d = val[k];
  dinv[r] = d != 0.0 ? 1.0 / d : 1.0;
}

// y = a x + b y
__global__ void axpby_kernel(uint64_t n, double a, const double *__restrict__ x, double b, double *__restrict__ y)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = b == 0.0 ? a * x[i] : fma(a, x[i], b * y[i]);
}
// z = x .* y
__global__ void mul_kernel(uint64_t n, const double *__restrict__ x, const double *__restrict__ y, double *__restrict__ z)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) z[i] = x[i] * y[i];
}

constexpr unsigned DOT_BLOCKS = 592;   // 4 per SM
__global__ void __launch_bounds__(256) dot_partial_kernel(uint64_t n, const double *__restrict__ x, const double *__restrict__ y, double *__restrict__ part)
{
  __shared__ double s[8];
  double a = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) a = fma(x[i], y[i], a);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += s[k];   // fixed order: bit-reproducible
    part[blockIdx.x] = t;
  }
}
__global__ void dot_final_kernel(unsigned nb, const double *__restrict__ part, double *__restrict__ out)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (unsigned k = 0; k < nb; ++k) t += part[k];
    *out = t;
  }
}

// u[c] = inhomogeneity + sum_k w_k u[master_k] for every constrained dof (AffineConstraints::distribute)
__global__ void distribute_kernel(uint64_t n_c, const uint32_t *__restrict__ c_dof, const uint64_t *__restrict__ c_ptr,
                                  const uint32_t *__restrict__ c_master, const double *__restrict__ c_weight,
                                  const double *__restrict__ inhom, double *__restrict__ u)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_c) return;
  double s = inhom ? inhom[i] : 0.0;
  for (uint64_t k = c_ptr[i]; k < c_ptr[i + 1]; ++k) s = fma(c_weight[k], u[c_master[k]], s);
  u[c_dof[i]] = s;
}

struct Csr {
  uint64_t n = 0;
  const uint64_t *rp = nullptr;
  const uint32_t *col = nullptr;
  const double *val = nullptr;
};

struct Solver {
  nm_ctx *ctx;
  uint64_t ns, ni;
  Csr K, M;
  double *dinv;           // 1 / diag(K)
  double *part, *scalar;  // reduction scratch
  uint64_t inner_its = 0;
  unsigned inner_max;
  double inner_tol, inner_reduce;
  // work vectors of the operators
  double *ws0, *ws1, *ws2, *ws3, *wi0;

  unsigned nb(uint64_t n) const { return nm_blocks(n, 256); }
  int spmv(const Csr &A, const double *x, double *y)
  {
    if (A.n) csr_spmv_kernel<<<nm_blocks(A.n * 8, 256), 256, 0, NM_LS(ctx)>>>(A.n, A.rp, A.col, A.val, x, y);
    return cudaGetLastError() == cudaSuccess ? NM_OK : NM_ERR_CUDA;
  }
  void axpby(uint64_t n, double a, const double *x, double b, double *y)
  {
    if (n) axpby_kernel<<<nb(n), 256, 0, NM_LS(ctx)>>>(n, a, x, b, y);
  }
  int dot(uint64_t n, const double *x, const double *y, double *out)
  {
    dot_partial_kernel<<<DOT_BLOCKS, 256, 0, NM_LS(ctx)>>>(n, x, y, part);
    dot_final_kernel<<<1, 32, 0, NM_LS(ctx)>>>(DOT_BLOCKS, part, scalar);
    NM_CUDA(cudaMemcpyAsync(out, scalar, 8, cudaMemcpyDeviceToHost, ctx->stream));
    NM_SYNC(ctx);
    return NM_OK;
  }
  // x = K^-1 b: preconditioned CG, ReductionControl(inner_max, inner_tol, inner_reduce), start vector 0
  int k_inv(const double *b, double *x)
  {
    double *r = ws1, *z = ws2, *p = ws3, *q = ws0;   // ws0 doubles as q: callers never pass it as b or x
    NM_CUDA(cudaMemsetAsync(x, 0, ns * 8, ctx->stream));
    NM_CUDA(cudaMemcpyAsync(r, b, ns * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    double rr = 0, rz = 0, rz_new = 0, pq = 0;
    NM_TRY(dot(ns, r, r, &rr));
    const double r0 = sqrt(rr);
    if (r0 <= inner_tol) return NM_OK;
    mul_kernel<<<nb(ns), 256, 0, NM_LS(ctx)>>>(ns, dinv, r, z);
    NM_CUDA(cudaMemcpyAsync(p, z, ns * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    NM_TRY(dot(ns, r, z, &rz));
    for (unsigned it = 1; it <= inner_max; ++it) {
      NM_TRY(spmv(K, p, q));
      NM_TRY(dot(ns, p, q, &pq));
      if (!(pq > 0.0)) return nm_fail(ctx, NM_ERR_INVALID, "the stiffness matrix is not positive definite (p.Kp = %g in the inner CG)", pq);
      const double alpha = rz / pq;
      axpby(ns, alpha, p, 1.0, x);
      axpby(ns, -alpha, q, 1.0, r);
      ++inner_its;
      NM_TRY(dot(ns, r, r, &rr));
      const double rn = sqrt(rr);
      if (rn <= inner_tol || rn <= inner_reduce * r0) return NM_OK;
      mul_kernel<<<nb(ns), 256, 0, NM_LS(ctx)>>>(ns, dinv, r, z);
      NM_TRY(dot(ns, r, z, &rz_new));
      axpby(ns, 1.0, z, rz_new / rz, p);
      rz = rz_new;
    }
    return nm_fail(ctx, NM_ERR_STATE, "the inner CG on the stiffness matrix did not reach its tolerance in %u iterations", inner_max);
  }
};

}  // namespace

extern "C" int nm_schur_solve(nm_ctx *ctx, const nm_schur_problem *P, double *lambda, double *solution, nm_schur_stats *stats, int where)
{
  if (!ctx) return NM_ERR_INVALID;
  NM_CUDA(cudaSetDevice(ctx->device));
  if (!P || !lambda || !solution) return nm_fail(ctx, NM_ERR_INVALID, "null problem or result vector");
  if (!ctx->matrix_valid) return nm_fail(ctx, NM_ERR_STATE, "the coupling matrix has not been assembled");
  if (P->n_space != ctx->n_space_dofs || P->n_immersed != ctx->n_im_dofs)
    return nm_fail(ctx, NM_ERR_INVALID, "matrix sizes %llu x %llu do not match the coupling matrix (%llu x %llu)", (unsigned long long)P->n_space,
                   (unsigned long long)P->n_immersed, (unsigned long long)ctx->n_space_dofs, (unsigned long long)ctx->n_im_dofs);
  if (!P->k_rowptr || !P->k_col || !P->k_val || !P->m_rowptr || !P->m_col || !P->m_val || !P->space_rhs || !P->embedded_rhs)
    return nm_fail(ctx, NM_ERR_INVALID, "null matrix or right-hand-side array");
  if (where != NM_HOST && where != NM_DEVICE) return nm_fail(ctx, NM_ERR_INVALID, "where must be NM_HOST or NM_DEVICE");
  const uint64_t ns = P->n_space, ni = P->n_immersed;
  const unsigned m = P->gmres_basis ? P->gmres_basis : 28;
  cudaEvent_t e0, e1;
  NM_CUDA(cudaEventCreate(&e0));
  NM_CUDA(cudaEventCreate(&e1));
  NM_CUDA(cudaEventRecord(e0, ctx->stream));

  // ---- inputs on the device ----
  uint64_t k_nnz = 0, m_nnz = 0;
  if (where == NM_HOST) { k_nnz = P->k_rowptr[ns]; m_nnz = P->m_rowptr[ni]; }
  else {
    NM_CUDA(cudaMemcpyAsync(&k_nnz, P->k_rowptr + ns, 8, cudaMemcpyDeviceToHost, ctx->stream));
    NM_CUDA(cudaMemcpyAsync(&m_nnz, P->m_rowptr + ni, 8, cudaMemcpyDeviceToHost, ctx->stream));
    NM_SYNC(ctx);
  }
  auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
  const size_t b_in = al((ns + 1) * 8) + al(k_nnz * 4) + al(k_nnz * 8) + al((ni + 1) * 8) + al(m_nnz * 4) + al(m_nnz * 8) + al(ns * 8) + al(ni * 8) +
                      al(ctx->n_constrained * 8 + 8) + al(ctx->n_constrained * 4 + 4);
  const size_t b_work = 10 * al(ns * 8) + (size_t)(m + 10) * al(ni * 8) + al(DOT_BLOCKS * 8) + (size_t(64) << 10);
  NM_TRY(nm_ensure(ctx, ctx->solver_buf, b_in + b_work + 1024));
  char *base = ctx->solver_buf.as<char>();
  auto carve = [&](size_t bytes) { char *p = base; base += al(bytes); return p; };
  const auto kind = where == NM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  auto put = [&](const void *src, size_t bytes) -> void * {
    void *d = carve(bytes);
    if (bytes && src) cudaMemcpyAsync(d, src, bytes, kind, ctx->stream);
    return d;
  };
  Solver S;
  S.ctx = ctx; S.ns = ns; S.ni = ni;
  S.K.n = ns; S.K.rp = (const uint64_t *)put(P->k_rowptr, (ns + 1) * 8); S.K.col = (const uint32_t *)put(P->k_col, k_nnz * 4); S.K.val = (const double *)put(P->k_val, k_nnz * 8);
  S.M.n = ni; S.M.rp = (const uint64_t *)put(P->m_rowptr, (ni + 1) * 8); S.M.col = (const uint32_t *)put(P->m_col, m_nnz * 4); S.M.val = (const double *)put(P->m_val, m_nnz * 8);
  const double *f = (const double *)put(P->space_rhs, ns * 8);
  const double *g = (const double *)put(P->embedded_rhs, ni * 8);
  const double *inhom = P->inhomogeneity ? (const double *)put(P->inhomogeneity, ctx->n_constrained * 8) : nullptr;
  // the constrained dof ids in line order (the context keeps line_of: rebuild the list)
  S.dinv = (double *)carve(ns * 8);
  S.ws0 = (double *)carve(ns * 8); S.ws1 = (double *)carve(ns * 8); S.ws2 = (double *)carve(ns * 8); S.ws3 = (double *)carve(ns * 8);
  double *sa = (double *)carve(ns * 8);                      // space-sized temporary of the operators
  S.wi0 = (double *)carve(ni * 8);
  double *x = (double *)carve(ni * 8), *r = (double *)carve(ni * 8), *w = (double *)carve(ni * 8), *b = (double *)carve(ni * 8), *t_i = (double *)carve(ni * 8);
  std::vector<double *> V(m + 1);
  for (unsigned j = 0; j <= m; ++j) V[j] = (double *)carve(ni * 8);
  S.part = (double *)carve(DOT_BLOCKS * 8);
  S.scalar = (double *)carve(256);
  NM_CUDA(cudaGetLastError());
  S.inner_max = P->inner_max_it ? P->inner_max_it : 2000;
  S.inner_tol = P->inner_tol;
  S.inner_reduce = P->inner_reduce;
  if (ns) csr_diag_inv_kernel<<<nm_blocks(ns, 256), 256, 0, NM_LS(ctx)>>>(ns, S.K.rp, S.K.col, S.K.val, S.dinv);

  // ---- the operators of the reference (every product with the coupling matrix is nm_spmv_run) ----
  auto Ct = [&](const double *xi, double *ys) { return nm_spmv_run(ctx, 0, xi, ys); };          // n_im -> n_space
  auto C = [&](const double *xs, double *yi) { return nm_spmv_run(ctx, 1, xs, yi); };           // n_space -> n_im
  double *sc = (double *)carve(ns * 8);   // results of K^-1 / K: k_inv uses ws0..ws3 internally, never sa or sc
  auto S_op = [&](const double *xi, double *yi) -> int {
    NM_TRY(Ct(xi, sa));
    NM_TRY(S.k_inv(sa, sc));
    return C(sc, yi);
  };
  auto P_op = [&](const double *xi, double *yi) -> int {                                        // (C K Ct + M) x
    NM_TRY(Ct(xi, sa));
    NM_TRY(S.spmv(S.K, sa, sc));
    NM_TRY(C(sc, yi));
    NM_TRY(S.spmv(S.M, xi, t_i));
    S.axpby(ni, 1.0, t_i, 1.0, yi);
    return NM_OK;
  };

  // ---- b = C K^-1 f - g ----
  NM_TRY(S.k_inv(f, sc));
  NM_TRY(C(sc, b));
  S.axpby(ni, -1.0, g, 1.0, b);

  // ---- left-preconditioned restarted GMRES on S x = b ----
  NM_CUDA(cudaMemsetAsync(x, 0, ni * 8, ctx->stream));
  unsigned it = 0;
  double rho = 0, rho0 = 0;
  bool converged = false;
  const unsigned max_it = P->outer_max_it ? P->outer_max_it : (unsigned)(ns < 0xffffffffULL ? ns : 0xffffffffULL);   // the 3D driver: ReductionControl(solution.size(), ...)
  std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), gamma(m + 1), y(m);
  bool first = true;
  while (!converged && it < max_it) {
    // r = P (b - S x)
    if (first) { NM_CUDA(cudaMemcpyAsync(w, b, ni * 8, cudaMemcpyDeviceToDevice, ctx->stream)); }
    else { NM_TRY(S_op(x, w)); S.axpby(ni, 1.0, b, -1.0, w); }
    NM_TRY(P_op(w, r));
    double rr = 0;
    NM_TRY(S.dot(ni, r, r, &rr));
    rho = sqrt(rr);
    if (first) { rho0 = rho; first = false; if (rho0 <= P->outer_tol) { converged = true; break; } }
    if (rho <= P->outer_tol || rho <= P->outer_reduce * rho0) { converged = true; break; }
    S.axpby(ni, 1.0 / rho, r, 0.0, V[0]);
    gamma[0] = rho;
    unsigned k = 0;
    for (; k < m && it < max_it; ++k) {
      NM_TRY(S_op(V[k], w));
      NM_TRY(P_op(w, r));                      // r = P S v_k
      for (unsigned i = 0; i <= k; ++i) {      // modified Gram-Schmidt
        double h = 0;
        NM_TRY(S.dot(ni, r, V[i], &h));
        H[(size_t)i * m + k] = h;
        S.axpby(ni, -h, V[i], 1.0, r);
      }
      double hh = 0;
      NM_TRY(S.dot(ni, r, r, &hh));
      hh = sqrt(hh);
      H[(size_t)(k + 1) * m + k] = hh;
      if (hh > 0) S.axpby(ni, 1.0 / hh, r, 0.0, V[k + 1]);
      for (unsigned i = 0; i < k; ++i) {       // earlier Givens rotations on the new column
        const double a = H[(size_t)i * m + k], c2 = H[(size_t)(i + 1) * m + k];
        H[(size_t)i * m + k] = cs[i] * a + sn[i] * c2;
        H[(size_t)(i + 1) * m + k] = -sn[i] * a + cs[i] * c2;
      }
      const double a = H[(size_t)k * m + k], c2 = H[(size_t)(k + 1) * m + k], d = hypot(a, c2);
      cs[k] = d > 0 ? a / d : 1.0;
      sn[k] = d > 0 ? c2 / d : 0.0;
      H[(size_t)k * m + k] = d;
      H[(size_t)(k + 1) * m + k] = 0.0;
      gamma[k + 1] = -sn[k] * gamma[k];
      gamma[k] = cs[k] * gamma[k];
      ++it;
      rho = fabs(gamma[k + 1]);
      if (rho <= P->outer_tol || rho <= P->outer_reduce * rho0) { converged = true; ++k; break; }
    }
    // x += V y,  H y = gamma
    for (int i = (int)k - 1; i >= 0; --i) {
      double s2 = gamma[i];
      for (unsigned j = i + 1; j < k; ++j) s2 -= H[(size_t)i * m + j] * y[j];
      y[i] = s2 / H[(size_t)i * m + i];
    }
    for (unsigned i = 0; i < k; ++i) S.axpby(ni, y[i], V[i], 1.0, x);
  }
  if (!converged) return nm_fail(ctx, NM_ERR_STATE, "GMRES on the Schur complement did not converge in %u iterations (residual %g)", it, rho);

  // ---- solution = K^-1 (f - Ct lambda); distribute ----
  NM_TRY(Ct(x, sa));
  S.axpby(ns, 1.0, f, -1.0, sa);
  double *u = (double *)carve(ns * 8);
  NM_TRY(S.k_inv(sa, u));
  if (ctx->n_constrained) {
    // constrained dof ids: recover them from line_of (the context does not keep the list)
    uint32_t *c_dof = (uint32_t *)carve(ctx->n_constrained * 4 + 4);
    NM_TRY(nm_constrained_dofs(ctx, c_dof));
    distribute_kernel<<<nm_blocks(ctx->n_constrained, 256), 256, 0, NM_LS(ctx)>>>(ctx->n_constrained, c_dof, ctx->c_ptr.as<uint64_t>(),
                                                                                  ctx->c_master.as<uint32_t>(), ctx->c_weight.as<double>(), inhom, u);
    NM_CUDA(cudaGetLastError());
  }
  double lam2 = 0;
  NM_TRY(S.dot(ni, x, x, &lam2));
  NM_TRY(nm_download(ctx, lambda, x, ni * 8, where));
  NM_TRY(nm_download(ctx, solution, u, ns * 8, where));
  NM_CUDA(cudaEventRecord(e1, ctx->stream));
  NM_SYNC(ctx);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (stats) {
    stats->outer_iterations = it;
    stats->inner_iterations = S.inner_its;
    stats->outer_residual = rho;
    stats->lambda_norm_sqr = lam2;
    stats->ms = ms;
  }
  return NM_OK;
}
