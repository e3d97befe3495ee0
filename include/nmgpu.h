/* nmgpu.h -- C ABI of libnmgpu.so: B200 (sm_100a) coupling-matrix assembly by exact
 * intersections, and the C / C^T sparse matrix-vector products.
 *
 * This is the drop-in boundary for ONE path of fdrmrc/non_matching_test_suite.  Each entry
 * point names the reference interface it replaces (file:line relative to the reference's
 * code/ directory).  The C++ wrapper include/coupling_utilities.h keeps the reference's
 * own signatures on top of this ABI; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C, no exceptions, no STL: every function returns an nm_status; a message for
 *     the last failure is available from nm_last_error().
 *   - the caller owns every buffer it passes or receives; nothing is retained from a host
 *     pointer after the call returns.  The context owns all device memory and its stream.
 *   - `where` says where the caller's buffers live: NM_HOST (pageable or pinned host
 *     memory) or NM_DEVICE (device memory of the context's GPU).
 *   - a context is not thread-safe; use one context per GPU / per rank.  Calls return
 *     after the work is complete (internally asynchronous on a private stream).
 *   - indices are 32-bit unsigned (deal.II types::global_dof_index without
 *     DEAL_II_WITH_64BIT_INDICES, as in the reference image), row pointers 64-bit.
 *   - matrix orientation: "stored" = n_space_dofs x n_immersed_dofs, the matrix the
 *     reference drivers call coupling_matrix (its vmult is their `Ct`, its Tvmult their `C`).
 *   - there is no CPU fallback: without a CUDA device every call fails with NM_ERR_CUDA.
 */
#ifndef NMGPU_H
#define NMGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nm_ctx nm_ctx;

typedef enum {
  NM_OK = 0,
  NM_ERR_INVALID = 1,     /* bad argument (sizes, null pointers, degree == 0, ...)          */
  NM_ERR_CUDA = 2,        /* CUDA runtime failure, or no device                             */
  NM_ERR_UNSUPPORTED = 3, /* outside the implemented scope (stated in the message)          */
  NM_ERR_STATE = 4,       /* call order: e.g. nm_build_sparsity before nm_intersect         */
  NM_ERR_NOMEM = 5
} nm_status;

enum { NM_HOST = 0, NM_DEVICE = 1,
       NM_HOST_ASYNC = 2,     /* nm_get_csr_compact only: pinned host buffers, copies complete at nm_sync_downloads() */
       NM_DEVICE_INPLACE = 3  /* nm_set_background*, nm_set_background_dofs, nm_set_immersed, nm_set_immersed_dofs only: device
                               * arrays; the DoF tables and the immersed vertices are READ WHERE THEY ARE instead of being copied
                               * into the context.  They must stay valid and unchanged until the same input is set again or the
                               * context is destroyed; the library never writes to them.  Everything else behaves as NM_DEVICE. */ };

/* Counters of one nm_intersect() call. */
typedef struct {
  uint64_t n_candidates;   /* closed bounding-box overlaps  (coupling_utilities.cc:53-55)   */
  uint64_t n_accepted;     /* pairs with sum of weights > tol (coupling_utilities.cc:59-65) */
  uint64_t n_qpoints;      /* quadrature points over all accepted pairs                     */
  uint64_t n_dropped;      /* sub-simplices dropped by |J| < 1e-12 (intersections.cc:710)   */
  uint64_t n_marginal;     /* candidates with 0 < measure < 1e-12 * measure(immersed cell): */
                           /* the pairs whose accept/reject decision is rounding-sensitive  */
  uint64_t n_split_ties;   /* immersed quads whose Delaunay split depends on CGAL's         */
                           /* insertion order (cocircular / coincident in xy-projection)    */
} nm_counts;

/* Device time of the phases of the last calls, milliseconds (CUDA events on the context's
 * stream).  Index with the NM_T_* constants. */
enum {
  NM_T_UPLOAD = 0,     /* host->device copies + layout kernels of nm_set_*                 */
  NM_T_INDEX = 1,      /* background grid-hash build                                       */
  NM_T_CANDIDATES = 2, /* count + scan + fill of candidate pairs                           */
  NM_T_SPLIT = 3,      /* exact Delaunay split of immersed quads (3D)                      */
  NM_T_CLIP = 4,       /* per-pair clipping + quadrature + local coupling vector           */
  NM_T_MERGE = 5,      /* per-immersed-cell merge through constraint lines -> rows of C    */
  NM_T_TRANSPOSE = 6,  /* C -> stored orientation CSR                                      */
  NM_T_SPMV = 7,       /* last nm_spmv, stored orientation                                 */
  NM_T_SPMV_T = 8,     /* last nm_spmv, transposed                                         */
  NM_T_DOWNLOAD = 9,   /* device->host copies of the last nm_get_*                         */
  /* single kernels, events directly around the launch */
  NM_T_K_CLIP = 10,        /* clip_moments3_kernel (3D) / clip_local_kernel (2D)           */
  NM_T_K_CAND_COUNT = 11,  /* cand_probe                                                   */
  NM_T_K_CAND_FILL = 12,   /* cand_place                                                   */
  NM_T_K_MERGE = 13,       /* merge_rows_fast (general path: merge_rows_hash)              */
  NM_T_COUNT = 16
};

int nm_version(void);

/* Create / destroy a context bound to CUDA device `device`. */
int nm_create(nm_ctx **ctx, int device);
int nm_destroy(nm_ctx *ctx);
const char *nm_last_error(const nm_ctx *ctx);

/* Background grid = what the wrapper reads from `space_cache` / `space_dh` /
 * `space_constraints`:
 *   verts  [n_cells][2^d][d]  mapping.get_vertices(cell), deal.II lexicographic order
 *                              (intersections.h:39-58 -- the CGAL re-ordering is not needed)
 *   dofs   [n_cells][2^d]     cell->get_dof_indices()      (coupling_utilities.h:109-115,276-283)
 *   constraint lines of the closed AffineConstraints object, CSR over the ascending list
 *   c_dof[n_constrained]: masters c_master[c_ptr[i]..c_ptr[i+1]) with weights c_weight
 *   (a line without masters = Dirichlet row).            (coupling_utilities.h:127-129,285-287)
 * Scope: FE_Q(1) on axis-aligned cells (every reference config).
 * Cells that are not axis-aligned boxes are rejected with NM_ERR_UNSUPPORTED. */
int nm_set_background(nm_ctx *ctx, int spacedim, uint64_t n_cells, const double *verts, const uint32_t *dofs,
                      uint64_t n_dofs, uint64_t n_constrained, const uint32_t *c_dof, const uint64_t *c_ptr,
                      const uint32_t *c_master, const double *c_weight, int where);

/* Same, with the cells given as boxes lo[n][d], hi[n][d] (one third of the transfer). */
int nm_set_background_boxes(nm_ctx *ctx, int spacedim, uint64_t n_cells, const double *lo, const double *hi,
                            const uint32_t *dofs, uint64_t n_dofs, uint64_t n_constrained, const uint32_t *c_dof,
                            const uint64_t *c_ptr, const uint32_t *c_master, const double *c_weight, int where);

/* DoF indices and constraint lines of the background, for callers that learn them after the
 * geometry (the reference's collect_quadratures_on_overlapped_grids sees only the two
 * GridTools::Cache objects; the DoFHandlers arrive with the sparsity / matrix calls).  `dofs`
 * may be NULL in nm_set_background* and nm_set_immersed when these follow.  Results of
 * nm_intersect stay valid. */
int nm_set_background_dofs(nm_ctx *ctx, const uint32_t *dofs, uint64_t n_dofs, uint64_t n_constrained, const uint32_t *c_dof,
                           const uint64_t *c_ptr, const uint32_t *c_master, const double *c_weight, int where);
int nm_set_immersed_dofs(nm_ctx *ctx, const uint32_t *dofs, uint32_t dofs_per_cell, uint64_t n_dofs, int where);

/* Immersed grid (or this rank's contiguous shard of it) = `immersed_cache` / `immersed_dh`:
 *   verts [n_cells][2^dim][spacedim] deal.II order, dofs [n_cells][dofs_per_cell].
 * Scope: dim = spacedim-1, FE_DGQ(0) (dofs_per_cell == 1, DoFs not shared between cells),
 * no immersed constraints (never filled in the reference: 2d3d/lagrange_multiplier.cc:149).
 * Also dim = spacedim = 2, the reference's codimension-0 instantiation <2,2,2> (src/coupling_utilities.cc:72-102,
 * src/intersections.cc:307-355): convex quads in the plane of the background (a quad that is not convex is refused with
 * NM_ERR_UNSUPPORTED); the pairs are background cell n immersed quad, the quadrature is QGaussSimplex<2>(degree) on the fan
 * triangles of the clipped polygon, the rest of the path (acceptance, sparsity, Q1 x P0 matrix, products) is the same. */
int nm_set_immersed(nm_ctx *ctx, int dim, int spacedim, uint64_t n_cells, const double *verts, const uint32_t *dofs,
                    uint32_t dofs_per_cell, uint64_t n_dofs, int where);

/* Replaces NonMatching::collect_quadratures_on_overlapped_grids<dim0,dim1,spacedim>
 * (include/coupling_utilities.h:508-515, src/coupling_utilities.cc:22-69) for <2,1,2> and
 * <3,2,3>: candidate pairs, exact intersection, quadrature of `degree` (the reference passes
 * it on as QGaussSimplex's n_points_1D: intersections.cc:695,776,812), acceptance by
 * sum(weights) > tol.  degree == 0 -> NM_ERR_INVALID (the reference AssertThrows).
 * Subdivision: the reference integrates over CGAL's pieces (tetrahedra x triangles), this library over the fan of the
 * clipped polygon.  With a rule that is exact for the integrand (degree >= 3 in 3D, >= 2 in 2D for Q1 x P0) both give the
 * same matrix up to rounding.  With a LOWER rule the pair set and the pattern are still the reference's, but the values
 * are that rule on a different subdivision: they differ from a CGAL run by the quadrature error of either side. */
int nm_intersect(nm_ctx *ctx, unsigned degree, double tol, nm_counts *counts);

/* flags[b] = 1 iff the closed bounding box of background cell b meets the bounding box of some
 * immersed cell: the refinement flagging loop of the drivers' adjust_grids()
 * (examples/lagrange_multiplier/1d2d/lagrange_multiplier.cc:361-470, same R-tree query as
 * src/coupling_utilities.cc:53-55).  Needs only nm_set_background + nm_set_immersed. */
int nm_flag_overlapped_background(nm_ctx *ctx, uint8_t *flags, int where);

/* Results of nm_intersect, ordered by (immersed cell, background cell) ascending.
 * Any output pointer may be NULL. */
int nm_get_candidates(nm_ctx *ctx, uint32_t *immersed, uint32_t *background, int where);
int nm_get_pairs(nm_ctx *ctx, uint32_t *immersed, uint32_t *background, double *sum_w, int where);
/* Quadrature<spacedim> of every accepted pair: q_offset[n_accepted+1], points
 * [n_qpoints][spacedim], weights[n_qpoints]  (what the reference returns in its tuples). */
int nm_get_quadrature(nm_ctx *ctx, uint64_t *q_offset, double *points, double *weights, int where);
/* Per-immersed-quad split code (bit t: triangle omitting vertex t is used; 0x80: tie). */
int nm_get_split_codes(nm_ctx *ctx, uint8_t *codes, int where);

/* Replaces create_coupling_sparsity_pattern_with_exact_intersections
 * (include/coupling_utilities.h:28-152): pattern of the stored matrix with
 * keep_constrained_entries = true. */
int nm_build_sparsity(nm_ctx *ctx, uint64_t *nnz);

/* Replaces create_coupling_mass_matrix_with_exact_intersections
 * (include/coupling_utilities.h:154-304) on the context's own intersection data (fused
 * path: the quadrature points never leave the device). */
int nm_assemble(nm_ctx *ctx);

/* The same entry point fed with the caller's `cells_and_quads` vector, exactly as the
 * reference signature passes it: inverse mapping of the given points to the background
 * reference cell, shape-function products, scatter (coupling_utilities.h:238-289).
 * pair (immersed[i], background[i]) owns points q_offset[i] .. q_offset[i+1]. */
int nm_assemble_from_quadrature(nm_ctx *ctx, uint64_t n_pairs, const uint32_t *immersed, const uint32_t *background,
                                const uint64_t *q_offset, const double *points, const double *weights, int where);

/* CSR of the matrix.  transpose = 0: stored orientation (rows = space DoFs, n_space+1
 * row pointers); transpose = 1: rows = immersed DoFs.  Columns ascending in every row.
 * val may be NULL (pattern only). */
int nm_get_csr(nm_ctx *ctx, int transpose, uint64_t *rowptr, uint32_t *col, double *val, int where);

/* The stored orientation without its empty rows -- what a caller needs to replay the reference's
 * dsp.add_entries(row, ...) / matrix.add(row, ...) calls (coupling_utilities.h:127-129, :285-287), and
 * the only form that stays small on a shard of a multi-GPU run, where the global dof count is far
 * larger than what one rank touches:
 *   row_ids[n_rows]  global space dof of every row that holds entries, ascending
 *   row_len[n_rows]  number of entries of that row
 *   col / val [nnz]  the rows back to back, columns (immersed dofs) ascending within a row.
 * Any pointer may be NULL; *n_rows is always set, so a first call with NULL arrays sizes the buffers
 * (nnz comes from nm_build_sparsity).  where = NM_HOST_ASYNC queues the copies on a separate stream
 * and returns: the (pinned) host buffers are complete after nm_sync_downloads(); the next assembly on
 * this context may be started in between -- its uploads and kernels overlap the read-back, and it waits
 * for the copies only where it would overwrite the matrix. */
int nm_get_csr_compact(nm_ctx *ctx, uint64_t *n_rows, uint32_t *row_ids, uint32_t *row_len, uint32_t *col, double *val, int where);
int nm_sync_downloads(nm_ctx *ctx);

/* transpose = 0: y[n_space] = A x[n_immersed]      (reference: Ct.vmult,  2d3d/...cc:627)
 * transpose = 1: y[n_immersed] = A^T x[n_space]    (reference: C = transpose_operator(Ct), :628)
 * With a sharded immersed grid y (transpose=0) is this rank's partial sum -- all-reduce it;
 * y (transpose=1) is non-zero only in this rank's immersed DoFs. */
int nm_spmv(nm_ctx *ctx, int transpose, const double *x, double *y, int where);

/* The same products in LOCAL numbering, for a consumer that is distributed like the reference's
 * (TrilinosWrappers::MPI::Vector, 2d3d/lagrange_multiplier.cc:627-643) and holds only the entries a
 * rank works on.  NM_LOCAL_IMMERSED: entry m belongs to immersed cell m of this context's (shard of
 * the) immersed grid; NM_LOCAL_ROWS: entry r belongs to row_ids[r] of nm_get_csr_compact.
 *   transpose = 0: y[n_rows]  = A x[n_cells]     transpose = 1: y[n_cells] = A^T x[n_rows]
 * No global-length vector crosses the host/device boundary and nothing global-length is zeroed. */
enum { NM_LOCAL_IMMERSED = 0, NM_LOCAL_ROWS = 1 };
int nm_spmv_local(nm_ctx *ctx, int transpose, const double *x, double *y, int where);
/* Scatter a local vector into a global-numbered DEVICE vector (entries outside the local set are left
 * untouched) / gather the local entries out of one.  `where_local` places the local vector. */
/* Both products of one Schur-complement iteration in local numbering, y_ct = Ct x_ct and y_c = C x_c (sizes as in
 * nm_spmv_local: x_ct, y_c one entry per immersed cell of this context; y_ct, x_c one per stored row).  With NM_HOST the
 * four transfers run on separate copy streams, so that the upload of x_c overlaps the download of y_ct (pinned buffers;
 * the two directions of the link are used at once); with NM_DEVICE it is the two nm_spmv_local calls. */
int nm_spmv_local_pair(nm_ctx *ctx, const double *x_ct, double *y_ct, const double *x_c, double *y_c, int where);
int nm_local_to_global(nm_ctx *ctx, int which, const double *local, int where_local, double *global_dev);
int nm_global_to_local(nm_ctx *ctx, int which, const double *global_dev, double *local, int where_local);

/* Ct.vmult fused with its reduction across GPUs (stored orientation, device pointers only):
 * every non-empty row result is ADDED into each of the `n_targets` vectors y_targets[i][n_space]
 * with a system-scope reduction -- pass this GPU's own result vector and the peer-mapped result
 * vectors of the other ranks (CUDA IPC / symmetric memory over NVLink), or, with multicast = 1,
 * ONE multicast address on which the NVSwitch applies the add to every rank's copy (multimem.red).
 * Contract: all target vectors are zero before any rank calls, and hold the full product once all
 * ranks' calls have returned (synchronise the ranks before and after).  Replaces nm_spmv(0) +
 * all-reduce: only the rows a rank owns cross the links. */
int nm_spmv_push(nm_ctx *ctx, const double *x, int n_targets, double *const *y_targets, int multicast);

/* The same fused product with reduce-scatter semantics, for a consumer that is distributed by background-dof ownership
 * like the reference's (Ct*lambda is a TrilinosWrappers::MPI::Vector, 2d3d/lagrange_multiplier.cc:627-643): rank k owns
 * the dofs [k*rows_per_rank, (k+1)*rows_per_rank) (the last rank up to n_space) and every row result is added only into
 * y_targets[owner][dof] -- this GPU's own vector or a peer-mapped one.  Each vector has n_space entries; after all
 * ranks' calls the owned range of a rank's vector holds the product.  Same zero-before / synchronise-after contract. */
int nm_spmv_push_owned(nm_ctx *ctx, const double *x, int n_ranks, double *const *y_targets, uint64_t rows_per_rank);

/* The same distributed result by PULLING, bit-reproducible: every rank writes the results of its stored rows (and their
 * dof ids, ascending) into a send buffer its peers can read; every owner then adds, source rank after source rank, the
 * segment of each buffer that falls into its dof range -- coalesced reads over NVLink and plain adds in a fixed order
 * instead of one 8-byte remote atomic per row (non_matching_test_suite_b200/distributed.py::FusedCtVmult.vmult_owned_pull
 * shows the set-up: symmetric memory for the send buffers, rank barriers between the steps).  Device pointers.
 *   nm_spmv_rows      y_rows[r] = (A x)[dof of stored row r] for the rows with entries; optionally their dof ids
 *   nm_owner_offsets  offsets[k] = first stored row owned by rank k (k = 0..n_ranks), dofs [k rows_per_rank, ...)
 *   nm_pull_rows      y_owned[id - first_row] += val over the segment [src_offsets[me], src_offsets[me + 1]) of a
 *                     (peer's) send buffer; an id outside [first_row, first_row + n_owned) -> NM_ERR_INVALID */
int nm_spmv_rows(nm_ctx *ctx, const double *x, double *y_rows, uint32_t *row_ids, uint64_t *n_rows);
int nm_owner_offsets(nm_ctx *ctx, int n_ranks, uint64_t rows_per_rank, uint64_t *offsets);
int nm_pull_rows(nm_ctx *ctx, const uint64_t *src_offsets, int me, const uint32_t *src_row_ids, const double *src_vals, uint64_t first_row,
                 uint64_t n_owned, double *y_owned);

/* ---- the whole assembly from HOST arrays in one call, uploads overlapped with the computation ----
 * Same result as nm_set_background_boxes + nm_set_immersed + nm_intersect + nm_build_sparsity + nm_assemble with
 * NM_HOST pointers, i.e. what collect_quadratures_on_overlapped_grids (code/src/coupling_utilities.cc:22-69) followed by
 * create_coupling_sparsity_pattern / create_coupling_mass_matrix_with_exact_intersections
 * (code/include/coupling_utilities.h:28-152, :154-304) compute, but all host->device copies are queued up front on a
 * copy stream and each phase waits only for the arrays it reads: the grid index is built while the immersed grid is
 * still in flight, candidates / clipping run while the DoF tables and constraint lines arrive.  With pinned host memory
 * the copies are asynchronous; with pageable memory the call is still correct, only not overlapped.
 * All pointers are host pointers and are not used after the call returns.  Read the matrix with nm_get_csr. */
typedef struct {
  int spacedim;                 /* 2 or 3 */
  uint64_t n_bg;                /* background cells (axis-aligned boxes) */
  const double *lo, *hi;        /* [n_bg][spacedim] */
  const uint32_t *bg_dofs;      /* [n_bg][2^spacedim] */
  uint64_t n_space_dofs;
  uint64_t n_constrained;       /* constraint lines as in nm_set_background */
  const uint32_t *c_dof;
  const uint64_t *c_ptr;
  const uint32_t *c_master;
  const double *c_weight;
  int dim;                      /* spacedim - 1 */
  uint64_t n_im;
  const double *im_verts;       /* [n_im][2^dim][spacedim] */
  const uint32_t *im_dofs;      /* [n_im] */
  uint64_t n_im_dofs;
  unsigned degree;              /* quadrature degree, > 0 */
  double tol;
  const double *edge;           /* optional, [n_bg]: CUBIC cells given by lower corner and edge length -- the upper corner is
                                 * fl(lo + edge) on every axis and `hi` is not read.  For grids where that holds bitwise
                                 * (hyper_cube refinements: dyadic coordinates), a third of the geometry bytes less to send.
                                 * NULL: boxes from lo / hi. */
} nm_host_problem;
int nm_assemble_host(nm_ctx *ctx, const nm_host_problem *problem, nm_counts *counts, uint64_t *nnz);

/* ---- higher polynomial degrees: FE_Q(p) on the background, FE_DGQ(q) on the immersed grid ----
 * The arithmetic of create_coupling_mass_matrix_with_exact_intersections for any tensor-product Lagrange pair
 * (code/include/coupling_utilities.h:238-289):
 *   xi_q = F_bg^-1(x_q) (:248-249), eta_q = F_im^-1(x_q) on the codimension-one cell (:251-254),
 *   local(i,j) += phi_i(xi_q) psi_j(eta_q) w_q (:258-274), distribute_local_to_global through the space constraints
 *   (:285-287), and the pattern of create_coupling_sparsity_pattern_with_exact_intersections with
 *   keep_constrained_entries = true (:100-139).
 * An element is described by its degree and its unit support points IN ITS OWN DOF ORDER (deal.II:
 * fe.get_unit_support_points()); the shape functions are the Lagrange polynomials of those points, so no numbering
 * convention is assumed.  Limits: degree <= 5, <= 64 background and <= 16 immersed DoFs per cell; background cells
 * axis-aligned as everywhere in this library; no immersed constraints.
 *   nm_set_finite_elements  after nm_set_background* / nm_set_immersed (their `dofs` may be NULL): DoF tables
 *                           [n_cells][dofs_per_cell] and the space constraint lines (as in nm_set_background)
 *   nm_assemble_fe          immersed == NULL: on the quadrature of the context's own nm_intersect (it never leaves the
 *                           device); otherwise on the caller's tuples, exactly as the reference signature passes them
 *                           (pair i = (immersed[i], background[i]), points q_offset[i] .. q_offset[i+1])
 * Afterwards nm_get_csr / nm_get_csr_compact / nm_spmv / nm_spmv_local work on the result; NM_LOCAL_IMMERSED then
 * numbers the immersed entries (cell, local dof) = cell * dofs_per_cell + local dof. */
typedef struct {
  unsigned degree;
  unsigned dofs_per_cell;
  const double *unit_support_points;   /* [dofs_per_cell][dim] (host); may be NULL for a one-dof element */
} nm_fe;
int nm_set_finite_elements(nm_ctx *ctx, const nm_fe *space_fe, const uint32_t *space_dofs, uint64_t n_space_dofs, uint64_t n_constrained,
                           const uint32_t *c_dof, const uint64_t *c_ptr, const uint32_t *c_master, const double *c_weight,
                           const nm_fe *immersed_fe, const uint32_t *immersed_dofs, uint64_t n_immersed_dofs, int where);
int nm_assemble_fe(nm_ctx *ctx, uint64_t n_pairs, const uint32_t *immersed, const uint32_t *background, const uint64_t *q_offset,
                   const double *points, const double *weights, int where, uint64_t *nnz);

/* ---- the Schur-complement solve around the GPU products (the rest of solve()) ----
 * Replaces PoissonLM::solve() (code/examples/lagrange_multiplier/2d3d/lagrange_multiplier.cc:613-659,
 * 1d2d/lagrange_multiplier.cc:675-730) with Ct = the context's assembled coupling matrix and C its transpose:
 *   K_inv = CG on K (ReductionControl inner_*), preconditioner = C K Ct + M applied by multiplication (:630),
 *   S = C K_inv Ct (:632), lambda = GMRES(S, ReductionControl outer_*) (C K_inv space_rhs - embedded_rhs) (:636-642),
 *   solution = K_inv (space_rhs - Ct lambda), then AffineConstraints::distribute with the context's constraint lines (:643,:658).
 * ReductionControl(n, tol, reduce) stops at |r| <= tol or |r| <= reduce |r_0| (deal.II); GMRES is restarted and left-
 * preconditioned with `gmres_basis` Krylov vectors (0 = deal.II's default, 28).  K (n_space x n_space, symmetric positive
 * definite as the drivers assemble it, constrained rows included) and M (n_immersed x n_immersed) are CSR matrices of the
 * caller; `inhomogeneity` (one value per constraint line, NULL = 0) is what distribute() adds.  The CG on K is
 * preconditioned with diag(K) instead of the reference's Trilinos AMG: the same K_inv up to inner_tol, other iteration counts.
 * All products with the coupling matrix are this library's SpMV kernels; every vector operation runs on the GPU. */
typedef struct {
  uint64_t n_space, n_immersed;
  const uint64_t *k_rowptr; const uint32_t *k_col; const double *k_val;
  const uint64_t *m_rowptr; const uint32_t *m_col; const double *m_val;
  const double *space_rhs, *embedded_rhs;
  const double *inhomogeneity;
  unsigned inner_max_it; double inner_tol, inner_reduce;
  unsigned outer_max_it; double outer_tol, outer_reduce;
  unsigned gmres_basis;
} nm_schur_problem;
typedef struct {
  unsigned outer_iterations;      /* reduction_control.last_step() of the GMRES on S */
  uint64_t inner_iterations;      /* CG iterations on K, summed over all applications of K_inv */
  double outer_residual;          /* preconditioned residual the GMRES stopped at */
  double lambda_norm_sqr;         /* lambda.norm_sqr(), printed by the drivers (:644) */
  float ms;                       /* device time of the whole solve */
} nm_schur_stats;
int nm_schur_solve(nm_ctx *ctx, const nm_schur_problem *problem, double *lambda, double *solution, nm_schur_stats *stats, int where);

/* ---- interface-penalisation (Nitsche) terms on a caller-supplied quadrature ----
 * Replaces the arithmetic of assemble_nitsche_with_exact_intersections (code/include/coupling_utilities.h:306-398) and
 * create_nitsche_rhs_with_exact_intersections (:400-481): for tuple p with background cell background[p] and points
 * q_offset[p]..q_offset[p+1]
 *   local(i,j) += coefficient[q] * (penalty / h) * phi_i(xi_q) * phi_j(xi_q) * weights[q]       (:366-370)
 *   rhs(i)     += coefficient[q] * (penalty / h) * phi_i(xi_q) * rhs_values[q] * weights[q]     (:447-451)
 * h = diameter of the background cell (:357,:429), xi_q = F^-1(x_q) on the axis-aligned Q1 cell; then
 * AffineConstraints::distribute_local_to_global, square version (:387-388): rows and columns resolved through the
 * constraint lines, |local(i,i)| (or the mean |diagonal| of the block if that is zero) added on the diagonal of every
 * constrained dof; vector version (:458-459): resolved rows only.
 * `coefficient` / `rhs_values`: values of the reference's Function objects at the points ([n_q]); coefficient may be
 * NULL (= 1).  `what`: NM_NITSCHE_MATRIX | NM_NITSCHE_RHS (rhs_values required for the latter).
 * Needs nm_set_background[_boxes] with dofs (or nm_set_background_dofs); no immersed grid and no nm_intersect.
 * The result is the CONTRIBUTION (to be added to the caller's system matrix / rhs): a CSR matrix n_space x n_space
 * with sorted columns holding every structural position the reference's add() would touch, and a dense vector. */
#define NM_NITSCHE_MATRIX 1
#define NM_NITSCHE_RHS 2
int nm_assemble_nitsche(nm_ctx *ctx, uint64_t n_pairs, const uint32_t *background, const uint64_t *q_offset, const double *points,
                        const double *weights, const double *coefficient, const double *rhs_values, double penalty, int what,
                        int where, uint64_t *nnz);
/* rowptr[n_space+1], col[nnz], val[nnz], rhs[n_space]; any pointer may be NULL (skipped). */
int nm_get_nitsche(nm_ctx *ctx, uint64_t *rowptr, uint32_t *col, double *val, double *rhs, int where);

/* Introspection (tests and the benchmark): which code path the last calls took. */
enum {
  NM_INFO_MERGE_PATH = 0,   /* matrix built by: 0 = single-pass hash merge, 1 = general two-pass merge        */
  NM_INFO_COMPACT_ROWS = 1, /* stored orientation keeps only touched rows internally (sharded grids)           */
  NM_INFO_STORED_ROWS = 2,  /* rows the stored orientation holds internally                                    */
  NM_INFO_NNZ = 3,
  NM_INFO_INDEX_KIND = 4,   /* candidate index: 1 = per-cell chains (any boxes), 2 = brick bitmasks (grid-aligned tree) */
  NM_INFO_HOST_SYNCS = 5,   /* host synchronisations with the stream since nm_create                           */
  NM_INFO_MERGE_SLOTS = 6   /* hash slots per immersed cell of the single-pass merge that built the matrix     */
};
int nm_get_info(nm_ctx *ctx, int key, uint64_t *value);

/* Measured FP64 throughput of this GPU: a register-resident chain of dependent-free DFMAs on every SM for about
 * `ms_target` milliseconds; *flops_per_s counts 2 flops per DFMA.  The denominator of the FP64 roofline fraction. */
int nm_measure_fp64_peak(nm_ctx *ctx, double ms_target, double *flops_per_s);

/* Device milliseconds per phase (NM_T_*), `n` entries copied. */
int nm_get_timings(nm_ctx *ctx, float *ms, int n);
/* Number of this library's kernel launches since nm_create (cub's internal kernels excluded). */
int nm_get_launch_count(nm_ctx *ctx, uint64_t *n);
/* Bytes of device memory currently held by the context. */
int nm_device_bytes(nm_ctx *ctx, uint64_t *bytes);
/* The CUDA stream (cudaStream_t) the context launches on, for callers that time with events. */
int nm_get_stream(nm_ctx *ctx, void **stream);

#ifdef __cplusplus
}
#endif
#endif /* NMGPU_H */
