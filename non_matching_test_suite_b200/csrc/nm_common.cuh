// Shared host-side state and helpers of libnmgpu.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nmgpu.h"

#define NM_SM_COUNT_B200 148
#define NM_STATS_BLOCKS (8 * NM_SM_COUNT_B200)  // grid of the extent-statistics kernels: 8 blocks per SM

// ---------------------------------------------------------------------------------------
// device buffer that only grows; reused across calls so that a step does no cudaMalloc
// ---------------------------------------------------------------------------------------
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  bool borrowed = false;   // p is the caller's device array (NM_DEVICE_INPLACE): never freed, never written by the library
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Background grid-hash slot: one occupied grid cell of one level.
struct HashSlot {
  unsigned long long key;  // packed (level, ix, iy, iz); ~0 = empty
  uint32_t start;          // head of the chain of entries (a background cell id for a grid-aligned tree)
  uint32_t count;          // length of the chain
};

struct GridIndex {
  int d = 0;
  double org[3] = {0, 0, 0};
  double brick_org[3] = {0, 0, 0};   // origin shared by the grids of all levels (brick index)
  double g0 = 0;            // spacing of level 0
  uint32_t level_mask = 0;  // bit k: level k holds cells
  uint32_t n_entries = 0;
  uint32_t hash_cap = 0;    // power of two
};

struct nm_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;           // host->device copies of nm_assemble_host
  cudaStream_t down_stream = nullptr;           // device->host copies of nm_get_csr_compact(NM_HOST_ASYNC)
  cudaEvent_t ev_down_ready = nullptr, ev_down_done = nullptr;
  // the copies of an asynchronous read-back are issued by a helper thread: the caller gets control back at once even
  // where the driver makes a multi-GB cudaMemcpyAsync wait (measured: 60 ms inside the call for 3.4 GB)
  std::thread down_thread;
  std::atomic<int> down_issued{0};              // 1: the helper has queued its copies and recorded ev_down_done
  int down_rc = 0;
  bool download_pending = false;                // the next matrix build must wait for ev_down_done
  uint64_t n_syncs = 0;                         // host synchronisations with the stream (counter read-backs) since nm_create
  int index_kind = 0;                           // grid index in use: 0 = none yet, 1 = per-cell chains, 2 = brick bitmasks
  cudaEvent_t ev_copy[3] = {nullptr, nullptr, nullptr};
  std::string err;
  int sm_count = NM_SM_COUNT_B200;

  // ---- background ----
  int spacedim = 0;
  uint64_t n_bg = 0, n_space_dofs = 0, n_constrained = 0, n_cmasters = 0;
  DevBuf bg_box;    // [n_bg][8] doubles: lo[0..2] pad hi[0..2] pad  (64 B, two sectors)
  DevBuf bg_dofs;   // [n_bg][2^d] u32
  DevBuf line_of;   // [n_space_dofs] int32: constraint line of a DoF or -1
  DevBuf c_ptr;     // [n_constrained+1] u64
  DevBuf c_master;  // u32
  DevBuf c_weight;  // f64
  bool bg_set = false, bg_dofs_set = false, index_valid = false;
  bool bg_stats_ready = false;  // idx_tmp holds the extent statistics of the current boxes (taken by the lo/hi layout kernel)

  // ---- immersed ----
  int dim1 = 0;
  uint64_t n_im = 0, n_im_dofs = 0;
  DevBuf im_verts;    // [n_im][2^dim1][spacedim] f64 (AoS, cell-major); codimension 0: [n_im][2][2] = the diagonal of each quad's box
  bool codim0 = false; // dim1 == spacedim == 2 (the reference's <2,2,2>): the quads themselves are in im_poly
  DevBuf im_poly;     // [n_im][4][2] f64, deal.II vertex order
  DevBuf im_dofs;     // [n_im] u32
  DevBuf im_split;    // [n_im] u8 (3D)
  DevBuf im_measure;  // [n_im] f64
  bool im_set = false, im_dofs_set = false;
  bool dofs_checked = false;   // both dof tables verified against the declared dof counts (reset by every upload)

  // ---- grid hash ----
  GridIndex gi;
  DevBuf idx_next;      // chain of the entries of one grid cell (next[entry id])
  DevBuf idx_extra_bg;  // background id of entry ids >= n_bg
  DevBuf idx_hash, idx_tmp;
  DevBuf brk_hdr, brk_bit;  // brick index: header (levels, origins, flags), bit of every cell inside its brick
  bool brick_roomy = false;        // the brick table was sized so that it cannot overflow
  bool force_chain_index = false;  // NMGPU_INDEX=chains: skip the brick index (testing)

  // ---- intersection results ----
  unsigned degree = 0;
  double tol = 0;
  nm_counts counts;
  DevBuf cand_tmp;    // [n_im][CAP] u32: hits of the single probing pass
  DevBuf cand_ptr;    // [n_im+1] u64
  DevBuf cand_im;     // [n_cand] u32
  DevBuf cand_bg;     // [n_cand] u32
  DevBuf cand_sumw;   // [n_cand] f64
  DevBuf cand_local;  // [n_cand][2^d] f64
  DevBuf cand_nq;     // [n_cand] u16
  DevBuf cand_acc;    // [n_cand] u8: pair accepted (sum of weights > tol)
  DevBuf counters;    // device counters (u64 x 8)
  bool candidates_only = false;  // candidate list computed without clipping (nm_flag_overlapped_background)
  bool intersect_valid = false;
  bool force_general_merge = false;      // NMGPU_FORCE_GENERAL_MERGE=1: skip the single-pass merge (testing)
  bool force_quadrature_kernel = false;  // NMGPU_FORCE_QUADRATURE=1: per-point kernel even where moments apply
  bool local_from_user = false;  // cand_local was filled by nm_assemble_from_quadrature

  // ---- matrix ----
  // C orientation (rows = immersed cells of this shard), non-compact row storage
  DevBuf c_rowstart;  // [n_im+1] u64 (capacity offsets)
  DevBuf c_rowlen;    // [n_im] u32
  DevBuf c_col;       // u32
  DevBuf c_val;       // f64
  uint64_t nnz = 0;
  uint64_t n_crows = 0;             // rows of C and the immersed dof of each (set by whichever path built the matrix)
  DevBuf *crow_dofs = nullptr;
  // stored orientation (rows = space dofs), compact CSR
  DevBuf a_rowptr32;  // same as u32 when nnz < 2^32
  bool a_rowptr32_valid = false;
  DevBuf a_rowptr;  // [n_space_dofs+1] u64
  DevBuf a_col;     // u32 (immersed dof)
  DevBuf a_val;     // f64
  DevBuf a_cursor;  // scratch
  DevBuf t_rec;     // (value, column) records between the transpose scatter and its row sort
  // compact-row mode (sharded grids: the global dof count is far larger than what this rank touches): the stored
  // orientation keeps only the rows with entries; a_rowptr has n_local_rows + 1 entries and row r is global dof
  // row_ids[r].  row_bits = bitmap of touched dofs, row_prefix[w] = touched dofs before word w (rank queries).
  bool compact_rows = false;
  int compact_rows_env = -1;  // NMGPU_COMPACT_ROWS: 0 / 1 force, -1 automatic
  uint64_t n_local_rows = 0;
  DevBuf row_bits, row_prefix, row_ids;
  // list of the non-empty rows for the compact read-back / local numbering (built on demand, nm_nonempty_rows)
  DevBuf out_ids, out_len;
  uint64_t n_out_rows = 0;
  bool out_rows_valid = false;
  bool matrix_valid = false;
  int merge_path = -1;   // which merge built the matrix: 0 = single-pass hash merge, 1 = general two-pass path
  int merge_slots = 0;   // hash slots per cell of the single-pass merge that built it (128 or 256 in 3D, 64 in 2D)
  bool merge_large_table = false;   // the small table overflowed on these grids: start with the large one

  // ---- interface-penalisation terms (nm_nitsche.cu): square matrix on the space dofs + right-hand side ----
  DevBuf nit_cnt, nit_keys, nit_vals, nit_rkeys, nit_rvals;
  DevBuf nit_rowptr, nit_col, nit_val, nit_rhs;
  uint64_t nit_nnz = 0;
  bool nit_has_matrix = false, nit_has_rhs = false;

  // ---- general finite elements (nm_fe.cu): FE_Q(p) background x FE_DGQ(q) immersed ----
  bool fe_general = false;          // nm_set_finite_elements was called: the matrix comes from nm_assemble_fe
  int fe_nb = 0, fe_ni = 0;         // dofs per background / immersed cell
  std::vector<unsigned char> fe;    // the device constant block (node tables, dof -> node maps)
  DevBuf fe_bg_dofs, fe_im_dofs;    // [n_bg][fe_nb], [n_im][fe_ni] u32
  DevBuf fe_pairs, fe_qoff, fe_pts, fe_w, fe_recoff, fe_keys, fe_vals;

  // ---- scratch ----
  DevBuf scan_tmp, scratch_a, scratch_b, stage_x, stage_y, stage_x2, stage_y2, h2d_stage;
  DevBuf solver_buf;          // matrices, right-hand sides and work vectors of nm_schur_solve
  DevBuf clip_lut;            // 256-entry table of the section algorithm (nm_clip_core.cuh)
  DevBuf glob_im, glob_dof;   // global-numbered device scratch vectors of the products in local numbering

  // ---- timing ----
  uint64_t n_launches = 0;  // kernels of this library launched so far (cub's internal ones not counted)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
  float t_ms[NM_T_COUNT];
};

// ---------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------
int nm_fail(nm_ctx *c, int code, const char *fmt, ...);

#define NM_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return nm_fail(ctx, NM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// host waits for the context's stream (counted: every one of them drains the GPU between two phases)
#define NM_SYNC(c)                                      \
  do {                                                  \
    ++(c)->n_syncs;                                     \
    NM_CUDA(cudaStreamSynchronize((c)->stream));        \
  } while (0)

#define NM_TRY(call)            \
  do {                          \
    int rc__ = (call);          \
    if (rc__ != NM_OK) return rc__; \
  } while (0)

int nm_ensure(nm_ctx *ctx, DevBuf &b, size_t bytes);
int nm_upload(nm_ctx *ctx, DevBuf &b, const void *src, size_t bytes, int where);
int nm_download(nm_ctx *ctx, void *dst, const void *src, size_t bytes, int where);

struct PhaseTimer {
  nm_ctx *c;
  int phase;
  bool accumulate;
  PhaseTimer(nm_ctx *ctx, int ph, bool acc = false) : c(ctx), phase(ph), accumulate(acc) { cudaEventRecord(c->ev0, c->stream); }
  void stop()
  {
    cudaEventRecord(c->ev1, c->stream);
    cudaEventSynchronize(c->ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    if (accumulate) c->t_ms[phase] += ms; else c->t_ms[phase] = ms;
  }
};

// launch-stream accessor that also counts the launch
#define NM_LS(c) (++(c)->n_launches, (c)->stream)

// events around ONE kernel launch (independent of the phase timers)
struct KernelTimer {
  nm_ctx *c;
  int slot;
  KernelTimer(nm_ctx *ctx, int s) : c(ctx), slot(s) { cudaEventRecord(c->ev2, c->stream); }
  void stop()
  {
    cudaEventRecord(c->ev3, c->stream);
    cudaEventSynchronize(c->ev3);
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev2, c->ev3);
    c->t_ms[slot] = ms;
  }
};

static inline unsigned nm_blocks(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

// ---- phases implemented in the other translation units ----
int nm_index_build(nm_ctx *ctx);
int nm_candidates_run(nm_ctx *ctx);
int nm_split_run(nm_ctx *ctx);
int nm_clip_run(nm_ctx *ctx);
int nm_quadrature_emit(nm_ctx *ctx, uint64_t *q_offset, double *points, double *weights, int where);
int nm_local_from_quadrature(nm_ctx *ctx, uint64_t n_pairs, const uint32_t *im, const uint32_t *bg, const uint64_t *q_offset,
                             const double *points, const double *weights, int where);
int nm_matrix_build(nm_ctx *ctx);
int nm_transpose_build(nm_ctx *ctx);
int nm_nitsche_run(nm_ctx *ctx, uint64_t n_pairs, const uint32_t *bg, const uint64_t *q_offset, const double *points,
                   const double *weights, const double *coefficient, const double *rhs_values, double penalty, int what, int where);
#define NM_SPMV_LOCAL_OUT 1
int nm_spmv_run(nm_ctx *ctx, int transpose, const double *x_dev, double *y_dev, int flags = 0);
int nm_stored_rowptr_global(nm_ctx *ctx, uint64_t *rowptr_dev);
int nm_spmv_push_owned_run(nm_ctx *ctx, const double *x, int n_ranks, double *const *targets, uint64_t rows_per_rank);
int nm_quad2d_prepare(nm_ctx *ctx);
int nm_nonempty_rows(nm_ctx *ctx, uint64_t *n_rows_out, const uint32_t **ids_dev, const uint32_t **len_dev);
int nm_local_map(nm_ctx *ctx, int which, int to_global, const double *src, double *dst);
int nm_wait_download_issued(nm_ctx *ctx);
int nm_constrained_dofs(nm_ctx *ctx, uint32_t *out_dev);
int nm_exclusive_scan_u64(nm_ctx *ctx, const uint64_t *in, uint64_t *out, uint64_t n);
int nm_exclusive_scan_u32_to_u64(nm_ctx *ctx, const uint32_t *in, uint64_t *out, uint64_t n);
int nm_exclusive_scan_u32_to_u64_and_u32(nm_ctx *ctx, const uint32_t *in, uint64_t *out, uint32_t *out32, uint64_t n);
