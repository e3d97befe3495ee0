// Times the reference's own three-call flow
//   collect_quadratures_on_overlapped_grids -> create_coupling_sparsity_pattern_with_exact_intersections ->
//   create_coupling_mass_matrix_with_exact_intersections
// (code/examples/lagrange_multiplier/2d3d/lagrange_multiplier.cc:750-753, 488-490, 599-602) through the drop-in header
// include/coupling_utilities.h on the deal.II stub, next to the direct C-ABI path (nm_assemble_host + nm_get_csr_compact)
// on the same arrays.  Input: the binary grid file bench.py writes.  Output: one JSON line.
#include <coupling_utilities.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

using namespace dealii;
using Clock = std::chrono::steady_clock;
static double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

template <typename T> static std::vector<T> rd(std::ifstream &in, std::size_t n)
{
  std::vector<T> v(n);
  if (n) in.read(reinterpret_cast<char *>(v.data()), std::streamsize(n * sizeof(T)));
  return v;
}

template <int dim0, int dim1, int spacedim> int run(std::ifstream &in)
{
  std::uint64_t hdr[6];
  in.read(reinterpret_cast<char *>(hdr), sizeof(hdr));
  const std::uint64_t n_bg = hdr[0], n_space_dofs = hdr[1], n_im = hdr[2], n_im_dofs = hdr[3], n_lines = hdr[4], n_masters = hdr[5];
  constexpr unsigned nv0 = 1u << dim0, nv1 = 1u << dim1;
  const auto lo = rd<double>(in, n_bg * spacedim), hi = rd<double>(in, n_bg * spacedim);
  const auto dofs = rd<std::uint32_t>(in, n_bg * nv0);
  const auto c_dof = rd<std::uint32_t>(in, n_lines);
  const auto c_ptr = rd<std::uint64_t>(in, n_lines + 1);
  const auto c_master = rd<std::uint32_t>(in, n_masters);
  const auto c_weight = rd<double>(in, n_masters);
  const auto im_verts = rd<double>(in, n_im * nv1 * spacedim);
  const auto im_dofs = rd<std::uint32_t>(in, n_im);
  if (!in) { std::cerr << "short grid file\n"; return 2; }

  Triangulation<dim0, spacedim> space;
  Triangulation<dim1, spacedim> immersed;
  DoFHandler<dim0, spacedim> space_dh;
  DoFHandler<dim1, spacedim> immersed_dh;
  AffineConstraints<double> constraints, immersed_constraints;
  space.cells.resize(n_bg);
  for (std::uint64_t c = 0; c < n_bg; ++c)
    for (unsigned v = 0; v < nv0; ++v)
      for (int d = 0; d < spacedim; ++d) space.cells[c][v][d] = ((v >> d) & 1u) ? hi[c * spacedim + d] : lo[c * spacedim + d];
  space_dh.fe.dpc = nv0;
  space_dh.table.assign(dofs.begin(), dofs.end());
  for (std::uint64_t l = 0; l < n_lines; ++l) {
    auto &line = constraints.lines[c_dof[l]];
    for (std::uint64_t k = c_ptr[l]; k < c_ptr[l + 1]; ++k) line.emplace_back(c_master[k], c_weight[k]);
  }
  immersed.cells.resize(n_im);
  for (std::uint64_t c = 0; c < n_im; ++c)
    for (unsigned v = 0; v < nv1; ++v)
      for (int d = 0; d < spacedim; ++d) immersed.cells[c][v][d] = im_verts[(c * nv1 + v) * spacedim + d];
  immersed_dh.fe.dpc = 1;
  immersed_dh.table.assign(im_dofs.begin(), im_dofs.end());
  space_dh.tria = &space; space_dh.ndofs = n_space_dofs;
  immersed_dh.tria = &immersed; immersed_dh.ndofs = n_im_dofs;
  Mapping<dim0, spacedim> space_mapping;
  Mapping<dim1, spacedim> immersed_mapping;
  GridTools::Cache<dim0, spacedim> space_cache(space, space_mapping);
  GridTools::Cache<dim1, spacedim> immersed_cache(immersed, immersed_mapping);

  double t_collect = 0, t_sparsity = 0, t_matrix = 0;
  std::size_t n_pairs = 0, n_q = 0, nnz_dsp = 0, nnz_mat = 0;
  double checksum = 0;
  for (int rep = 0; rep < 2; ++rep) {   // the first pass allocates (device buffers, host vectors); the second is reported
    auto t0 = Clock::now();
    const auto cells_and_quads = NonMatching::collect_quadratures_on_overlapped_grids(space_cache, immersed_cache, 3, 1e-15);
    t_collect = ms_since(t0);
    t0 = Clock::now();
    DynamicSparsityPattern dsp(n_space_dofs, n_im_dofs);
    NonMatching::create_coupling_sparsity_pattern_with_exact_intersections(cells_and_quads, space_dh, immersed_dh, dsp, constraints,
                                                                          ComponentMask(), ComponentMask(), immersed_constraints);
    t_sparsity = ms_since(t0);
    t0 = Clock::now();
    StubFlatMatrix<double> matrix(n_space_dofs, n_im_dofs);
    NonMatching::create_coupling_mass_matrix_with_exact_intersections(space_dh, immersed_dh, cells_and_quads, matrix, constraints,
                                                                     ComponentMask(), ComponentMask(), space_mapping, immersed_mapping,
                                                                     immersed_constraints);
    t_matrix = ms_since(t0);
    n_pairs = cells_and_quads.size();
    n_q = 0;
    for (const auto &t : cells_and_quads) n_q += std::get<2>(t).size();
    nnz_dsp = 0;
    for (const auto &r : dsp.entries) nnz_dsp += r.size();
    nnz_mat = matrix.v_.size();
    checksum = 0;
    for (const double v : matrix.v_) checksum += v;
  }

  // the direct path on the same arrays: one call from host arrays + compact read-back, and the quadrature read-back alone
  nm_ctx *ctx = nullptr;
  if (nm_create(&ctx, 0) != NM_OK) return 3;
  nm_host_problem P;
  std::memset(&P, 0, sizeof(P));
  P.spacedim = spacedim; P.n_bg = n_bg; P.lo = lo.data(); P.hi = hi.data(); P.bg_dofs = dofs.data(); P.n_space_dofs = n_space_dofs;
  P.n_constrained = n_lines; P.c_dof = c_dof.data(); P.c_ptr = c_ptr.data(); P.c_master = c_master.data(); P.c_weight = c_weight.data();
  P.dim = dim1; P.n_im = n_im; P.im_verts = im_verts.data(); P.im_dofs = im_dofs.data(); P.n_im_dofs = n_im_dofs; P.degree = 3; P.tol = 1e-15;
  double t_direct = 0, t_quad = 0, direct_sum = 0;
  std::uint64_t nnz = 0;
  for (int rep = 0; rep < 2; ++rep) {
    auto t0 = Clock::now();
    nm_counts counts;
    if (nm_assemble_host(ctx, &P, &counts, &nnz) != NM_OK) { std::cerr << nm_last_error(ctx) << "\n"; return 4; }
    std::uint64_t n_rows = 0;
    nm_get_csr_compact(ctx, &n_rows, nullptr, nullptr, nullptr, nullptr, NM_HOST);
    std::vector<std::uint32_t> ids(n_rows), len(n_rows), col(nnz);
    std::vector<double> val(nnz);
    if (nm_get_csr_compact(ctx, &n_rows, ids.data(), len.data(), col.data(), val.data(), NM_HOST) != NM_OK) return 5;
    t_direct = ms_since(t0);
    direct_sum = 0;
    for (const double v : val) direct_sum += v;
    t0 = Clock::now();
    std::vector<std::uint64_t> off(counts.n_accepted + 1);
    std::vector<double> pts(counts.n_qpoints * spacedim), w(counts.n_qpoints);
    if (nm_get_quadrature(ctx, off.data(), pts.data(), w.data(), NM_HOST) != NM_OK) return 6;
    t_quad = ms_since(t0);
  }
  nm_destroy(ctx);
  NonMatching::nmgpu_release();
  std::printf("{\"pairs\": %zu, \"qpoints\": %zu, \"nnz_pattern\": %zu, \"nnz_values\": %zu, \"collect_ms\": %.3f, \"sparsity_ms\": %.3f, "
              "\"matrix_ms\": %.3f, \"three_calls_ms\": %.3f, \"direct_assemble_host_plus_readback_ms\": %.3f, "
              "\"quadrature_readback_ms\": %.3f, \"sum_of_values_dropin\": %.17g, \"sum_of_values_direct\": %.17g}\n",
              n_pairs, n_q, nnz_dsp, nnz_mat, t_collect, t_sparsity, t_matrix, t_collect + t_sparsity + t_matrix, t_direct, t_quad, checksum,
              direct_sum);
  return 0;
}

int main(int argc, char **argv)
{
  if (argc < 2) { std::cerr << "usage: bench_dropin <grid.bin>\n"; return 2; }
  std::ifstream in(argv[1], std::ios::binary);
  std::int64_t spacedim = 0;
  in.read(reinterpret_cast<char *>(&spacedim), 8);
  try {
    return spacedim == 2 ? run<2, 1, 2>(in) : run<3, 2, 3>(in);
  } catch (const std::exception &e) {
    std::cerr << "exception: " << e.what() << "\n";
    return 1;
  }
}
