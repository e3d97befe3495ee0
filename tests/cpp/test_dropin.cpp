// Drives include/coupling_utilities.h (the reference's three entry points, same signatures)
// against the deal.II stub on grids written by tests/test_gpu_dropin.py; prints what a driver
// would see: number of pairs, total area, the pattern and the matrix.
#include <coupling_utilities.h>

#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>

using namespace dealii;

// the Function objects a Nitsche driver passes (interface_penalisation/1d2d/smooth_disk.cc:435-450 uses
// ConstantFunction(2.0) and its Solution class); non-constant here so that the per-point values matter
template <int spacedim> struct Coefficient : Function<spacedim> {
  double value(const Point<spacedim> &p, const unsigned int = 0) const override { return 2.0 + p[0]; }
};
template <int spacedim> struct BoundaryValue : Function<spacedim> {
  double value(const Point<spacedim> &p, const unsigned int = 0) const override { return 1.0 + p[1] * p[1]; }
};

template <int dim0, int dim1, int spacedim> int run(std::ifstream &in, const char *out_path)
{
  Triangulation<dim0, spacedim> space;
  Triangulation<dim1, spacedim> immersed;
  DoFHandler<dim0, spacedim> space_dh;
  DoFHandler<dim1, spacedim> immersed_dh;
  AffineConstraints<double> constraints, immersed_constraints;
  unsigned n_bg, n_im, n_space_dofs, n_im_dofs, n_lines;
  in >> n_bg >> n_space_dofs;
  space.cells.resize(n_bg);
  space_dh.fe.dpc = 1u << dim0;
  space_dh.table.resize(std::size_t(n_bg) * space_dh.fe.dpc);
  for (unsigned c = 0; c < n_bg; ++c) {
    for (unsigned v = 0; v < (1u << dim0); ++v)
      for (int d = 0; d < spacedim; ++d) in >> space.cells[c][v][d];
    for (unsigned v = 0; v < (1u << dim0); ++v) in >> space_dh.table[std::size_t(c) * space_dh.fe.dpc + v];
  }
  in >> n_lines;
  for (unsigned l = 0; l < n_lines; ++l) {
    unsigned dof, nm;
    in >> dof >> nm;
    auto &line = constraints.lines[dof];
    for (unsigned k = 0; k < nm; ++k) {
      unsigned m; double w;
      in >> m >> w;
      line.emplace_back(m, w);
    }
  }
  in >> n_im >> n_im_dofs;
  immersed.cells.resize(n_im);
  immersed_dh.fe.dpc = 1;
  immersed_dh.table.resize(n_im);
  for (unsigned c = 0; c < n_im; ++c) {
    for (unsigned v = 0; v < (1u << dim1); ++v)
      for (int d = 0; d < spacedim; ++d) in >> immersed.cells[c][v][d];
    in >> immersed_dh.table[c];
  }
  space_dh.tria = &space; space_dh.ndofs = n_space_dofs;
  immersed_dh.tria = &immersed; immersed_dh.ndofs = n_im_dofs;
  Mapping<dim0, spacedim> space_mapping;
  Mapping<dim1, spacedim> immersed_mapping;
  GridTools::Cache<dim0, spacedim> space_cache(space, space_mapping);
  GridTools::Cache<dim1, spacedim> immersed_cache(immersed, immersed_mapping);

  // the three calls of PoissonLM::run / setup_coupling / assemble_system
  const auto cells_and_quads = NonMatching::collect_quadratures_on_overlapped_grids(space_cache, immersed_cache, 3, 1e-15);
  double area = 0;
  for (const auto &t : cells_and_quads)
    for (const double w : std::get<2>(t).get_weights()) area += w;
  DynamicSparsityPattern dsp(n_space_dofs, n_im_dofs);
  NonMatching::create_coupling_sparsity_pattern_with_exact_intersections(cells_and_quads, space_dh, immersed_dh, dsp, constraints,
                                                                        ComponentMask(), ComponentMask(), immersed_constraints);
  StubSparseMatrix<double> matrix(n_space_dofs, n_im_dofs);
  NonMatching::create_coupling_mass_matrix_with_exact_intersections(space_dh, immersed_dh, cells_and_quads, matrix, constraints,
                                                                   ComponentMask(), ComponentMask(), space_mapping, immersed_mapping,
                                                                   immersed_constraints);
  if (!matrix.compressed) return 3;
  std::FILE *f = std::fopen(out_path, "w");
  std::fprintf(f, "%zu %.17g\n", cells_and_quads.size(), area);
  std::size_t nnz = 0;
  for (const auto &r : dsp.entries) nnz += r.size();
  std::fprintf(f, "%zu\n", nnz);
  for (unsigned r = 0; r < dsp.rows; ++r)
    for (auto c : dsp.entries[r]) {
      auto it = matrix.data[r].find(c);
      std::fprintf(f, "%u %u %.17g\n", r, c, it == matrix.data[r].end() ? 0.0 : it->second);
    }
  // the two interface-penalisation calls of the Nitsche drivers, on the same tuples
  StubSparseMatrix<double> system_matrix(n_space_dofs, n_space_dofs);
  Vector<double> system_rhs(n_space_dofs);
  NonMatching::assemble_nitsche_with_exact_intersections<dim0, dim1, spacedim>(space_dh, cells_and_quads, system_matrix, constraints,
                                                                              ComponentMask(), space_mapping, Coefficient<spacedim>(),
                                                                              10.0);
  NonMatching::create_nitsche_rhs_with_exact_intersections<dim0, dim1, spacedim>(space_dh, cells_and_quads, system_rhs, constraints,
                                                                                space_mapping, BoundaryValue<spacedim>(),
                                                                                Functions::ConstantFunction<spacedim>(2.0), 10.0);
  if (system_matrix.compressed) return 6;   // the reference leaves compress() to the driver
  std::size_t n_nit = 0;
  for (const auto &r : system_matrix.data) n_nit += r.size();
  std::fprintf(f, "%zu\n", n_nit);
  for (unsigned r = 0; r < n_space_dofs; ++r)
    for (const auto &e : system_matrix.data[r]) std::fprintf(f, "%u %u %.17g\n", r, e.first, e.second);
  for (unsigned r = 0; r < n_space_dofs; ++r) std::fprintf(f, "%.17g\n", system_rhs(r));
  std::fclose(f);
  // optional second part of the grid file: higher-degree handlers on the same triangulations (FE_Q(p) x FE_DGQ(q)); the
  // same three calls, with a rule of 2p + 1 points as the drivers choose it (2d3d/lagrange_multiplier.cc:750-753)
  unsigned dpc0 = 0;
  if (in >> dpc0) {
    DoFHandler<dim0, spacedim> space_dh2;
    DoFHandler<dim1, spacedim> immersed_dh2;
    unsigned dpc1 = 0, ns2 = 0, ni2 = 0, nl2 = 0;
    in >> space_dh2.fe.degree;
    space_dh2.fe.dpc = dpc0;
    space_dh2.fe.unit_support_points.resize(dpc0);
    for (unsigned i = 0; i < dpc0; ++i)
      for (int d = 0; d < dim0; ++d) in >> space_dh2.fe.unit_support_points[i][d];
    in >> ns2;
    space_dh2.table.resize(std::size_t(n_bg) * dpc0);
    for (auto &v : space_dh2.table) in >> v;
    AffineConstraints<double> constraints2;
    in >> nl2;
    for (unsigned l = 0; l < nl2; ++l) {
      unsigned dof, nm;
      in >> dof >> nm;
      auto &line = constraints2.lines[dof];
      for (unsigned k = 0; k < nm; ++k) { unsigned m; double w; in >> m >> w; line.emplace_back(m, w); }
    }
    in >> dpc1 >> immersed_dh2.fe.degree;
    immersed_dh2.fe.dpc = dpc1;
    immersed_dh2.fe.unit_support_points.resize(dpc1);
    for (unsigned i = 0; i < dpc1; ++i)
      for (int d = 0; d < dim1; ++d) in >> immersed_dh2.fe.unit_support_points[i][d];
    in >> ni2;
    immersed_dh2.table.resize(std::size_t(n_im) * dpc1);
    for (auto &v : immersed_dh2.table) in >> v;
    space_dh2.tria = &space; space_dh2.ndofs = ns2;
    immersed_dh2.tria = &immersed; immersed_dh2.ndofs = ni2;
    const auto info2 = NonMatching::collect_quadratures_on_overlapped_grids(space_cache, immersed_cache, 2 * space_dh2.fe.degree + 1, 1e-15);
    DynamicSparsityPattern dsp2(ns2, ni2);
    NonMatching::create_coupling_sparsity_pattern_with_exact_intersections(info2, space_dh2, immersed_dh2, dsp2, constraints2, ComponentMask(),
                                                                          ComponentMask(), immersed_constraints);
    StubSparseMatrix<double> matrix2(ns2, ni2);
    NonMatching::create_coupling_mass_matrix_with_exact_intersections(space_dh2, immersed_dh2, info2, matrix2, constraints2, ComponentMask(),
                                                                     ComponentMask(), space_mapping, immersed_mapping, immersed_constraints);
    // and once more from a COPY of the tuples (a foreign vector: the quadrature travels through the host)
    const auto copy2 = info2;
    StubSparseMatrix<double> matrix3(ns2, ni2);
    NonMatching::create_coupling_mass_matrix_with_exact_intersections(space_dh2, immersed_dh2, copy2, matrix3, constraints2, ComponentMask(),
                                                                     ComponentMask(), space_mapping, immersed_mapping, immersed_constraints);
    std::FILE *g = std::fopen((std::string(out_path) + ".fe").c_str(), "w");
    std::size_t nnz2 = 0;
    for (const auto &r : dsp2.entries) nnz2 += r.size();
    std::fprintf(g, "%zu\n", nnz2);
    for (unsigned r = 0; r < dsp2.rows; ++r)
      for (auto c : dsp2.entries[r]) {
        auto it = matrix2.data[r].find(c);
        auto jt = matrix3.data[r].find(c);
        std::fprintf(g, "%u %u %.17g %.17g\n", r, c, it == matrix2.data[r].end() ? 0.0 : it->second, jt == matrix3.data[r].end() ? 0.0 : jt->second);
      }
    std::fclose(g);
  }
  // error behaviour of the reference: AssertThrow(degree > 0)
  try {
    NonMatching::collect_quadratures_on_overlapped_grids(space_cache, immersed_cache, 0, 1e-15);
    return 4;
  } catch (const std::exception &e) {
    if (std::string(e.what()).find("Invalid quadrature degree") == std::string::npos) return 5;
  }
  return 0;
}

int main(int argc, char **argv)
{
  if (argc < 3) { std::cerr << "usage: test_dropin <grid.txt> <out.txt>\n"; return 2; }
  std::ifstream in(argv[1]);
  int spacedim;
  in >> spacedim;
  try {
    if (spacedim == 22) return run<2, 2, 2>(in, argv[2]);     // codimension 0: a 2D body in the 2D background
    return spacedim == 2 ? run<2, 1, 2>(in, argv[2]) : run<3, 2, 3>(in, argv[2]);
  } catch (const std::exception &e) {
    std::cerr << "exception: " << e.what() << "\n";
    return 1;
  }
}
