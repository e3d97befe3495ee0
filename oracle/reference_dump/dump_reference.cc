// TEST INFRASTRUCTURE -- pins the oracle (and through it the GPU path) to the REAL reference.
//
// Build and run INSIDE the reference's own Docker image (environment/Dockerfile: deal.II 9.4.2 + CGAL 5.5.2), where the
// reference library is built (code/run.sh).  This program calls the reference's own entry points
//   NonMatching::collect_quadratures_on_overlapped_grids            code/src/coupling_utilities.cc:22-69
//   create_coupling_sparsity_pattern_with_exact_intersections       code/include/coupling_utilities.h:28-152
//   create_coupling_mass_matrix_with_exact_intersections            code/include/coupling_utilities.h:154-304
// on small grids with hanging nodes and writes everything a reader without deal.II needs to replay the case:
// the cells (vertices in deal.II order, DoF indices), the closed constraint lines, the accepted pairs with their sum of
// weights and quadrature, and the assembled matrix.  tests/golden/ref_to_npz.py turns the text dumps into
// tests/golden/ref_<case>.npz; tests/test_reference_dumps.py then checks oracle and GPU against them.  Until such a file
// exists, parity stays "unpinned" (DESIGN.md).
//
//   cases: disk_c<k> (<2,1,2>, k = 0..2)   sphere_c<k> (<3,2,3>, k = 0..1)      usage: dump_reference <out-dir>
#include <deal.II/base/quadrature.h>
#include <deal.II/dofs/dof_handler.h>
#include <deal.II/dofs/dof_tools.h>
#include <deal.II/fe/fe_dgq.h>
#include <deal.II/fe/fe_q.h>
#include <deal.II/fe/mapping_q1.h>
#include <deal.II/grid/grid_generator.h>
#include <deal.II/grid/grid_tools.h>
#include <deal.II/grid/grid_tools_cache.h>
#include <deal.II/grid/tria.h>
#include <deal.II/lac/affine_constraints.h>
#include <deal.II/lac/dynamic_sparsity_pattern.h>
#include <deal.II/lac/sparse_matrix.h>
#include <deal.II/lac/sparsity_pattern.h>
#include <deal.II/numerics/vector_tools.h>

#include <coupling_utilities.h>   // the REFERENCE's header (code/include), not this repository's drop-in

#include <fstream>
#include <iomanip>
#include <iostream>

using namespace dealii;

template <int dim, int spacedim>
void run_case(const std::string &name, const unsigned int cycle, const std::string &dir, const unsigned int fe_space_degree = 1,
              const unsigned int fe_immersed_degree = 0)
{
  constexpr int dim1 = spacedim - 1;
  Triangulation<spacedim>       space;
  Triangulation<dim1, spacedim> immersed;
  GridGenerator::hyper_cube(space, -1., 1.);
  space.refine_global(3 + cycle);
  Point<spacedim> center;
  for (unsigned int d = 0; d < spacedim; ++d)
    center[d] = .5;
  GridGenerator::hyper_sphere(immersed, center, .3);
  immersed.refine_global((spacedim == 2 ? 3 : 1) + cycle);

  MappingQ1<spacedim>       space_mapping;
  MappingQ1<dim1, spacedim> immersed_mapping;
  // one round of local refinement around the interface: hanging nodes on the path
  {
    GridTools::Cache<spacedim, spacedim> c0(space, space_mapping);
    GridTools::Cache<dim1, spacedim>     c1(immersed, immersed_mapping);
    const auto info = NonMatching::collect_quadratures_on_overlapped_grids(c0, c1, 3, 1e-15);
    for (const auto &t : info)
      std::get<0>(t)->set_refine_flag();
    space.execute_coarsening_and_refinement();
  }
  GridTools::Cache<spacedim, spacedim> space_cache(space, space_mapping);
  GridTools::Cache<dim1, spacedim>     immersed_cache(immersed, immersed_mapping);

  FE_Q<spacedim>         space_fe(fe_space_degree);
  FE_DGQ<dim1, spacedim> immersed_fe(fe_immersed_degree);
  DoFHandler<spacedim>       space_dh(space);
  DoFHandler<dim1, spacedim> immersed_dh(immersed);
  space_dh.distribute_dofs(space_fe);
  immersed_dh.distribute_dofs(immersed_fe);
  AffineConstraints<double> constraints, immersed_constraints;
  DoFTools::make_hanging_node_constraints(space_dh, constraints);
  VectorTools::interpolate_boundary_values(space_dh, 0, Functions::ZeroFunction<spacedim>(), constraints);
  constraints.close();
  immersed_constraints.close();

  const unsigned int degree = 2 * fe_space_degree + 1;
  const auto cells_and_quads = NonMatching::collect_quadratures_on_overlapped_grids(space_cache, immersed_cache, degree, 1e-15);

  DynamicSparsityPattern dsp(space_dh.n_dofs(), immersed_dh.n_dofs());
  NonMatching::create_coupling_sparsity_pattern_with_exact_intersections(cells_and_quads, space_dh, immersed_dh, dsp, constraints, ComponentMask(),
                                                                        ComponentMask(), immersed_constraints);
  SparsityPattern sp;
  sp.copy_from(dsp);
  SparseMatrix<double> matrix(sp);
  NonMatching::create_coupling_mass_matrix_with_exact_intersections(space_dh, immersed_dh, cells_and_quads, matrix, constraints, ComponentMask(),
                                                                   ComponentMask(), space_mapping, immersed_mapping, immersed_constraints);

  std::ofstream out(dir + "/ref_" + name + "_c" + std::to_string(cycle) + "_p" + std::to_string(fe_space_degree) + "q" +
                    std::to_string(fe_immersed_degree) + ".txt");
  out << std::setprecision(17);
  out << "spacedim " << spacedim << " degree " << degree << " tol 1e-15 fe " << fe_space_degree << " " << fe_immersed_degree << "\n";
  auto dump_fe = [&](const auto &fe, const char *tag) {
    out << tag << " " << fe.n_dofs_per_cell() << "\n";
    for (const auto &p : fe.get_unit_support_points()) {
      for (unsigned int d = 0; d < p.dimension; ++d)
        out << p[d] << " ";
      out << "\n";
    }
  };
  dump_fe(space_fe, "space_fe");
  dump_fe(immersed_fe, "immersed_fe");
  auto dump_cells = [&](const auto &dh, const auto &mapping, const char *tag) {
    out << tag << " " << dh.get_triangulation().n_active_cells() << " " << dh.n_dofs() << "\n";
    std::vector<types::global_dof_index> dofs(dh.get_fe().n_dofs_per_cell());
    for (const auto &cell : dh.active_cell_iterators()) {
      for (const auto &v : mapping.get_vertices(cell))
        for (unsigned int d = 0; d < spacedim; ++d)
          out << v[d] << " ";
      cell->get_dof_indices(dofs);
      for (const auto i : dofs)
        out << i << " ";
      out << "\n";
    }
  };
  dump_cells(space_dh, space_mapping, "space_cells");
  dump_cells(immersed_dh, immersed_mapping, "immersed_cells");
  out << "constraints " << constraints.n_constraints() << "\n";
  for (const auto &line : constraints.get_lines()) {
    out << line.index << " " << line.entries.size();
    for (const auto &e : line.entries)
      out << " " << e.first << " " << e.second;
    out << "\n";
  }
  out << "pairs " << cells_and_quads.size() << "\n";
  for (const auto &t : cells_and_quads) {
    const auto &q = std::get<2>(t);
    double sum_w = 0;
    for (const double w : q.get_weights())
      sum_w += w;
    out << std::get<0>(t)->active_cell_index() << " " << std::get<1>(t)->active_cell_index() << " " << sum_w << " " << q.size();
    for (unsigned int k = 0; k < q.size(); ++k) {
      for (unsigned int d = 0; d < spacedim; ++d)
        out << " " << q.point(k)[d];
      out << " " << q.weight(k);
    }
    out << "\n";
  }
  out << "matrix " << sp.n_nonzero_elements() << "\n";
  for (types::global_dof_index r = 0; r < space_dh.n_dofs(); ++r)
    for (auto it = matrix.begin(r); it != matrix.end(r); ++it)
      out << r << " " << it->column() << " " << it->value() << "\n";
  std::cout << name << " cycle " << cycle << " (p, q) = (" << fe_space_degree << ", " << fe_immersed_degree << "): " << cells_and_quads.size()
            << " pairs, " << sp.n_nonzero_elements() << " stored entries\n";
}

int main(int argc, char **argv)
{
  const std::string dir = argc > 1 ? argv[1] : ".";
  for (unsigned int c = 0; c < 3; ++c)
    run_case<1, 2>("disk", c, dir);
  for (unsigned int c = 0; c < 2; ++c)
    run_case<2, 3>("sphere", c, dir);
  // the general path: the driver defaults are degree 1 for both spaces
  run_case<1, 2>("disk", 0, dir, 1, 1);
  run_case<1, 2>("disk", 0, dir, 2, 1);
  run_case<2, 3>("sphere", 0, dir, 1, 1);
  return 0;
}
