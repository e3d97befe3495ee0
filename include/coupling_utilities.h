// Drop-in replacement for the reference's code/include/coupling_utilities.h, hot path only.
//
// Same namespace, names and signatures as the reference, implemented as thin marshalling over
// the C ABI of libnmgpu.so (include/nmgpu.h):
//
//   collect_quadratures_on_overlapped_grids<dim0,dim1,spacedim>          reference :508-515, src/coupling_utilities.cc:22-69
//   create_coupling_sparsity_pattern_with_exact_intersections<...>       reference :28-152
//   create_coupling_mass_matrix_with_exact_intersections<...>            reference :154-304
//
// Implemented instantiations: <2,1,2> and <3,2,3>, and the codimension-0 <2,2,2> (convex immersed quads, FE_Q(1) x FE_DGQ(0)),
// on Cartesian background cells.  FE_Q(1) x FE_DGQ(0) (every LM
// configuration in data/*.prm) runs on the fused kernels; any other FE_Q(p) x FE_DGQ(q) pair (p <= 5, <= 64 / 16 DoFs per
// cell) on the general path (nm_set_finite_elements / nm_assemble_fe), described to the library by the elements' unit
// support points.  Anything else raises ExcNotImplemented -- there is no CPU fallback.
//
// First widening row, same boundary:
//   assemble_nitsche_with_exact_intersections<...>                       reference :306-398
//   create_nitsche_rhs_with_exact_intersections<...>                     reference :400-481
// for FE_Q(1) on Cartesian background cells; the Function objects are evaluated here on the host
// (value_list, as the reference does at :349,:438-441), everything else runs on the GPU.
//
// Build: add -I<repo>/include and link -L<repo>/non_matching_test_suite_b200 -lnmgpu.
// The GPU is chosen with the environment variable NMGPU_DEVICE (default 0; with MPI set it to the
// local rank).  See INTEGRATION.md.
#ifndef nmgpu_coupling_utilities_h
#define nmgpu_coupling_utilities_h

#include <deal.II/base/function.h>
#include <deal.II/base/point.h>
#include <deal.II/base/quadrature.h>
#include <deal.II/dofs/dof_handler.h>
#include <deal.II/fe/component_mask.h>
#include <deal.II/fe/mapping.h>
#include <deal.II/grid/grid_tools_cache.h>
#include <deal.II/grid/tria.h>
#include <deal.II/lac/affine_constraints.h>

#include <cstdint>
#include <algorithm>
#include <cstdlib>
#include <map>
#include <memory>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

#include "nmgpu.h"

namespace dealii {
namespace NonMatching {

namespace nmgpu_detail {

// One GPU context per (space triangulation, immersed triangulation) pair.  The three reference
// calls of one refinement cycle find each other through the triangulation addresses.
struct Session {
  nm_ctx *ctx = nullptr;
  bool grids_set = false, space_set = false;
  bool general_fe = false;   // the dof tables on the device belong to FE_Q(p) x FE_DGQ(q) beyond (1, 0): nm_assemble_fe builds the matrix
  std::uint64_t n_space = 0, n_immersed = 0;
  std::size_t space_active_cells = 0, immersed_active_cells = 0;   // n_active_cells() of the grids on the device
  std::vector<unsigned int> space_local_of_active;     // active_cell_index -> row in the arrays sent to the GPU (or -1)
  std::vector<unsigned int> immersed_local_of_active;
  // Fingerprint of the vector collect_quadratures_on_overlapped_grids returned last.  When the sparsity / mass-matrix
  // calls receive that same vector (the reference's flow: 2d3d/lagrange_multiplier.cc:750-753 -> 488-490 -> 599-602)
  // the quadrature is still on the device and nothing is sent back; any other vector takes the upload path.
  struct {
    bool valid = false;
    std::size_t size = 0;
    const void *first = nullptr, *last = nullptr;      // point storage of the first / last Quadrature
    double w_first = 0, w_last = 0;
  } own;
  // the DoF tables on the device were taken from these objects
  struct {
    bool valid = false;
    const void *space_dh = nullptr, *immersed_dh = nullptr, *constraints = nullptr;
    std::uint64_t n_space = 0, n_immersed = 0, n_constraints = 0;
  } dofs;
  ~Session() { if (ctx) nm_destroy(ctx); }
};

using SessionMap = std::map<std::pair<const void *, const void *>, std::unique_ptr<Session>>;
inline SessionMap &sessions()
{
  // never destroyed: no CUDA call may run during static destruction, when the runtime may already be gone;
  // nmgpu_release() frees the contexts (and their device memory) explicitly
  static SessionMap *m = new SessionMap();
  return *m;
}

inline Session &session(const void *space_tria, const void *immersed_tria)
{
  auto &slot = sessions()[{space_tria, immersed_tria}];
  if (!slot) {
    slot = std::make_unique<Session>();
    const char *dev = std::getenv("NMGPU_DEVICE");
    const int rc = nm_create(&slot->ctx, dev ? std::atoi(dev) : 0);
    AssertThrow(rc == NM_OK, ExcMessage("nmgpu: no CUDA device / context creation failed (there is no CPU fallback)"));
  }
  return *slot;
}

// The grids on the device are those of an earlier call: still the same mesh?  (A refine/coarsen between
// collect_quadratures_on_overlapped_grids and the matrix calls changes the active cell count.)
template <int dim0, int dim1, int spacedim>
bool grids_current(const Session &s, const Triangulation<dim0, spacedim> &space, const Triangulation<dim1, spacedim> &immersed)
{
  return s.grids_set && s.space_active_cells == space.n_active_cells() && s.immersed_active_cells == immersed.n_active_cells();
}

template <typename Info> bool is_own_result(const Session &s, const Info &info)
{
  if (!s.own.valid || info.size() != s.own.size)
    return false;
  if (info.empty())
    return true;
  const auto &qf = std::get<2>(info.front());
  const auto &ql = std::get<2>(info.back());
  return qf.size() > 0 && ql.size() > 0 && static_cast<const void *>(qf.get_points().data()) == s.own.first &&
         static_cast<const void *>(ql.get_points().data()) == s.own.last && qf.weight(0) == s.own.w_first &&
         ql.weight(ql.size() - 1) == s.own.w_last;
}

// Host threads for the embarrassingly parallel part of the marshalling (building the millions of Quadrature objects the
// reference's return type demands).  NMGPU_HOST_THREADS overrides; the drivers' own threading is not touched.
inline unsigned int host_threads()
{
  if (const char *e = std::getenv("NMGPU_HOST_THREADS"))
    return std::max(1, std::atoi(e));
  const unsigned int hw = std::thread::hardware_concurrency();
  return hw == 0 ? 1u : std::min(hw, 8u);
}

template <typename F> void parallel_for(const std::size_t n, F &&body)
{
  const unsigned int nt = static_cast<unsigned int>(std::min<std::size_t>(host_threads(), n / 4096 + 1));
  if (nt <= 1) {
    for (std::size_t i = 0; i < n; ++i)
      body(i);
    return;
  }
  std::vector<std::thread> pool;
  for (unsigned int t = 0; t < nt; ++t)
    pool.emplace_back([&, t]() {
      const std::size_t b = n * t / nt, e = n * (t + 1) / nt;
      for (std::size_t i = b; i < e; ++i)
        body(i);
    });
  for (auto &th : pool)
    th.join();
}

// Receive buffer for a large read-back: not value-initialised (a std::vector would clear 2.5 GB of points on one thread
// before the copy overwrites them), and its pages are faulted in by all host threads at once -- a first touch by the
// copy itself costs more than the transfer.
template <typename T> struct HostBuffer {
  std::unique_ptr<T[]> p;
  std::size_t n = 0;
  explicit HostBuffer(std::size_t count) : p(count ? new T[count] : nullptr), n(count)
  {
    constexpr std::size_t page = 4096 / sizeof(T) ? 4096 / sizeof(T) : 1;
    T *base = p.get();
    parallel_for((n + page - 1) / page, [base, page](const std::size_t i) { reinterpret_cast<volatile char *>(base + i * page)[0] = 0; });
  }
  T *data() { return p.get(); }
  const T &operator[](std::size_t i) const { return p[i]; }
};

inline void check(const Session &s, int rc)
{
  AssertThrow(rc == NM_OK, ExcMessage(std::string("nmgpu: ") + nm_last_error(s.ctx)));
}

// mapping.get_vertices(cell) of every (locally owned) active cell, cell-major, deal.II order
// (what the reference feeds to CGAL through get_vertices_in_cgal_order, intersections.h:39-58).
template <int dim, int spacedim>
void gather_vertices(const Triangulation<dim, spacedim> &tria, const Mapping<dim, spacedim> &mapping, const bool owned_only,
                     std::vector<double> &verts, std::vector<unsigned int> &local_of_active,
                     std::vector<typename Triangulation<dim, spacedim>::cell_iterator> *cells)
{
  constexpr unsigned int nv = 1u << dim;
  local_of_active.assign(tria.n_active_cells(), static_cast<unsigned int>(-1));
  verts.clear();
  unsigned int n = 0;
  for (const auto &cell : tria.active_cell_iterators()) {
    if (owned_only && !cell->is_locally_owned())
      continue;
    const auto v = mapping.get_vertices(cell);
    AssertThrow(v.size() == nv, ExcNotImplemented());
    for (unsigned int i = 0; i < nv; ++i)
      for (unsigned int d = 0; d < spacedim; ++d)
        verts.push_back(v[i][d]);
    local_of_active[cell->active_cell_index()] = n++;
    if (cells)
      cells->push_back(cell);
  }
}

// Background cells as boxes: every configuration of the reference has Cartesian background cells
// (GridGenerator::hyper_cube + refinements), so lo/hi corners carry the geometry in a quarter of the
// bytes.  A cell whose vertices are not those of an axis-aligned box in deal.II order is refused
// here exactly as the library itself would refuse it.
template <int dim, int spacedim>
void gather_boxes(const Triangulation<dim, spacedim> &tria, const Mapping<dim, spacedim> &mapping, std::vector<double> &lo,
                  std::vector<double> &hi, std::vector<unsigned int> &local_of_active,
                  std::vector<typename Triangulation<dim, spacedim>::cell_iterator> *cells)
{
  constexpr unsigned int nv = 1u << dim;
  local_of_active.assign(tria.n_active_cells(), static_cast<unsigned int>(-1));
  lo.clear();
  hi.clear();
  unsigned int n = 0;
  for (const auto &cell : tria.active_cell_iterators()) {
    if (!cell->is_locally_owned())
      continue;
    const auto v = mapping.get_vertices(cell);
    AssertThrow(v.size() == nv, ExcNotImplemented());
    bool ok = true;
    for (unsigned int i = 0; i < nv; ++i)
      for (unsigned int d = 0; d < spacedim; ++d)
        ok = ok && v[i][d] == (((i >> d) & 1u) ? v[nv - 1][d] : v[0][d]);
    AssertThrow(ok, ExcMessage("nmgpu: only Cartesian background cells are implemented"));
    for (unsigned int d = 0; d < spacedim; ++d) {
      lo.push_back(v[0][d]);
      hi.push_back(v[nv - 1][d]);
    }
    local_of_active[cell->active_cell_index()] = n++;
    if (cells)
      cells->push_back(cell);
  }
}

template <int dim0, int dim1, int spacedim>
void set_grids(Session &s, const Triangulation<dim0, spacedim> &space, const Mapping<dim0, spacedim> &mapping0,
               const Triangulation<dim1, spacedim> &immersed, const Mapping<dim1, spacedim> &mapping1,
               std::vector<typename Triangulation<dim0, spacedim>::cell_iterator> *space_cells,
               std::vector<typename Triangulation<dim1, spacedim>::cell_iterator> *immersed_cells)
{
  std::vector<double> lo, hi, v1;
  // locally owned background cells, the whole immersed grid (coupling_utilities.cc:40-45)
  gather_boxes(space, mapping0, lo, hi, s.space_local_of_active, space_cells);
  gather_vertices(immersed, mapping1, false, v1, s.immersed_local_of_active, immersed_cells);
  s.n_space = lo.size() / spacedim;
  s.n_immersed = v1.size() / ((1u << dim1) * spacedim);
  check(s, nm_set_background_boxes(s.ctx, spacedim, s.n_space, lo.data(), hi.data(), nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr,
                                   NM_HOST));
  check(s, nm_set_immersed(s.ctx, dim1, spacedim, s.n_immersed, v1.data(), nullptr, 1, 0, NM_HOST));
  s.grids_set = true;
  s.space_set = true;
  s.space_active_cells = space.n_active_cells();
  s.immersed_active_cells = immersed.n_active_cells();
  s.own.valid = false;
  s.dofs.valid = false;
}

// DoF indices of the background cells + the lines of the closed AffineConstraints object as CSR.
template <int dim0, int spacedim, typename number>
void set_space_dofs(Session &s, const DoFHandler<dim0, spacedim> &space_dh, const AffineConstraints<number> &constraints)
{
  const auto &space_fe = space_dh.get_fe();
  AssertThrow(space_fe.n_components() == 1, ExcNotImplemented());
  AssertThrow(space_fe.n_dofs_per_cell() == (1u << dim0), ExcMessage("nmgpu: the background space must be FE_Q(1)"));
  std::vector<std::uint32_t> d0(s.n_space * (1u << dim0));
  std::vector<types::global_dof_index> tmp0(1u << dim0);
  for (const auto &cell : space_dh.active_cell_iterators()) {
    const unsigned int l = s.space_local_of_active[cell->active_cell_index()];
    if (l == static_cast<unsigned int>(-1))
      continue;
    cell->get_dof_indices(tmp0);
    for (unsigned int i = 0; i < tmp0.size(); ++i)
      d0[std::size_t(l) * tmp0.size() + i] = static_cast<std::uint32_t>(tmp0[i]);
  }
  std::vector<std::uint32_t> c_dof, c_master;
  std::vector<std::uint64_t> c_ptr(1, 0);
  std::vector<double> c_weight;
  // the lines of the (closed, hence index-sorted) constraint object; never is_constrained(i) over all dofs, which
  // asserts on a distributed AffineConstraints with local_lines
  for (const auto &line : constraints.get_lines()) {
    c_dof.push_back(static_cast<std::uint32_t>(line.index));
    for (const auto &e : line.entries) {
      c_master.push_back(static_cast<std::uint32_t>(e.first));
      c_weight.push_back(static_cast<double>(e.second));
    }
    c_ptr.push_back(c_master.size());
  }
  check(s, nm_set_background_dofs(s.ctx, d0.data(), space_dh.n_dofs(), c_dof.size(), c_dof.data(), c_ptr.data(), c_master.data(),
                                  c_weight.data(), NM_HOST));
  s.dofs.valid = false;   // set_dofs marks the pair of tables valid again
}

template <int dim0, int dim1, int spacedim, typename number>
void set_dofs(Session &s, const DoFHandler<dim0, spacedim> &space_dh, const DoFHandler<dim1, spacedim> &immersed_dh,
              const AffineConstraints<number> &constraints, const AffineConstraints<number> &immersed_constraints)
{
  const auto &space_fe = space_dh.get_fe();
  const auto &immersed_fe = immersed_dh.get_fe();
  AssertThrow(space_fe.n_components() == 1 && immersed_fe.n_components() == 1, ExcNotImplemented());
  AssertThrow(immersed_constraints.n_constraints() == 0, ExcNotImplemented());
  // the sparsity and the mass-matrix call of one cycle pass the same handlers and constraints: send the tables once
  if (s.dofs.valid && s.dofs.space_dh == &space_dh && s.dofs.immersed_dh == &immersed_dh && s.dofs.constraints == &constraints &&
      s.dofs.n_space == space_dh.n_dofs() && s.dofs.n_immersed == immersed_dh.n_dofs() && s.dofs.n_constraints == constraints.n_constraints())
    return;
  s.dofs.valid = false;
  s.general_fe = !(space_fe.n_dofs_per_cell() == (1u << dim0) && immersed_fe.n_dofs_per_cell() == 1);
  if (!s.general_fe) {   // FE_Q(1) x FE_DGQ(0): the configuration of every shipped .prm, on the fused kernels
    set_space_dofs<dim0, spacedim>(s, space_dh, constraints);
    std::vector<std::uint32_t> d1(s.n_immersed);
    std::vector<types::global_dof_index> tmp1(1);
    for (const auto &cell : immersed_dh.active_cell_iterators()) {
      cell->get_dof_indices(tmp1);
      d1[s.immersed_local_of_active[cell->active_cell_index()]] = static_cast<std::uint32_t>(tmp1[0]);
    }
    check(s, nm_set_immersed_dofs(s.ctx, d1.data(), 1, immersed_dh.n_dofs(), NM_HOST));
  } else {
    // FE_Q(p) x FE_DGQ(q): the elements are described by their unit support points in their own dof order
    auto table = [&](const auto &dh, const std::vector<unsigned int> &local_of_active, const std::uint64_t n_cells) {
      const unsigned int dpc = dh.get_fe().n_dofs_per_cell();
      std::vector<std::uint32_t> tab(n_cells * dpc);
      std::vector<types::global_dof_index> tmp(dpc);
      for (const auto &cell : dh.active_cell_iterators()) {
        const unsigned int l = local_of_active[cell->active_cell_index()];
        if (l == static_cast<unsigned int>(-1))
          continue;
        cell->get_dof_indices(tmp);
        for (unsigned int i = 0; i < dpc; ++i)
          tab[std::size_t(l) * dpc + i] = static_cast<std::uint32_t>(tmp[i]);
      }
      return tab;
    };
    auto support = [](const auto &fe, const unsigned int dim) {
      const auto &pts = fe.get_unit_support_points();
      AssertThrow(pts.size() == fe.n_dofs_per_cell(), ExcMessage("nmgpu: the element has no unit support points (not a Lagrange element)"));
      std::vector<double> flat;
      for (const auto &p : pts)
        for (unsigned int k = 0; k < dim; ++k)
          flat.push_back(p[k]);
      return flat;
    };
    const auto d0 = table(space_dh, s.space_local_of_active, s.n_space);
    const auto d1 = table(immersed_dh, s.immersed_local_of_active, s.n_immersed);
    const auto sp0 = support(space_fe, dim0), sp1 = support(immersed_fe, dim1);
    std::vector<std::uint32_t> c_dof, c_master;
    std::vector<std::uint64_t> c_ptr(1, 0);
    std::vector<double> c_weight;
    for (const auto &line : constraints.get_lines()) {
      c_dof.push_back(static_cast<std::uint32_t>(line.index));
      for (const auto &e : line.entries) {
        c_master.push_back(static_cast<std::uint32_t>(e.first));
        c_weight.push_back(static_cast<double>(e.second));
      }
      c_ptr.push_back(c_master.size());
    }
    const nm_fe f0{space_fe.degree, space_fe.n_dofs_per_cell(), sp0.data()}, f1{immersed_fe.degree, immersed_fe.n_dofs_per_cell(), sp1.data()};
    check(s, nm_set_finite_elements(s.ctx, &f0, d0.data(), space_dh.n_dofs(), c_dof.size(), c_dof.data(), c_ptr.data(), c_master.data(),
                                    c_weight.data(), &f1, d1.data(), immersed_dh.n_dofs(), NM_HOST));
  }
  s.dofs.space_dh = &space_dh;
  s.dofs.immersed_dh = &immersed_dh;
  s.dofs.constraints = &constraints;
  s.dofs.n_space = space_dh.n_dofs();
  s.dofs.n_immersed = immersed_dh.n_dofs();
  s.dofs.n_constraints = constraints.n_constraints();
  s.dofs.valid = true;
}

// Background geometry only (the Nitsche terms do not look at the immersed grid).
template <int dim0, int spacedim>
void set_space_grid(Session &s, const Triangulation<dim0, spacedim> &space, const Mapping<dim0, spacedim> &mapping0)
{
  std::vector<double> lo, hi;
  gather_boxes(space, mapping0, lo, hi, s.space_local_of_active, nullptr);
  s.n_space = lo.size() / spacedim;
  check(s, nm_set_background_boxes(s.ctx, spacedim, s.n_space, lo.data(), hi.data(), nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr,
                                   NM_HOST));
  s.space_set = true;
  s.grids_set = false;   // the immersed grid of this context (if any) no longer matches
  s.space_active_cells = space.n_active_cells();
  s.own.valid = false;
  s.dofs.valid = false;
}

// Flatten the tuples for nm_assemble_nitsche and evaluate the caller's Function objects at the points.
template <int dim0, int dim1, int spacedim, typename number>
void nitsche_from_info(Session &s,
                       const std::vector<std::tuple<typename Triangulation<dim0, spacedim>::cell_iterator,
                                                    typename Triangulation<dim1, spacedim>::cell_iterator, Quadrature<spacedim>>> &info,
                       const bool owned_only, const Function<spacedim, number> &coefficient, const Function<spacedim, number> *rhs_function,
                       const double penalty, const int what)
{
  std::vector<std::uint32_t> bg;
  std::vector<std::uint64_t> off(1, 0);
  std::vector<double> pts, w, cv, fv;
  std::vector<number> tmp;
  for (const auto &t : info) {
    const auto &cell = std::get<0>(t);
    if (!cell->is_active() || (owned_only && !cell->is_locally_owned()))   // reference :345 / :428
      continue;
    const unsigned int l = s.space_local_of_active[cell->active_cell_index()];
    if (l == static_cast<unsigned int>(-1))
      continue;
    bg.push_back(l);
    const auto &q = std::get<2>(t);
    tmp.resize(q.size());
    coefficient.value_list(q.get_points(), tmp);
    cv.insert(cv.end(), tmp.begin(), tmp.end());
    if (rhs_function) {
      rhs_function->value_list(q.get_points(), tmp);
      fv.insert(fv.end(), tmp.begin(), tmp.end());
    }
    for (unsigned int k = 0; k < q.size(); ++k) {
      for (unsigned int d = 0; d < spacedim; ++d)
        pts.push_back(q.point(k)[d]);
      w.push_back(q.weight(k));
    }
    off.push_back(w.size());
  }
  std::uint64_t nnz = 0;
  check(s, nm_assemble_nitsche(s.ctx, bg.size(), bg.data(), off.data(), pts.data(), w.data(), cv.data(), rhs_function ? fv.data() : nullptr,
                               penalty, what, NM_HOST, &nnz));
}

// Hand the caller's (cell, cell, Quadrature) tuples to the device: inverse mapping, shape
// values and scatter all happen there (reference :238-289).
template <int dim0, int dim1, int spacedim>
void assemble_from_info(Session &s,
                        const std::vector<std::tuple<typename Triangulation<dim0, spacedim>::cell_iterator,
                                                     typename Triangulation<dim1, spacedim>::cell_iterator, Quadrature<spacedim>>> &info)
{
  std::vector<std::uint32_t> im, bg;
  std::vector<std::uint64_t> off(1, 0);
  std::vector<double> pts, w;
  for (const auto &t : info) {
    bg.push_back(s.space_local_of_active[std::get<0>(t)->active_cell_index()]);
    im.push_back(s.immersed_local_of_active[std::get<1>(t)->active_cell_index()]);
    const auto &q = std::get<2>(t);
    for (unsigned int k = 0; k < q.size(); ++k) {
      for (unsigned int d = 0; d < spacedim; ++d)
        pts.push_back(q.point(k)[d]);
      w.push_back(q.weight(k));
    }
    off.push_back(w.size());
  }
  if (s.general_fe)
    check(s, nm_assemble_fe(s.ctx, info.size(), im.data(), bg.data(), off.data(), pts.data(), w.data(), NM_HOST, nullptr));
  else
    check(s, nm_assemble_from_quadrature(s.ctx, info.size(), im.data(), bg.data(), off.data(), pts.data(), w.data(), NM_HOST));
}

// the matrix from the quadrature that is still on the device (the caller handed back the vector collect_...() returned)
inline void assemble_from_device_result(Session &s)
{
  if (s.general_fe)
    check(s, nm_assemble_fe(s.ctx, 0, nullptr, nullptr, nullptr, nullptr, nullptr, NM_DEVICE, nullptr));
  else
    check(s, nm_assemble(s.ctx));
}

// The matrix of the context, rows with entries only, handed to `visit(row, n, cols, vals)` (vals == nullptr: pattern).
template <typename Visit> void for_each_stored_row(Session &s, const bool with_values, Visit &&visit)
{
  std::uint64_t nnz = 0, n_rows = 0;
  check(s, nm_build_sparsity(s.ctx, &nnz));
  check(s, nm_get_csr_compact(s.ctx, &n_rows, nullptr, nullptr, nullptr, nullptr, NM_HOST));
  HostBuffer<std::uint32_t> ids(n_rows), len(n_rows), col(nnz);
  HostBuffer<double> val(with_values ? nnz : 0);
  check(s, nm_get_csr_compact(s.ctx, &n_rows, ids.data(), len.data(), col.data(), with_values ? val.data() : nullptr, NM_HOST));
  std::vector<types::global_dof_index> cols;
  std::size_t o = 0;
  for (std::uint64_t r = 0; r < n_rows; ++r) {
    cols.assign(col.data() + o, col.data() + o + len[r]);
    visit(static_cast<types::global_dof_index>(ids[r]), static_cast<types::global_dof_index>(len[r]), cols, with_values ? val.data() + o : nullptr);
    o += len[r];
  }
}

} // namespace nmgpu_detail

template <int dim0, int dim1, int spacedim>
std::vector<std::tuple<typename Triangulation<dim0, spacedim>::cell_iterator, typename Triangulation<dim1, spacedim>::cell_iterator,
                       Quadrature<spacedim>>>
collect_quadratures_on_overlapped_grids(const GridTools::Cache<dim0, spacedim> &space_cache,
                                        const GridTools::Cache<dim1, spacedim> &immersed_cache, const unsigned int degree,
                                        const double tol = 1e-14)
{
  AssertThrow(dim1 <= dim0, ExcMessage("Intrinsic dimension of the immersed object must be smaller than dim0."));
  AssertThrow(degree > 0, ExcMessage("Invalid quadrature degree."));
  // codimension 1 in 2D and 3D, and the codimension-0 instantiation <2,2,2> (code/src/coupling_utilities.cc:72-102)
  AssertThrow(dim0 == spacedim && ((dim1 + 1 == dim0 && (spacedim == 2 || spacedim == 3)) || (dim1 == 2 && dim0 == 2)), ExcNotImplemented());

  using SpaceIt = typename Triangulation<dim0, spacedim>::cell_iterator;
  using ImmIt = typename Triangulation<dim1, spacedim>::cell_iterator;
  auto &s = nmgpu_detail::session(&space_cache.get_triangulation(), &immersed_cache.get_triangulation());
  std::vector<SpaceIt> space_cells;
  std::vector<ImmIt> immersed_cells;
  nmgpu_detail::set_grids<dim0, dim1, spacedim>(s, space_cache.get_triangulation(), space_cache.get_mapping(),
                                                immersed_cache.get_triangulation(), immersed_cache.get_mapping(), &space_cells,
                                                &immersed_cells);
  nm_counts counts;
  nmgpu_detail::check(s, nm_intersect(s.ctx, degree, tol, &counts));

  nmgpu_detail::HostBuffer<std::uint32_t> im(counts.n_accepted), bg(counts.n_accepted);
  nmgpu_detail::HostBuffer<std::uint64_t> off(counts.n_accepted + 1);
  nmgpu_detail::HostBuffer<double> pts(counts.n_qpoints * spacedim), w(counts.n_qpoints);
  nmgpu_detail::check(s, nm_get_pairs(s.ctx, im.data(), bg.data(), nullptr, NM_HOST));
  nmgpu_detail::check(s, nm_get_quadrature(s.ctx, off.data(), pts.data(), w.data(), NM_HOST));

  // one tuple per accepted pair, built in parallel: every tuple owns two heap vectors (what Quadrature<spacedim> is)
  std::vector<std::tuple<SpaceIt, ImmIt, Quadrature<spacedim>>> cells_with_quadratures(counts.n_accepted);
  nmgpu_detail::parallel_for(counts.n_accepted, [&](const std::size_t p) {
    const std::size_t nq = off[p + 1] - off[p];
    std::vector<Point<spacedim>> qp(nq);
    std::vector<double> qw(w.data() + off[p], w.data() + off[p + 1]);
    for (std::size_t k = 0; k < nq; ++k)
      for (unsigned int d = 0; d < spacedim; ++d)
        qp[k][d] = pts[(off[p] + k) * spacedim + d];
    cells_with_quadratures[p] = std::make_tuple(space_cells[bg[p]], immersed_cells[im[p]], Quadrature<spacedim>(std::move(qp), std::move(qw)));
  });
  // remember what was handed out (the element storage does not move when the vector is returned)
  s.own.valid = true;
  s.own.size = cells_with_quadratures.size();
  if (!cells_with_quadratures.empty()) {
    const auto &qf = std::get<2>(cells_with_quadratures.front());
    const auto &ql = std::get<2>(cells_with_quadratures.back());
    s.own.first = qf.get_points().data();
    s.own.last = ql.get_points().data();
    s.own.w_first = qf.weight(0);
    s.own.w_last = ql.weight(ql.size() - 1);
  }
  return cells_with_quadratures;
}

template <int dim0, int dim1, int spacedim, typename Sparsity, typename number>
void create_coupling_sparsity_pattern_with_exact_intersections(
    const std::vector<std::tuple<typename dealii::Triangulation<dim0, spacedim>::cell_iterator,
                                 typename dealii::Triangulation<dim1, spacedim>::cell_iterator, dealii::Quadrature<spacedim>>>
        &intersections_info,
    const DoFHandler<dim0, spacedim> &space_dh, const DoFHandler<dim1, spacedim> &immersed_dh, Sparsity &sparsity,
    const AffineConstraints<number> &constraints, const ComponentMask &space_comps, const ComponentMask &immersed_comps,
    const AffineConstraints<number> &immersed_constraints)
{
  AssertDimension(sparsity.n_rows(), space_dh.n_dofs());
  AssertDimension(sparsity.n_cols(), immersed_dh.n_dofs());
  Assert((dim1 <= dim0) && (dim0 <= spacedim), ExcMessage("This function can only work if dim1<=dim0<=spacedim"));
  AssertThrow(space_comps.size() <= 1 && immersed_comps.size() <= 1, ExcNotImplemented());

  auto &s = nmgpu_detail::session(&space_dh.get_triangulation(), &immersed_dh.get_triangulation());
  AssertThrow((nmgpu_detail::grids_current<dim0, dim1, spacedim>(s, space_dh.get_triangulation(), immersed_dh.get_triangulation())),
              ExcMessage("nmgpu: collect_quadratures_on_overlapped_grids must be called on these grids first (and again after a mesh change)"));
  nmgpu_detail::set_dofs<dim0, dim1, spacedim>(s, space_dh, immersed_dh, constraints, immersed_constraints);
  if (!nmgpu_detail::is_own_result(s, intersections_info)) {   // a foreign vector: its quadrature has to travel to the device
    nmgpu_detail::assemble_from_info<dim0, dim1, spacedim>(s, intersections_info);
    s.own.valid = false;
  } else if (s.general_fe)
    nmgpu_detail::assemble_from_device_result(s);               // pattern and values come out of one pass
  nmgpu_detail::for_each_stored_row(s, false, [&](types::global_dof_index r, types::global_dof_index, const std::vector<types::global_dof_index> &cols,
                                                  const double *) { sparsity.add_entries(r, cols.begin(), cols.end(), /*indices_are_sorted=*/true); });
}

template <int dim0, int dim1, int spacedim, typename Matrix>
void create_coupling_mass_matrix_with_exact_intersections(
    const DoFHandler<dim0, spacedim> &space_dh, const DoFHandler<dim1, spacedim> &immersed_dh,
    const std::vector<std::tuple<typename Triangulation<dim0, spacedim>::cell_iterator,
                                 typename Triangulation<dim1, spacedim>::cell_iterator, Quadrature<spacedim>>> &cells_and_quads,
    Matrix &matrix, const AffineConstraints<typename Matrix::value_type> &space_constraints, const ComponentMask &space_comps,
    const ComponentMask &immersed_comps, const Mapping<dim0, spacedim> &space_mapping, const Mapping<dim1, spacedim> &immersed_mapping,
    const AffineConstraints<typename Matrix::value_type> &immersed_constraints)
{
  AssertDimension(matrix.m(), space_dh.n_dofs());
  AssertDimension(matrix.n(), immersed_dh.n_dofs());
  Assert((dim1 <= dim0) && (dim0 <= spacedim), ExcMessage("This function can only work if dim1<=dim0<=spacedim"));
  AssertThrow(space_comps.size() <= 1 && immersed_comps.size() <= 1, ExcNotImplemented());

  auto &s = nmgpu_detail::session(&space_dh.get_triangulation(), &immersed_dh.get_triangulation());
  // the caller built cells_and_quads elsewhere, or the mesh changed since: take the geometry from the mappings given here
  if (!nmgpu_detail::grids_current<dim0, dim1, spacedim>(s, space_dh.get_triangulation(), immersed_dh.get_triangulation()))
    nmgpu_detail::set_grids<dim0, dim1, spacedim>(s, space_dh.get_triangulation(), space_mapping, immersed_dh.get_triangulation(),
                                                  immersed_mapping, nullptr, nullptr);
  nmgpu_detail::set_dofs<dim0, dim1, spacedim>(s, space_dh, immersed_dh, space_constraints, immersed_constraints);
  if (nmgpu_detail::is_own_result(s, cells_and_quads))
    nmgpu_detail::assemble_from_device_result(s);               // the quadrature never left the device
  else {
    nmgpu_detail::assemble_from_info<dim0, dim1, spacedim>(s, cells_and_quads);
    s.own.valid = false;
  }
  std::vector<typename Matrix::value_type> vals;
  nmgpu_detail::for_each_stored_row(s, true, [&](types::global_dof_index r, types::global_dof_index n, const std::vector<types::global_dof_index> &cols,
                                                 const double *v) {
    vals.assign(v, v + n);
    // zeros are skipped exactly like distribute_local_to_global does (constrained rows stay structural zeros)
    matrix.add(r, n, cols.data(), vals.data(), /*elide_zero_values=*/true, /*col_indices_are_sorted=*/true);
  });
  matrix.compress(VectorOperation::add); // reference :289
}

// Reference: code/include/coupling_utilities.h:306-398.  Adds
//   sum_q c(x_q) (penalty / h) phi_i(x_q) phi_j(x_q) w_q,   h = space_cell->diameter(),
// distributed through `space_constraints`, to `matrix` (which must already hold these positions, as the reference's
// system matrix does); like the reference it does not compress.
template <int dim0, int dim1, int spacedim, typename Matrix>
void assemble_nitsche_with_exact_intersections(
    const DoFHandler<dim0, spacedim> &space_dh,
    const std::vector<std::tuple<typename Triangulation<dim0, spacedim>::cell_iterator,
                                 typename Triangulation<dim1, spacedim>::cell_iterator, Quadrature<spacedim>>> &cells_and_quads,
    Matrix &matrix, const AffineConstraints<typename Matrix::value_type> &space_constraints, const ComponentMask &space_comps,
    const Mapping<dim0, spacedim> &space_mapping, const Function<spacedim, typename Matrix::value_type> &nitsche_coefficient,
    const double penalty)
{
  AssertDimension(matrix.m(), space_dh.n_dofs());
  AssertDimension(matrix.n(), space_dh.n_dofs());
  Assert((dim1 <= dim0) && (dim0 <= spacedim), ExcMessage("This function can only work if dim1<=dim0<=spacedim"));
  Assert(cells_and_quads.size() > 0,
         ExcMessage("The background and immersed mesh must overlap in order to use this function."));
  AssertThrow(space_comps.size() <= 1, ExcNotImplemented());

  auto &s = nmgpu_detail::session(&space_dh.get_triangulation(), &std::get<1>(cells_and_quads.front())->get_triangulation());
  if (!s.space_set || s.space_active_cells != space_dh.get_triangulation().n_active_cells())   // first use, or the mesh changed
    nmgpu_detail::set_space_grid<dim0, spacedim>(s, space_dh.get_triangulation(), space_mapping);
  nmgpu_detail::set_space_dofs<dim0, spacedim>(s, space_dh, space_constraints);
  nmgpu_detail::nitsche_from_info<dim0, dim1, spacedim, typename Matrix::value_type>(s, cells_and_quads, /*owned_only=*/true,
                                                                                     nitsche_coefficient, nullptr, penalty,
                                                                                     NM_NITSCHE_MATRIX);
  const types::global_dof_index n = space_dh.n_dofs();
  std::vector<std::uint64_t> rowptr(n + 1);
  nmgpu_detail::check(s, nm_get_nitsche(s.ctx, rowptr.data(), nullptr, nullptr, nullptr, NM_HOST));
  std::vector<std::uint32_t> col(rowptr[n]);
  std::vector<double> val(rowptr[n]);
  nmgpu_detail::check(s, nm_get_nitsche(s.ctx, nullptr, col.data(), val.data(), nullptr, NM_HOST));
  std::vector<types::global_dof_index> cols;
  std::vector<typename Matrix::value_type> vals;
  for (types::global_dof_index r = 0; r < n; ++r)
    if (rowptr[r + 1] > rowptr[r]) {
      cols.assign(col.begin() + rowptr[r], col.begin() + rowptr[r + 1]);
      vals.assign(val.begin() + rowptr[r], val.begin() + rowptr[r + 1]);
      matrix.add(r, static_cast<types::global_dof_index>(cols.size()), cols.data(), vals.data(), /*elide_zero_values=*/true,
                 /*col_indices_are_sorted=*/true);
    }
}

// Reference: code/include/coupling_utilities.h:400-481.  Adds
//   sum_q c(x_q) (penalty / h) phi_i(x_q) f(x_q) w_q
// distributed through `space_constraints`, to `rhs_vector`; no compress, like the reference.
template <int dim0, int dim1, int spacedim, typename Vector>
void create_nitsche_rhs_with_exact_intersections(
    const DoFHandler<dim0, spacedim> &space_dh,
    const std::vector<std::tuple<typename Triangulation<dim0, spacedim>::cell_iterator,
                                 typename Triangulation<dim1, spacedim>::cell_iterator, Quadrature<spacedim>>> &cells_and_quads,
    Vector &rhs_vector, const AffineConstraints<double> &space_constraints, const Mapping<dim0, spacedim> &space_mapping,
    const Function<spacedim, double> &rhs_function, const Function<spacedim, double> &coefficient, const double penalty)
{
  AssertDimension(rhs_vector.size(), space_dh.n_dofs());
  Assert(dim1 <= dim0, ExcMessage("This function can only work if dim1<=dim0"));
  if (cells_and_quads.empty())
    return;
  auto &s = nmgpu_detail::session(&space_dh.get_triangulation(), &std::get<1>(cells_and_quads.front())->get_triangulation());
  if (!s.space_set || s.space_active_cells != space_dh.get_triangulation().n_active_cells())   // first use, or the mesh changed
    nmgpu_detail::set_space_grid<dim0, spacedim>(s, space_dh.get_triangulation(), space_mapping);
  nmgpu_detail::set_space_dofs<dim0, spacedim>(s, space_dh, space_constraints);
  // the reference's rhs loop takes every active cell (:428), not only locally owned ones; the arrays sent to the GPU
  // hold the locally owned cells, which is the same set on one rank
  nmgpu_detail::nitsche_from_info<dim0, dim1, spacedim, double>(s, cells_and_quads, /*owned_only=*/false, coefficient, &rhs_function,
                                                                penalty, NM_NITSCHE_RHS);
  std::vector<double> rhs(space_dh.n_dofs());
  nmgpu_detail::check(s, nm_get_nitsche(s.ctx, nullptr, nullptr, nullptr, rhs.data(), NM_HOST));
  // one batched add: off-process rows are legal for the distributed vector types and reach their owner in the caller's
  // compress(VectorOperation::add), exactly like the element-wise adds of distribute_local_to_global
  std::vector<types::global_dof_index> rows;
  std::vector<typename Vector::value_type> vals;
  for (types::global_dof_index r = 0; r < space_dh.n_dofs(); ++r)
    if (rhs[r] != 0.) {
      rows.push_back(r);
      vals.push_back(static_cast<typename Vector::value_type>(rhs[r]));
    }
  if (!rows.empty())
    rhs_vector.add(rows, vals);
}

// Frees every GPU context (and its device memory) this header created.  The contexts are otherwise kept for the life
// of the process, so that the calls of one refinement cycle find the device-resident results of the previous call.
inline void nmgpu_release() { nmgpu_detail::sessions().clear(); }

} // namespace NonMatching
} // namespace dealii

#endif
