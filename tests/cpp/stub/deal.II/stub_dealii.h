// Minimal, functional stand-in for the handful of deal.II 9.4 types that include/coupling_utilities.h
// touches.  TEST INFRASTRUCTURE: deal.II is not installed in this image, so the drop-in header is
// compile-checked and exercised against this stub (same member names and call shapes as deal.II).
#pragma once
#include <array>
#include <cstddef>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace dealii {
namespace types { using global_dof_index = unsigned int; }

struct ExcMessage { std::string m; explicit ExcMessage(std::string s) : m(std::move(s)) {} };
struct ExcNotImplemented { std::string m = "ExcNotImplemented"; };
#define AssertThrow(cond, exc) do { if (!(cond)) throw std::runtime_error((exc).m); } while (0)
#define Assert(cond, exc) AssertThrow(cond, exc)
#define AssertDimension(a, b) AssertThrow((a) == (b), ::dealii::ExcMessage("dimension mismatch"))

namespace VectorOperation { enum values { unknown, insert, add }; }

template <int dim> struct Point {
  double c[dim] = {};
  double &operator[](unsigned i) { return c[i]; }
  const double &operator[](unsigned i) const { return c[i]; }
};

template <int dim> class Quadrature {
public:
  Quadrature() = default;
  Quadrature(std::vector<Point<dim>> p, std::vector<double> w) : pts(std::move(p)), wts(std::move(w)) {}
  unsigned int size() const { return wts.size(); }
  const Point<dim> &point(unsigned k) const { return pts[k]; }
  double weight(unsigned k) const { return wts[k]; }
  const std::vector<double> &get_weights() const { return wts; }
  const std::vector<Point<dim>> &get_points() const { return pts; }
private:
  std::vector<Point<dim>> pts;
  std::vector<double> wts;
};

template <int dim, typename number = double> class Function {
public:
  virtual ~Function() = default;
  virtual number value(const Point<dim> &p, const unsigned int component = 0) const = 0;
  virtual void value_list(const std::vector<Point<dim>> &points, std::vector<number> &values, const unsigned int component = 0) const
  {
    for (std::size_t i = 0; i < points.size(); ++i) values[i] = value(points[i], component);
  }
};
namespace Functions {
template <int dim, typename number = double> class ConstantFunction : public Function<dim, number> {
public:
  explicit ConstantFunction(number v) : v(v) {}
  number value(const Point<dim> &, const unsigned int = 0) const override { return v; }
private:
  number v;
};
} // namespace Functions

template <typename T> class Vector {
public:
  using value_type = T;
  explicit Vector(std::size_t n) : v(n, T(0)) {}
  void add(const std::vector<types::global_dof_index> &indices, const std::vector<T> &values)
  {
    for (std::size_t k = 0; k < indices.size(); ++k) v[indices[k]] += values[k];
  }
  std::size_t size() const { return v.size(); }
  T &operator()(std::size_t i) { return v[i]; }
  const T &operator()(std::size_t i) const { return v[i]; }
private:
  std::vector<T> v;
};

template <int dim, int spacedim> class DoFHandler;

template <int dim, int spacedim = dim> class Triangulation {
public:
  static constexpr unsigned nv = 1u << dim;
  std::vector<std::array<Point<spacedim>, nv>> cells;
  struct Accessor {
    const Triangulation *t = nullptr;
    const DoFHandler<dim, spacedim> *dh = nullptr;
    unsigned idx = 0;
    unsigned active_cell_index() const { return idx; }
    bool is_locally_owned() const { return true; }
    bool is_active() const { return true; }
    const Triangulation &get_triangulation() const { return *t; }
    int level() const { return 0; }
    int index() const { return int(idx); }
    void get_dof_indices(std::vector<types::global_dof_index> &v) const;
  };
  struct cell_iterator {
    Accessor a;
    const Accessor *operator->() const { return &a; }
    const Accessor &operator*() const { return a; }
    bool operator!=(const cell_iterator &o) const { return a.idx != o.a.idx; }
    cell_iterator &operator++() { ++a.idx; return *this; }
  };
  using active_cell_iterator = cell_iterator;
  // like deal.II's IteratorRange: dereferencing the range iterator yields the cell iterator itself
  struct RangeIt {
    cell_iterator it;
    const cell_iterator &operator*() const { return it; }
    RangeIt &operator++() { ++it; return *this; }
    bool operator!=(const RangeIt &o) const { return it != o.it; }
  };
  struct Range {
    cell_iterator b, e;
    RangeIt begin() const { return {b}; }
    RangeIt end() const { return {e}; }
  };
  unsigned n_active_cells() const { return cells.size(); }
  Range active_cell_iterators() const { return {{{this, nullptr, 0}}, {{this, nullptr, unsigned(cells.size())}}}; }
};

template <int dim, int spacedim = dim> class Mapping {
public:
  std::vector<Point<spacedim>> get_vertices(const typename Triangulation<dim, spacedim>::cell_iterator &c) const
  {
    const auto &v = c->t->cells[c->idx];
    return std::vector<Point<spacedim>>(v.begin(), v.end());
  }
};

namespace GridTools {
template <int dim, int spacedim = dim> class Cache {
public:
  Cache(const Triangulation<dim, spacedim> &t, const Mapping<dim, spacedim> &m) : tria(&t), map(&m) {}
  const Triangulation<dim, spacedim> &get_triangulation() const { return *tria; }
  const Mapping<dim, spacedim> &get_mapping() const { return *map; }
private:
  const Triangulation<dim, spacedim> *tria;
  const Mapping<dim, spacedim> *map;
};
} // namespace GridTools

template <int dim, int spacedim = dim> class FiniteElement {
public:
  unsigned dpc = 1;
  unsigned degree = 0;
  std::vector<Point<dim>> unit_support_points;   // in the element's dof order (deal.II: FiniteElement::get_unit_support_points())
  unsigned n_components() const { return 1; }
  unsigned n_dofs_per_cell() const { return dpc; }
  const std::vector<Point<dim>> &get_unit_support_points() const { return unit_support_points; }
};

template <int dim, int spacedim = dim> class DoFHandler {
public:
  const Triangulation<dim, spacedim> *tria = nullptr;
  FiniteElement<dim, spacedim> fe;
  std::vector<types::global_dof_index> table;   // [n_cells][dofs_per_cell]
  types::global_dof_index ndofs = 0;
  const Triangulation<dim, spacedim> &get_triangulation() const { return *tria; }
  const FiniteElement<dim, spacedim> &get_fe() const { return fe; }
  types::global_dof_index n_dofs() const { return ndofs; }
  typename Triangulation<dim, spacedim>::Range active_cell_iterators() const
  {
    return {{{tria, this, 0}}, {{tria, this, unsigned(tria->cells.size())}}};
  }
};

template <int dim, int spacedim>
void Triangulation<dim, spacedim>::Accessor::get_dof_indices(std::vector<types::global_dof_index> &v) const
{
  for (unsigned i = 0; i < v.size(); ++i) v[i] = dh->table[std::size_t(idx) * dh->fe.dpc + i];
}

template <typename number = double> class AffineConstraints {
public:
  using size_type = types::global_dof_index;
  std::map<size_type, std::vector<std::pair<size_type, number>>> lines;
  bool is_constrained(size_type i) const { return lines.count(i) != 0; }
  const std::vector<std::pair<size_type, number>> *get_constraint_entries(size_type i) const
  {
    auto it = lines.find(i);
    return it == lines.end() ? nullptr : &it->second;
  }
  size_type n_constraints() const { return lines.size(); }
  // deal.II: AffineConstraints::get_lines() -> range of ConstraintLine {index, entries, inhomogeneity}, sorted by index
  struct ConstraintLine {
    size_type index;
    std::vector<std::pair<size_type, number>> entries;
    number inhomogeneity;
  };
  std::vector<ConstraintLine> get_lines() const
  {
    std::vector<ConstraintLine> r;
    for (const auto &l : lines) r.push_back({l.first, l.second, number(0)});
    return r;
  }
};

class ComponentMask {
public:
  unsigned size() const { return 0; }
};

class DynamicSparsityPattern {
public:
  DynamicSparsityPattern(unsigned r, unsigned c) : rows(r), cols(c), entries(r) {}
  unsigned n_rows() const { return rows; }
  unsigned n_cols() const { return cols; }
  template <typename It> void add_entries(unsigned row, It b, It e, bool = false) { entries[row].insert(entries[row].end(), b, e); }
  unsigned rows, cols;
  std::vector<std::vector<types::global_dof_index>> entries;
};

template <typename T> class StubSparseMatrix {
public:
  using value_type = T;
  StubSparseMatrix(unsigned r, unsigned c) : rows(r), cols(c), data(r) {}
  unsigned m() const { return rows; }
  unsigned n() const { return cols; }
  void add(types::global_dof_index row, types::global_dof_index n_cols, const types::global_dof_index *c, const T *v, bool elide_zero = true,
           bool = false)
  {
    for (unsigned k = 0; k < n_cols; ++k)
      if (!elide_zero || v[k] != T(0)) data[row][c[k]] += v[k];
  }
  void compress(VectorOperation::values) { compressed = true; }
  unsigned rows, cols;
  std::vector<std::map<types::global_dof_index, T>> data;
  bool compressed = false;
};
// Append-only sink with the add()/compress() shape of a deal.II matrix: what a timing run needs (no per-entry map).
template <typename T> class StubFlatMatrix {
public:
  using value_type = T;
  StubFlatMatrix(unsigned r, unsigned c) : rows(r), cols(c) {}
  unsigned m() const { return rows; }
  unsigned n() const { return cols; }
  void add(types::global_dof_index row, types::global_dof_index n_cols, const types::global_dof_index *c, const T *v, bool elide_zero = true,
           bool = false)
  {
    for (unsigned k = 0; k < n_cols; ++k)
      if (!elide_zero || v[k] != T(0)) { r_.push_back(row); c_.push_back(c[k]); v_.push_back(v[k]); }
  }
  void compress(VectorOperation::values) { compressed = true; }
  unsigned rows, cols;
  std::vector<types::global_dof_index> r_, c_;
  std::vector<T> v_;
  bool compressed = false;
};
} // namespace dealii
