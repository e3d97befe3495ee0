// Host-side unit test of the brick-index helpers (non_matching_test_suite_b200/csrc/nm_brick.cuh): the exact per-axis
// index range and the per-brick query mask against brute force, and a whole query on a two-level band of grid cells
// against the closed-box test of the reference (code/src/coupling_utilities.cc:53-55) on the same doubles.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <set>
#include <tuple>
#include <vector>

#include "../../non_matching_test_suite_b200/csrc/nm_brick.cuh"

using namespace nmbrick;

static int fails = 0;
#define CHECK(c)                                              \
  do {                                                        \
    if (!(c)) { if (fails++ < 10) printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); } \
  } while (0)

static void test_axis_range(std::mt19937_64 &rng)
{
  std::uniform_real_distribution<double> U(0, 1);
  for (int rep = 0; rep < 200000; ++rep) {
    const double org = (U(rng) - 0.5) * 4;
    const double g = (rep & 1) ? ldexp(1.0, -(int)(U(rng) * 20)) : (0.1 + U(rng)) * ldexp(1.0, -(int)(U(rng) * 20));
    const int top = 200;
    double qlo, qhi;
    const int mode = rep % 5;
    if (mode == 0) { const int i = (int)(U(rng) * 150); qlo = grid_corner(org, i, g); qhi = grid_corner(org, i + (int)(U(rng) * 4), g); }  // touching corners
    else if (mode == 1) { const int i = (int)(U(rng) * 150); qhi = grid_corner(org, i, g); qlo = qhi - U(rng) * 3 * g; }
    else { qlo = org + (U(rng) * 220 - 10) * g; qhi = qlo + U(rng) * 5 * g; }
    int L, H;
    const bool ok = axis_range(org, g, 1.0 / g, (int)top, qlo, qhi, L, H);
    // brute force
    int bl = 1, bh = 0;
    bool any = false;
    for (int i = 0; i <= top; ++i) {
      const double lo = grid_corner(org, i, g), hi = grid_corner(org, i + 1, g);
      if (lo <= qhi && qlo <= hi) { if (!any) bl = i; bh = i; any = true; }
    }
    CHECK(ok == any);
    if (ok && any) { CHECK(L == bl); CHECK(H == bh); }
  }
}

template <int D>
static void test_query_mask(std::mt19937_64 &rng)
{
  constexpr int SH = BrickShape<D>::SH, E = BrickShape<D>::EDGE;
  std::uniform_int_distribution<int> I(0, 40);
  for (int rep = 0; rep < 100000; ++rep) {
    int L[3] = {0, 0, 0}, H[3] = {0, 0, 0}, b[3] = {0, 0, 0};
    for (int k = 0; k < D; ++k) { L[k] = I(rng); H[k] = L[k] + I(rng) % 7; b[k] = I(rng) / E + (I(rng) % 3 - 1); if (b[k] < 0) b[k] = 0; }
    const unsigned long long m = brick_query_mask<D>(L, H, b[0], b[1], b[2]);
    unsigned long long ref = 0;
    for (int bit = 0; bit < 64; ++bit) {
      int idx[3] = {0, 0, 0};
      if (D == 3) { idx[0] = (b[0] << SH) + (bit & 3); idx[1] = (b[1] << SH) + ((bit >> 2) & 3); idx[2] = (b[2] << SH) + ((bit >> 4) & 3); }
      else { idx[0] = (b[0] << SH) + (bit & 7); idx[1] = (b[1] << SH) + ((bit >> 3) & 7); }
      bool in = true;
      for (int k = 0; k < D; ++k) in = in && idx[k] >= L[k] && idx[k] <= H[k];
      if (in) ref |= 1ULL << bit;
    }
    CHECK(m == ref);
  }
}

// a band of grid-aligned cells on two levels around a circle / sphere; queries = random boxes near the band
template <int D>
static void test_whole_query(std::mt19937_64 &rng)
{
  constexpr int SH = BrickShape<D>::SH;
  std::uniform_real_distribution<double> U(0, 1);
  const double g0 = 1.0 / 64, org0[3] = {-1.0, -1.0, -1.0};   // band cells start at arbitrary fine indices: min corner is generally NOT on the coarse grid
  struct Cell { int level; long long idx[3]; double lo[3], hi[3]; };
  std::vector<Cell> cells;
  const int n0 = 128;
  auto dist = [&](const double *c) { double r = 0; for (int k = 0; k < D; ++k) r += (c[k] - 0.1) * (c[k] - 0.1); return sqrt(r) - 0.6; };
  for (int level = 0; level < 2; ++level) {
    const double g = g0 * (1 << level);
    const int n = n0 >> level;
    long long idx[3] = {0, 0, 0};
    const long long tot = D == 3 ? (long long)n * n * n : (long long)n * n;
    for (long long f = 0; f < tot; ++f) {
      idx[0] = f % n; idx[1] = (f / n) % n; idx[2] = D == 3 ? f / ((long long)n * n) : 0;
      double c[3];
      for (int k = 0; k < D; ++k) c[k] = org0[k] + (idx[k] + 0.5) * g;
      const double s = fabs(dist(c));
      const bool fine = s < 2.5 * g0, coarse = s >= 2.5 * g0 && s < 6 * g0;
      if ((level == 0 && fine) || (level == 1 && coarse)) {
        Cell C; C.level = level;
        for (int k = 0; k < 3; ++k) { C.idx[k] = k < D ? idx[k] : 0; C.lo[k] = k < D ? grid_corner(org0[k], idx[k], g) : 0; C.hi[k] = k < D ? grid_corner(org0[k], idx[k] + 1, g) : 0; }
        cells.push_back(C);
      }
    }
  }
  // ONE origin for both levels (aligned_origin): the corner of a coarsest cell moved down by whole coarsest cells
  double lo_min[3] = {1e300, 1e300, 1e300}, lo_coarse[3] = {0, 0, 0}, gmax = 0;
  for (auto &c : cells) {
    for (int k = 0; k < D; ++k) lo_min[k] = fmin(lo_min[k], c.lo[k]);
    if (c.hi[0] - c.lo[0] > gmax) { gmax = c.hi[0] - c.lo[0]; for (int k = 0; k < D; ++k) lo_coarse[k] = c.lo[k]; }
  }
  double org[3] = {0, 0, 0};
  for (int k = 0; k < D; ++k) org[k] = aligned_origin(lo_min[k], lo_coarse[k], gmax);
  std::map<std::tuple<int, int, int, int>, std::pair<unsigned long long, std::vector<int>>> bricks;
  for (size_t i = 0; i < cells.size(); ++i) {
    auto &c = cells[i];
    const int level = level_of_extent(c.hi[0] - c.lo[0], g0);
    CHECK(level == c.level);
    const double g = g0 * (1 << c.level);
    int j[3] = {0, 0, 0};
    for (int k = 0; k < D; ++k) {
      j[k] = (int)llrint((c.lo[k] - org[k]) / g);
      CHECK(j[k] >= 0 && grid_corner(org[k], j[k], g) == c.lo[k] && grid_corner(org[k], j[k] + 1, g) == c.hi[k]);
    }
    const unsigned bit = D == 3 ? (unsigned)((j[0] & 3) | ((j[1] & 3) << 2) | ((j[2] & 3) << 4)) : (unsigned)((j[0] & 7) | ((j[1] & 7) << 3));
    auto &b = bricks[{c.level, j[0] >> SH, j[1] >> SH, j[2] >> SH}];
    CHECK(!((b.first >> bit) & 1));
    b.first |= 1ULL << bit;
    b.second.resize(64, -1);
    b.second[bit] = (int)i;
  }
  size_t total = 0;
  for (int q = 0; q < 3000; ++q) {
    double qlo[3] = {0, 0, 0}, qhi[3] = {0, 0, 0};
    const Cell &c = cells[(size_t)(U(rng) * cells.size()) % cells.size()];
    for (int k = 0; k < D; ++k) {
      const int m = (int)(U(rng) * 4);
      qlo[k] = m == 0 ? c.hi[k] : (m == 1 ? c.lo[k] - U(rng) * 2 * g0 : c.lo[k] + (U(rng) - 0.5) * g0);   // touching from above / overlapping
      qhi[k] = m == 0 ? c.hi[k] + U(rng) * 2 * g0 : (m == 1 ? c.lo[k] : qlo[k] + U(rng) * 3 * g0);
    }
    std::set<int> got, ref;
    for (int level = 0; level < 2; ++level) {
      const double g = g0 * (1 << level);
      int L[3] = {0, 0, 0}, H[3] = {0, 0, 0};
      bool ok = true;
      for (int k = 0; k < D; ++k) ok = ok && axis_range(org[k], g, 1.0 / g, (1 << 19) - 2, qlo[k], qhi[k], L[k], H[k]);
      if (!ok) continue;
      for (int bz = L[2] >> SH; bz <= (H[2] >> SH); ++bz)
        for (int by = L[1] >> SH; by <= (H[1] >> SH); ++by)
          for (int bx = L[0] >> SH; bx <= (H[0] >> SH); ++bx) {
            auto it = bricks.find({level, bx, by, bz});
            if (it == bricks.end()) continue;
            unsigned long long hits = it->second.first & brick_query_mask<D>(L, H, bx, by, bz);
            for (int bit = 0; bit < 64; ++bit) if ((hits >> bit) & 1) got.insert(it->second.second[bit]);
          }
    }
    for (size_t i = 0; i < cells.size(); ++i) {
      bool hit = true;
      for (int k = 0; k < D; ++k) hit = hit && cells[i].lo[k] <= qhi[k] && qlo[k] <= cells[i].hi[k];
      if (hit) ref.insert((int)i);
    }
    CHECK(got == ref);
    total += ref.size();
  }
  printf("  D=%d: %zu cells, %zu bricks, %zu candidate hits checked\n", D, cells.size(), bricks.size(), total);
}

int main()
{
  std::mt19937_64 rng(12345);
  for (double x : {0.0, -0.0, 1.5, -1.5, 1e-300, -1e300}) CHECK(ord_decode(ord_encode(x)) == x);
  CHECK(ord_encode(-2.0) < ord_encode(-1.0) && ord_encode(-1.0) < ord_encode(0.0) && ord_encode(0.0) < ord_encode(1e-300) && ord_encode(1.0) < ord_encode(2.0));
  test_axis_range(rng);
  test_query_mask<2>(rng);
  test_query_mask<3>(rng);
  test_whole_query<2>(rng);
  test_whole_query<3>(rng);
  printf(fails ? "test_brick: %d FAILURES\n" : "test_brick: ok\n", fails);
  return fails ? 1 : 0;
}
