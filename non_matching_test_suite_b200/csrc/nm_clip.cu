// Subsystems (2) and (3a): per-pair FP64 intersection + quadrature + local coupling vector.
//
// Replaces, for every candidate pair,
//   CGALWrappers::compute_intersection_of_cells<2,1,2> / <3,2,3>   code/src/intersections.cc:358-393, 437-509
//   QGaussSimplexExt::mapped_quadrature                             code/src/intersections.cc:693-746
//   the acceptance test sum(weights) > tol                          code/src/coupling_utilities.cc:59-65
//   local(i,0) += phi_i(F^-1(x_q)) * 1 * w_q                        code/include/coupling_utilities.h:248-274
// by one thread per candidate: the immersed simplex is clipped against the axis-aligned
// background cell (Liang-Barsky for a segment, Sutherland-Hodgman for a triangle), the
// result is fanned into simplices, each gets the reference's QGaussSimplex rule, sub-simplices
// with |J| < 1e-12 are dropped as in the reference, and the Q1 x P0 products are accumulated
// without the points ever leaving registers.  Work is done in cell-local coordinates
// (x - lo), which is exact in sign and keeps full precision for band cells of size 1e-6.
#include "nm_common.cuh"
#include "nm_predicates.cuh"
#include "nm_clip_core.cuh"
using namespace nmclip;

namespace {

struct Rule {
  int n;
  double x[16][2];
  double w[16];
};
__constant__ Rule c_rule;

// QGaussSimplex<dim>(n_points_1D) of deal.II 9.4 for the sizes the path uses
// (code/src/intersections.cc:693-695: the `degree` argument is passed as n_points_1D).
bool make_rule(int dim1, unsigned n, Rule &r)
{
  memset(&r, 0, sizeof(r));
  if (dim1 == 1) {  // Gauss-Legendre on [0,1]
    switch (n) {
      case 1: r.n = 1; r.x[0][0] = 0.5; r.w[0] = 1.0; return true;
      case 2: { double a = 0.5 / sqrt(3.0); r.n = 2; r.x[0][0] = 0.5 - a; r.x[1][0] = 0.5 + a; r.w[0] = r.w[1] = 0.5; return true; }
      case 3: { double a = 0.5 * sqrt(0.6); r.n = 3; r.x[0][0] = 0.5 - a; r.x[1][0] = 0.5; r.x[2][0] = 0.5 + a;
                r.w[0] = r.w[2] = 5.0 / 18.0; r.w[1] = 4.0 / 9.0; return true; }
      case 4: { double s = sqrt(1.2), a = 0.5 * sqrt(3.0 / 7.0 - 2.0 / 7.0 * s), b = 0.5 * sqrt(3.0 / 7.0 + 2.0 / 7.0 * s);
                double wa = (18.0 + sqrt(30.0)) / 72.0, wb = (18.0 - sqrt(30.0)) / 72.0;
                r.n = 4; r.x[0][0] = 0.5 - b; r.x[1][0] = 0.5 - a; r.x[2][0] = 0.5 + a; r.x[3][0] = 0.5 + b;
                r.w[0] = r.w[3] = wb; r.w[1] = r.w[2] = wa; return true; }
    }
    if (n >= 5 && n <= 8) {  // Newton on P_n: the rules the drivers ask for at degree p >= 2 (n = 2p + 1, 2d3d/lagrange_multiplier.cc:750-753)
      r.n = (int)n;
      for (int i = 0; i < (int)n; ++i) {
        double z = cos(3.14159265358979323846 * (i + 0.75) / (n + 0.5)), pp = 1.0;
        for (int it = 0; it < 100; ++it) {
          double p1 = 1.0, p2 = 0.0;
          for (int j = 0; j < (int)n; ++j) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0); }
          pp = n * (z * p1 - p2) / (z * z - 1.0);
          const double dz = p1 / pp;
          z -= dz;
          if (fabs(dz) < 1e-16) break;
        }
        {
          double p1 = 1.0, p2 = 0.0;
          for (int j = 0; j < (int)n; ++j) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0); }
          pp = n * (z * p1 - p2) / (z * z - 1.0);
        }
        r.x[n - 1 - i][0] = 0.5 + 0.5 * z;
        r.w[n - 1 - i] = 1.0 / ((1.0 - z * z) * pp * pp);
      }
      return true;
    }
    return false;
  }
  if (n == 1) { r.n = 1; r.x[0][0] = r.x[0][1] = 1.0 / 3.0; r.w[0] = 0.5; return true; }
  if (n == 2) {
    r.n = 3;
    const double a = 1.0 / 6.0, b = 2.0 / 3.0;
    r.x[0][0] = a; r.x[0][1] = a; r.x[1][0] = b; r.x[1][1] = a; r.x[2][0] = a; r.x[2][1] = b;
    r.w[0] = r.w[1] = r.w[2] = 1.0 / 6.0;
    return true;
  }
  if (n == 3) {  // 7-point, degree 5
    r.n = 7;
    const double s = sqrt(15.0), a = (6.0 - s) / 21.0, b = (6.0 + s) / 21.0;
    const double wa = 0.5 * (155.0 - s) / 1200.0, wb = 0.5 * (155.0 + s) / 1200.0;
    r.x[0][0] = r.x[0][1] = 1.0 / 3.0; r.w[0] = 0.5 * 9.0 / 40.0;
    r.x[1][0] = a; r.x[1][1] = a; r.x[2][0] = 1 - 2 * a; r.x[2][1] = a; r.x[3][0] = a; r.x[3][1] = 1 - 2 * a;
    r.x[4][0] = b; r.x[4][1] = b; r.x[5][0] = 1 - 2 * b; r.x[5][1] = b; r.x[6][0] = b; r.x[6][1] = 1 - 2 * b;
    r.w[1] = r.w[2] = r.w[3] = wa; r.w[4] = r.w[5] = r.w[6] = wb;
    return true;
  }
  return false;
}

__device__ __constant__ int c_triples[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};

// ---------------------------------------------------------------------------------------
// quad split (rule R2) + immersed cell measure
// ---------------------------------------------------------------------------------------
__global__ void split_kernel(uint64_t n, const double *__restrict__ verts, uint8_t *__restrict__ code_out,
                             double *__restrict__ measure, unsigned long long *n_ties)
{
  uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  double q[4][3];
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int k = 0; k < 3; ++k) q[v][k] = verts[m * 12 + v * 3 + k];
  const double *P[4] = {q[0], q[1], q[2], q[3]};
  int code = -1;
  // points coinciding in projection: CGAL keeps one of them (which one depends on its
  // spatial sort) -> flagged; the later index is dropped
  for (int i = 0; i < 4 && code < 0; ++i)
    for (int j = i + 1; j < 4 && code < 0; ++j)
      if (P[i][0] == P[j][0] && P[i][1] == P[j][1]) {
        int keep[3], c = 0;
        for (int k = 0; k < 4; ++k) if (k != j) keep[c++] = k;
        const double *a = P[keep[0]], *b = P[keep[1]], *cc = P[keep[2]];
        bool dup = (a[0] == b[0] && a[1] == b[1]) || (b[0] == cc[0] && b[1] == cc[1]) || (a[0] == cc[0] && a[1] == cc[1]);
        int o = dup ? 0 : nmpred::orient2d(a, b, cc);
        code = (dup || o == 0 || o == 2) ? 0x80 : (0x80 | (1 << j));
      }
  if (code < 0) {
    code = 0;
    bool tie = false;
    for (int t = 0; t < 4; ++t) {
      const double *a = P[c_triples[t][0]], *b = P[c_triples[t][1]], *c = P[c_triples[t][2]];
      int o = nmpred::orient2d(a, b, c);
      if (o == 2) { tie = true; continue; }
      if (o == 0) continue;
      int ic = nmpred::incircle(a, b, c, P[t]);
      if (ic == 2) { tie = true; continue; }
      ic *= o;
      if (ic < 0) code |= 1 << t;
      else if (ic == 0) tie = true;
    }
    if (tie) code = 0x80 | (1 << 2) | (1 << 1);
  }
  code_out[m] = (uint8_t)code;
  if (code & 0x80) atomicAdd(n_ties, 1ULL);
  double area = 0;
  for (int t = 0; t < 4; ++t)
    if ((code >> t) & 1) {
      const double *a = P[c_triples[t][0]], *b = P[c_triples[t][1]], *c = P[c_triples[t][2]];
      double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
      double nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
      area += 0.5 * sqrt(nx * nx + ny * ny + nz * nz);
    }
  measure[m] = area;
}

__global__ void segment_measure_kernel(uint64_t n, const double *__restrict__ verts, double *__restrict__ measure)
{
  uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  const double *v = verts + m * 4;
  measure[m] = hypot(v[2] - v[0], v[3] - v[1]);
}

// ---------------------------------------------------------------------------------------
// geometry helpers (cell-local coordinates: box = [0, H])
// ---------------------------------------------------------------------------------------
#define NM_MAXV 12

struct Poly {
  double x[NM_MAXV], y[NM_MAXV], z[NM_MAXV];
  int n;
};

// one Sutherland-Hodgman step against the closed half-space  s * (p[axis] - bound) >= 0
template <int AXIS, int UPPER>
__device__ inline void clip_step(const Poly &in, Poly &out, double bound)
{
  int m = 0;
  const int n = in.n;
  for (int i = 0; i < n; ++i) {
    const int j = (i + 1 == n) ? 0 : i + 1;
    const double pc = AXIS == 0 ? in.x[i] : (AXIS == 1 ? in.y[i] : in.z[i]);
    const double qc = AXIS == 0 ? in.x[j] : (AXIS == 1 ? in.y[j] : in.z[j]);
    const double dp = UPPER ? bound - pc : pc - bound;
    const double dq = UPPER ? bound - qc : qc - bound;
    if (dp >= 0.0) { out.x[m] = in.x[i]; out.y[m] = in.y[i]; out.z[m] = in.z[i]; ++m; }
    if ((dp > 0.0 && dq < 0.0) || (dp < 0.0 && dq > 0.0)) {
      const double t = dp / (dp - dq);
      double nx = in.x[i] + t * (in.x[j] - in.x[i]);
      double ny = in.y[i] + t * (in.y[j] - in.y[i]);
      double nz = in.z[i] + t * (in.z[j] - in.z[i]);
      // the new vertex lies on the plane by construction: pin it there so that later
      // inside-tests see it as inside (no sliver re-clipping)
      if (AXIS == 0) nx = bound; else if (AXIS == 1) ny = bound; else nz = bound;
      if (m < NM_MAXV) { out.x[m] = nx; out.y[m] = ny; out.z[m] = nz; ++m; }
    }
  }
  out.n = m;
}


template <class F>
__device__ inline void triangle_in_box(const double (*tri)[3], const double *H, F &&visit_subtriangle)
{
  // trivial reject on the triangle's own box (candidates come from the quad's box)
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double mn = fmin(tri[0][k], fmin(tri[1][k], tri[2][k]));
    const double mx = fmax(tri[0][k], fmax(tri[1][k], tri[2][k]));
    if (mx < 0.0 || mn > H[k]) return;
  }
  Poly A, B;
  A.n = 3;
#pragma unroll
  for (int v = 0; v < 3; ++v) { A.x[v] = tri[v][0]; A.y[v] = tri[v][1]; A.z[v] = tri[v][2]; }
  clip_step<0, 0>(A, B, 0.0);  if (B.n < 3) return;
  clip_step<0, 1>(B, A, H[0]); if (A.n < 3) return;
  clip_step<1, 0>(A, B, 0.0);  if (B.n < 3) return;
  clip_step<1, 1>(B, A, H[1]); if (A.n < 3) return;
  clip_step<2, 0>(A, B, 0.0);  if (B.n < 3) return;
  clip_step<2, 1>(B, A, H[2]); if (A.n < 3) return;
  for (int j = 1; j + 1 < A.n; ++j) {
    const double a[3] = {A.x[0], A.y[0], A.z[0]};
    const double e1[3] = {A.x[j] - a[0], A.y[j] - a[1], A.z[j] - a[2]};
    const double e2[3] = {A.x[j + 1] - a[0], A.y[j + 1] - a[1], A.z[j + 1] - a[2]};
    visit_subtriangle(a, e1, e2);
  }
}

// Liang-Barsky on the closed box [0,H]; false for empty or single-point results
__device__ inline bool segment_in_box(const double *p0, const double *p1, const double *H, double &t0, double &t1)
{
  t0 = 0.0; t1 = 1.0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double d = p1[k] - p0[k];
    if (d == 0.0) {
      if (p0[k] < 0.0 || p0[k] > H[k]) return false;
    } else {
      double ta = (0.0 - p0[k]) / d, tb = (H[k] - p0[k]) / d;
      if (ta > tb) { const double s = ta; ta = tb; tb = s; }
      t0 = fmax(t0, ta); t1 = fmin(t1, tb);
      if (t0 > t1) return false;
    }
  }
  return t0 < t1;
}

struct PairAcc {
  double sumw;
  double loc[8];
  unsigned nq;
  unsigned dropped;
};

__device__ inline void q1_add_3d(PairAcc &o, double u0, double u1, double u2, double w)
{
  const double a0 = w - w * u0, a1 = w * u0;
  const double b00 = a0 - a0 * u1, b10 = a1 - a1 * u1, b01 = a0 * u1, b11 = a1 * u1;
  o.loc[4] += b00 * u2; o.loc[5] += b10 * u2; o.loc[6] += b01 * u2; o.loc[7] += b11 * u2;
  o.loc[0] += b00 - b00 * u2; o.loc[1] += b10 - b10 * u2; o.loc[2] += b01 - b01 * u2; o.loc[3] += b11 - b11 * u2;
  o.sumw += w;
}
__device__ inline void q1_add_2d(PairAcc &o, double u0, double u1, double w)
{
  const double a0 = w - w * u0, a1 = w * u0;
  o.loc[2] += a0 * u1; o.loc[3] += a1 * u1;
  o.loc[0] += a0 - a0 * u1; o.loc[1] += a1 - a1 * u1;
  o.sumw += w;
}

// sub-triangle (a, a+e1, a+e2) in cell-local coordinates; `emit(x_local[3], w)` per point
template <class E>
__device__ inline void subtriangle_rule(const double *a, const double *e1, const double *e2, double tol, PairAcc &o, E &&emit)
{
  const double nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
  const double J2 = nx * nx + ny * ny + nz * nz;
  if (!(0.25 * J2 > tol * tol)) return;                 // squared_area > tol^2   intersections.cc:477,493
  const double J = sqrt(J2);                             // 2 * area = |det B|      intersections.cc:702-707
  if (J < 1e-12) { o.dropped++; return; }                // intersections.cc:710
  const int nq = c_rule.n;
  for (int q = 0; q < nq; ++q) {
    const double s = c_rule.x[q][0], t = c_rule.x[q][1];
    const double x[3] = {a[0] + s * e1[0] + t * e2[0], a[1] + s * e1[1] + t * e2[1], a[2] + s * e1[2] + t * e2[2]};
    emit(x, J * c_rule.w[q]);
  }
  o.nq += nq;
}

template <class E>
__device__ inline void pair_3d(const double *__restrict__ box, const double *__restrict__ quad, int code, double tol, PairAcc &o, E &&emit)
{
  const double lo[3] = {box[0], box[1], box[2]};
  const double H[3] = {box[4] - lo[0], box[5] - lo[1], box[6] - lo[2]};
  double q[4][3];
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int k = 0; k < 3; ++k) q[v][k] = quad[v * 3 + k] - lo[k];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    if (!((code >> t) & 1)) continue;
    // triple t omits vertex t
    const int i0 = t == 0 ? 1 : 0, i1 = t <= 1 ? 2 : 1, i2 = t <= 2 ? 3 : 2;
    const double tri[3][3] = {{q[i0][0], q[i0][1], q[i0][2]}, {q[i1][0], q[i1][1], q[i1][2]}, {q[i2][0], q[i2][1], q[i2][2]}};
    triangle_in_box(tri, H, [&](const double *a, const double *e1, const double *e2) { subtriangle_rule(a, e1, e2, tol, o, emit); });
  }
}

template <class E>
__device__ inline void pair_2d(const double *__restrict__ box, const double *__restrict__ seg, PairAcc &o, E &&emit)
{
  const double lo[2] = {box[0], box[1]};
  const double H[2] = {box[4] - lo[0], box[5] - lo[1]};
  const double p0[2] = {seg[0] - lo[0], seg[1] - lo[1]}, p1[2] = {seg[2] - lo[0], seg[3] - lo[1]};
  double t0, t1;
  if (!segment_in_box(p0, p1, H, t0, t1)) return;
  const double d[2] = {p1[0] - p0[0], p1[1] - p0[1]};
  const double a[2] = {p0[0] + t0 * d[0], p0[1] + t0 * d[1]};
  const double e[2] = {(t1 - t0) * d[0], (t1 - t0) * d[1]};
  const double J = sqrt(e[0] * e[0] + e[1] * e[1]);       // |det B| = length     intersections.cc:702-707
  if (J < 1e-12) { o.dropped++; return; }                  // intersections.cc:710
  const int nq = c_rule.n;
  for (int q = 0; q < nq; ++q) {
    const double s = c_rule.x[q][0];
    const double x[2] = {a[0] + s * e[0], a[1] + s * e[1]};
    emit(x, J * c_rule.w[q]);
  }
  o.nq += nq;
}

// Codimension 0 in 2D, the reference's <2,2,2> instantiation (code/src/intersections.cc:307-355, compute_intersection_quad_quad):
// background rectangle n convex immersed quad.  The reference intersects the polygons with CGAL, triangulates the result
// (CDT), keeps the triangles of area > tol and puts QGaussSimplex<2>(degree) on them (:693-746); here the quad is clipped
// by the four lines of the cell (Sutherland-Hodgman in cell-local coordinates) and the polygon is fanned from its first
// vertex -- the same region, and for a rule that is exact for the bilinear integrand (n_points_1D >= 2) the same integrals.
template <class E>
__device__ inline void pair_2d2d(const double *__restrict__ box, const double *__restrict__ quad /* [4][2], deal.II order */, double tol,
                                 PairAcc &o, E &&emit)
{
  const double lo[2] = {box[0], box[1]};
  const double H[2] = {box[4] - lo[0], box[5] - lo[1]};
  double ax[8], ay[8], bx[8], by[8];
  {
    const int ring[4] = {0, 1, 3, 2};   // deal.II vertex order -> around the cell
#pragma unroll
    for (int i = 0; i < 4; ++i) { ax[i] = quad[2 * ring[i]] - lo[0]; ay[i] = quad[2 * ring[i] + 1] - lo[1]; }
  }
  int n = 4;
  // one closed half-plane after the other: coordinate `axis` >= 0, then <= H[axis]
#pragma unroll 1
  for (int p = 0; p < 4 && n >= 3; ++p) {
    const int axis = p >> 1;
    const double sgn = (p & 1) ? -1.0 : 1.0, bound = (p & 1) ? H[axis] : 0.0;
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const int j = i + 1 == n ? 0 : i + 1;
      const double da = sgn * ((axis ? ay[i] : ax[i]) - bound), db = sgn * ((axis ? ay[j] : ax[j]) - bound);
      if (da >= 0.0) { bx[m] = ax[i]; by[m] = ay[i]; ++m; }
      if ((da > 0.0 && db < 0.0) || (da < 0.0 && db > 0.0)) {
        const double t = da / (da - db);
        double x = ax[i] + t * (ax[j] - ax[i]), y = ay[i] + t * (ay[j] - ay[i]);
        if (axis) y = bound; else x = bound;          // on the line by construction
        bx[m] = x; by[m] = y; ++m;
      }
    }
    n = m;
    for (int i = 0; i < n; ++i) { ax[i] = bx[i]; ay[i] = by[i]; }
  }
  if (n < 3) return;
  const int nq = c_rule.n;
  for (int j = 1; j + 1 < n; ++j) {
    const double e1[2] = {ax[j] - ax[0], ay[j] - ay[0]}, e2[2] = {ax[j + 1] - ax[0], ay[j + 1] - ay[0]};
    const double J = fabs(e1[0] * e2[1] - e1[1] * e2[0]);       // |det B| = 2 * area     intersections.cc:702-707
    if (!(0.5 * J > tol)) continue;                             // triangle area > tol    intersections.cc:336-337
    if (J < 1e-12) { o.dropped++; continue; }                   // intersections.cc:710
    for (int q = 0; q < nq; ++q) {
      const double s = c_rule.x[q][0], t = c_rule.x[q][1];
      const double x[2] = {ax[0] + s * e1[0] + t * e2[0], ay[0] + s * e1[1] + t * e2[1]};
      emit(x, J * c_rule.w[q]);
    }
    o.nq += nq;
  }
}

// bounding box (as the two end points of a diagonal, so that the candidate search treats the quad like a segment with
// the same box), area and convexity of the immersed quads of the codimension-0 case
__global__ void quad2d_prepare_kernel(uint64_t n, const double *__restrict__ poly, double *__restrict__ diag, double *__restrict__ measure,
                                      unsigned long long *n_bad)
{
  const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n) return;
  const double *q = poly + m * 8;
  const int ring[4] = {0, 1, 3, 2};
  double x[4], y[4];
  for (int i = 0; i < 4; ++i) { x[i] = q[2 * ring[i]]; y[i] = q[2 * ring[i] + 1]; }
  double lox = x[0], hix = x[0], loy = y[0], hiy = y[0], a2 = 0.0;
  int pos = 0, neg = 0;
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3, k = (i + 2) & 3;
    lox = x[i] < lox ? x[i] : lox; hix = x[i] > hix ? x[i] : hix;
    loy = y[i] < loy ? y[i] : loy; hiy = y[i] > hiy ? y[i] : hiy;
    a2 += x[i] * y[j] - x[j] * y[i];
    const double c = (x[j] - x[i]) * (y[k] - y[j]) - (y[j] - y[i]) * (x[k] - x[j]);
    pos += c > 0.0; neg += c < 0.0;
  }
  diag[m * 4 + 0] = lox; diag[m * 4 + 1] = loy; diag[m * 4 + 2] = hix; diag[m * 4 + 3] = hiy;
  measure[m] = 0.5 * fabs(a2);
  if ((pos && neg) || !(fabs(a2) > 0.0)) atomicAdd(n_bad, 1ULL);     // turns both ways, or no area: not a convex quad
}

// ---------------------------------------------------------------------------------------
// the fused kernel: one thread per candidate
// ---------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128) clip_local_kernel(uint64_t n_cand, const uint32_t *__restrict__ cand_im,
                                                         const uint32_t *__restrict__ cand_bg, const double *__restrict__ bg_box,
                                                         const double *__restrict__ im_verts, const uint8_t *__restrict__ split,
                                                         const double *__restrict__ im_measure, double tol,
                                                         double *__restrict__ sumw_out, double *__restrict__ local_out,
                                                         uint16_t *__restrict__ nq_out, uint8_t *__restrict__ acc_out,
                                                         unsigned long long *counters, int codim0)
{
  constexpr int NL = 1 << D;
  const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned acc = 0, nq = 0, dropped = 0, marginal = 0;
  if (c < n_cand) {
    const uint32_t m = cand_im[c], b = cand_bg[c];
    const double *box = bg_box + (uint64_t)b * 8;
    PairAcc o;
    o.sumw = 0; o.nq = 0; o.dropped = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) o.loc[i] = 0;
    if (D == 3) {
      const double ih[3] = {1.0 / (box[4] - box[0]), 1.0 / (box[5] - box[1]), 1.0 / (box[6] - box[2])};
      pair_3d(box, im_verts + (uint64_t)m * 12, split[m], tol, o,
              [&](const double *x, double w) { q1_add_3d(o, x[0] * ih[0], x[1] * ih[1], x[2] * ih[2], w); });
    } else {
      const double ih[2] = {1.0 / (box[4] - box[0]), 1.0 / (box[5] - box[1])};
      if (codim0)   // im_verts = the quads [n][4][2]
        pair_2d2d(box, im_verts + (uint64_t)m * 8, tol, o, [&](const double *x, double w) { q1_add_2d(o, x[0] * ih[0], x[1] * ih[1], w); });
      else
        pair_2d(box, im_verts + (uint64_t)m * 4, o, [&](const double *x, double w) { q1_add_2d(o, x[0] * ih[0], x[1] * ih[1], w); });
    }
    const bool ok = o.sumw > tol;                          // coupling_utilities.cc:61
    sumw_out[c] = o.sumw;
    acc_out[c] = ok ? 1 : 0;
    nq_out[c] = (uint16_t)(ok ? o.nq : 0);
    if (ok) {
      double2 *dst = reinterpret_cast<double2 *>(local_out + c * NL);
#pragma unroll
      for (int i = 0; i < NL / 2; ++i) dst[i] = make_double2(o.loc[2 * i], o.loc[2 * i + 1]);
      acc = 1; nq = o.nq;
    }
    dropped = o.dropped;
    marginal = (o.sumw > 0.0 && o.sumw < 1e-12 * im_measure[m]) ? 1 : 0;
  }
  // block-level counters
  __shared__ unsigned s_cnt[4];
  if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  acc = __reduce_add_sync(0xffffffffu, acc);
  nq = __reduce_add_sync(0xffffffffu, nq);
  dropped = __reduce_add_sync(0xffffffffu, dropped);
  marginal = __reduce_add_sync(0xffffffffu, marginal);
  if ((threadIdx.x & 31) == 0) {
    if (acc) atomicAdd(&s_cnt[0], acc);
    if (nq) atomicAdd(&s_cnt[1], nq);
    if (dropped) atomicAdd(&s_cnt[2], dropped);
    if (marginal) atomicAdd(&s_cnt[3], marginal);
  }
  __syncthreads();
  if (threadIdx.x < 4 && s_cnt[threadIdx.x]) atomicAdd(&counters[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

// ---------------------------------------------------------------------------------------
// 3D throughput kernel: warp-level work queue + closed-form moments
// ---------------------------------------------------------------------------------------
// One thread per candidate leaves two thirds of the lanes idle (candidates whose triangles miss
// the cell retire at once, the others clip polygons of different sizes).  Here a warp walks a
// contiguous range of candidates; stage 1 (all lanes: one candidate each) only loads the pair and
// rejects triangles by their own bounding box, the surviving (candidate, triangle) items go to a
// per-warp ring in shared memory, and stage 2 always runs on 32 queued items (one per lane):
// Sutherland-Hodgman clip, then the integrals of 1, u, v, w, uv, uw, vw, uvw over the clipped
// polygon in closed form (u,v,w = reference coordinates of the background cell, affine on the
// triangle):  for a triangle of area A with vertex values f_i, g_i, h_i of affine f, g, h
//     int f     = A/3  * Sf
//     int f g   = A/12 * (Sf Sg + S(fg))
//     int f g h = A/60 * (Sf Sg Sh + Sf S(gh) + Sg S(fh) + Sh S(fg) + 2 S(fgh))
// which is what the reference's degree-5 rule evaluates exactly for the Q1 x P0 integrand
// (degree <= 3 on a planar piece; SURVEY.md rule R1).  The items of one candidate sit next to
// each other in the ring, so they are combined with two shuffles in a fixed order.
// block shape (same 24 warps per SM at 80 registers), measured at sphere cycle 8: 8 warps x 3 blocks 13.27 ms, 4 x 6 13.18,
// 2 x 12 13.12, 4 x 7 (72 registers) 13.66, 16 x 1 (128 registers, no spills) 14.71
#ifndef CM_MINB
#define CM_MINB 12
#endif
#ifndef CM_WARPS_N
#define CM_WARPS_N 2
#endif
constexpr int CM_WARPS = CM_WARPS_N;    // warps per block
constexpr int CM_CHUNKS = 8;   // chunks of 32 candidates per warp
// Vertex pool of the clip in SHARED memory (1) or in local memory (0).  The pool is indexed by values that differ from lane
// to lane; as local memory every such load touches one 32-byte sector per lane (ncu: 1.7 useful bytes per sector, the L1
// pipe the busiest unit of the kernel), in shared memory with the layout [coordinate][slot][thread] the bank depends on
// the lane only, whatever the slot.  Nine slots per thread (item_clip_polygon_v2 reuses the slots of cut-off vertices).
// Measured on B200, sphere cycle 8 (1.06e8 candidates), kernel time:
//   local pool, slots by counter (default)                         13.28 ms
//   local pool, slots of cut-off vertices reused (NM_CLIP_SLOTS=1) 14.83 ms     slots by plane (=2) 14.74 ms
//   shared pool (needs NM_CLIP_SLOTS=1)                            14.69 ms     the compiler runs the fan loop per exit group of
//                                                                               the plane loop (7 of 32 lanes active, +21 % instructions)
//   shared pool + CM_JOIN=1 (explicit __syncwarp before the fan)   13.80 ms     long-scoreboard stalls 3.9 -> 3.4 per issue, but
//                                                                               +11 % instructions; local pool + CM_JOIN=1: 14.23 ms
// The kernel is bound by issue slots and dependent-instruction latency, not by the L1 pipe: the default stays.
#ifndef CM_SHARED_POOL
#define CM_SHARED_POOL 0
#endif
#ifndef CM_JOIN
#define CM_JOIN 0
#endif
#if CM_SHARED_POOL && (NM_CLIP_SLOTS != 1 || NM_CLIP_SECTION || NM_CLIP_BITS != 2)
#error "the nine-slot shared pool needs NM_CLIP_SLOTS=1 and the plane-major box-plane clip"
#endif
constexpr int CM_SLOTS = 9;
constexpr size_t CM_POOL_BYTES = CM_SHARED_POOL ? sizeof(double) * 3 * CM_SLOTS * CM_WARPS * 32 : 0;
struct SPool {
  double *p;   // element (k, i) of this thread at p[(k * CM_SLOTS + i) * blockDim]
  __device__ __forceinline__ double get(int k, int i) const { return p[(k * CM_SLOTS + i) * (CM_WARPS * 32)]; }
  __device__ __forceinline__ void set(int k, int i, double v) { p[(k * CM_SLOTS + i) * (CM_WARPS * 32)] = v; }
};

__global__ void __launch_bounds__(CM_WARPS * 32, CM_MINB) clip_moments3_kernel(
    uint64_t n_cand, const uint32_t *__restrict__ cand_im, const uint32_t *__restrict__ cand_bg, const double *__restrict__ bg_box,
    const double *__restrict__ im_verts, const uint8_t *__restrict__ split, const double *__restrict__ im_measure, double tol,
    unsigned nq_per_tri, double *__restrict__ sumw_out, double *__restrict__ local_out, uint16_t *__restrict__ nq_out,
    uint8_t *__restrict__ acc_out, unsigned long long *counters, const uint32_t *__restrict__ lut_global)
{
  __shared__ uint32_t s_ring[CM_WARPS][128];
#if NM_CLIP_SECTION
  __shared__ uint32_t s_lut[256];   // section polygons: cyclic order of the cut box edges per corner sign pattern
  for (unsigned i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = lut_global[i];
  __syncthreads();
#else
  const uint32_t *s_lut = lut_global;   // not read by the box-plane clip
#endif
  extern __shared__ double s_pool[];
  const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t *ring = s_ring[w];
  const uint64_t wstart = ((uint64_t)blockIdx.x * CM_WARPS + w) * (32 * CM_CHUNKS);
  const uint64_t wend = wstart + 32 * CM_CHUNKS < n_cand ? wstart + 32 * CM_CHUNKS : n_cand;
  unsigned head = 0, count = 0;  // ring state, warp-uniform
  unsigned c_acc = 0, c_nq = 0, c_drop = 0, c_marg = 0;
  for (uint64_t base = wstart; base < wend || count > 0; base += 32) {
    // ---- stage 1: one candidate per lane -> items ------------------------------------------------
    unsigned nitems = 0, items = 0;  // up to three 2-bit triangle ids packed in `items`
    const uint64_t c = base + lane;
    if (base < wend && c < wend) {
      const uint32_t m = cand_im[c], b = cand_bg[c];
      const double *box = bg_box + (uint64_t)b * 8;
      const double *quad = im_verts + (uint64_t)m * 12;
      const int code = split[m];
      const double H[3] = {box[4] - box[0], box[5] - box[1], box[6] - box[2]};
      // k-th triangle of the split code: every lane runs the same (usually two) iterations
      unsigned rem = (unsigned)code & 15u;
#pragma unroll 1
      while (rem) {
        const int t = __ffs(rem) - 1;
        rem &= rem - 1;
        const int iv[3] = {t == 0 ? 1 : 0, t <= 1 ? 2 : 1, t <= 2 ? 3 : 2};
        double tri[3][3];
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
          for (int k = 0; k < 3; ++k) tri[v][k] = quad[iv[v] * 3 + k] - box[k];
        if (tri_gate(tri, H)) { items |= (unsigned)t << (2 * nitems); ++nitems; }
      }
      if (nitems == 0) { sumw_out[c] = 0.0; acc_out[c] = 0; nq_out[c] = 0; }
    }
    // push (warp-aggregated, candidate order preserved)
    unsigned inc = nitems;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (unsigned)o) inc += v; }
    const unsigned total = __shfl_sync(0xffffffffu, inc, 31);
    unsigned pos = head + count + inc - nitems;
    for (unsigned k = 0; k < nitems; ++k) ring[(pos + k) & 127] = ((uint32_t)(c - wstart) << 2) | ((items >> (2 * k)) & 3);
    count += total;
    __syncwarp();
    const bool last = base + 32 >= wend;
    // ---- stage 2: 32 queued items at a time ------------------------------------------------------
    while (count >= 32 || (last && count > 0)) {
      unsigned take = count < 32 ? count : 32;
      if (take < count) {  // do not split the items of one candidate between two rounds
        const uint32_t nxt = ring[(head + take) & 127] >> 2;
        while (take > 0 && (ring[(head + take - 1) & 127] >> 2) == nxt) --take;
      }
      const bool on = lane < take;
      const uint32_t it = on ? ring[(head + lane) & 127] : 0xffffffffu;
      const uint32_t cl = it >> 2;
      ItemOut o;
      o.area2 = 0; o.nfan = 0; o.dropped = 0;
#pragma unroll
      for (int i = 0; i < 7; ++i) o.m[i] = 0;
      uint32_t m = 0;
      const uint64_t cc = wstart + cl;
#if CM_JOIN
      // clip first, then -- with the warp converged again (lanes leave the plane loop at different trips and by two
      // exits) -- integrate: without the explicit join the compiler may run the fan loop once per exit group
#if CM_SHARED_POOL
      ClippedPolyT<SPool> C;
      C.P.p = s_pool + threadIdx.x;
#else
      ClippedPoly C;
#endif
      bool ok = false;
      const double *box2 = bg_box;
      if (on) {
        m = cand_im[cc];
        box2 = bg_box + (uint64_t)cand_bg[cc] * 8;
#if CM_SHARED_POOL
        ok = item_clip_polygon_v2(box2, im_verts + (uint64_t)m * 12, (int)(it & 3), C);
#else
        ok = item_polygon(box2, im_verts + (uint64_t)m * 12, (int)(it & 3), s_lut, C);
#endif
      }
      __syncwarp();
      if (ok) poly_moments(C, box2, tol, o);
      __syncwarp();
#else
      if (on) {
        m = cand_im[cc];
#if CM_SHARED_POOL
        ClippedPolyT<SPool> C;
        C.P.p = s_pool + threadIdx.x;
        item_moments_pool(bg_box + (uint64_t)cand_bg[cc] * 8, im_verts + (uint64_t)m * 12, (int)(it & 3), tol, C, o);
#else
        item_moments(bg_box + (uint64_t)cand_bg[cc] * 8, im_verts + (uint64_t)m * 12, (int)(it & 3), tol, s_lut, o);
#endif
      }
#endif
      // combine the (adjacent, at most three) items of a candidate in ring order
      const uint32_t cl1 = __shfl_down_sync(0xffffffffu, cl, 1), cl2 = __shfl_down_sync(0xffffffffu, cl, 2);
      const uint32_t clp = __shfl_up_sync(0xffffffffu, cl, 1);
      const bool same1 = on && lane + 1 < 32 && cl1 == cl, same2 = on && lane + 2 < 32 && cl2 == cl;
      const bool is_head = on && (lane == 0 || clp != cl);
      double tot[8];
      tot[0] = o.area2;
#pragma unroll
      for (int i = 0; i < 7; ++i) tot[i + 1] = o.m[i];
      unsigned nfan = o.nfan;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const double a1 = __shfl_down_sync(0xffffffffu, tot[i], 1), a2 = __shfl_down_sync(0xffffffffu, tot[i], 2);
        tot[i] = tot[i] + (same1 ? a1 : 0.0) + (same2 ? a2 : 0.0);
      }
      {
        const unsigned n1 = __shfl_down_sync(0xffffffffu, nfan, 1), n2 = __shfl_down_sync(0xffffffffu, nfan, 2);
        nfan += (same1 ? n1 : 0u) + (same2 ? n2 : 0u);
      }
      c_drop += o.dropped;
      if (is_head) {
        const double sumw = 0.5 * tot[0];
        const bool ok = sumw > tol;                        // coupling_utilities.cc:61
        sumw_out[cc] = sumw;
        acc_out[cc] = ok ? 1 : 0;
        const unsigned nq = ok ? nfan * nq_per_tri : 0u;
        nq_out[cc] = (uint16_t)nq;
        if (ok) {
          double l8[8], sw_;
          moments_to_local(tot, sw_, l8);
          const double l0 = l8[0], l1 = l8[1], l2 = l8[2], l3 = l8[3], l4 = l8[4], l5 = l8[5], l6 = l8[6], l7 = l8[7];
          double2 *dst = reinterpret_cast<double2 *>(local_out + cc * 8);
          dst[0] = make_double2(l0, l1); dst[1] = make_double2(l2, l3); dst[2] = make_double2(l4, l5); dst[3] = make_double2(l6, l7);
          c_acc++; c_nq += nq;
        }
        if (sumw > 0.0 && sumw < 1e-12 * im_measure[m]) c_marg++;
      }
      head = (head + take) & 127;
      count -= take;
      __syncwarp();
    }
    if (base >= wend) break;
  }
  c_acc = __reduce_add_sync(0xffffffffu, c_acc);
  c_nq = __reduce_add_sync(0xffffffffu, c_nq);
  c_drop = __reduce_add_sync(0xffffffffu, c_drop);
  c_marg = __reduce_add_sync(0xffffffffu, c_marg);
  if (lane == 0) {
    if (c_acc) atomicAdd(&counters[0], (unsigned long long)c_acc);
    if (c_nq) atomicAdd(&counters[1], (unsigned long long)c_nq);
    if (c_drop) atomicAdd(&counters[2], (unsigned long long)c_drop);
    if (c_marg) atomicAdd(&counters[3], (unsigned long long)c_marg);
  }
}

// ---------------------------------------------------------------------------------------
// unfused variants for the drop-in wrapper
// ---------------------------------------------------------------------------------------
// points/weights of every accepted pair (what the reference returns as Quadrature<spacedim>).
// MOMENT_PATH: the pair was accepted by clip_moments3_kernel -- walk exactly its polygons (same SAT
// gate, same clip, same fan, same thresholds) so that the number of points equals the count it stored.
template <int D, bool MOMENT_PATH>
__global__ void emit_quadrature_kernel(uint64_t n_acc, const uint64_t *__restrict__ acc_idx, const uint32_t *__restrict__ cand_im,
                                       const uint32_t *__restrict__ cand_bg, const double *__restrict__ bg_box,
                                       const double *__restrict__ im_verts, const uint8_t *__restrict__ split, double tol,
                                       const uint64_t *__restrict__ q_off, double *__restrict__ pts, double *__restrict__ wts,
                                       unsigned long long *n_mismatch, const uint32_t *__restrict__ lut, int codim0)
{
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_acc) return;
  const uint64_t c = acc_idx[p];
  const uint32_t m = cand_im[c], b = cand_bg[c];
  const double *box = bg_box + (uint64_t)b * 8;
  PairAcc o;
  o.sumw = 0; o.nq = 0; o.dropped = 0;
  uint64_t w = q_off[p];
  const uint64_t w_end = q_off[p + 1];
  if (D == 3 && MOMENT_PATH) {
    const double *quad = im_verts + (uint64_t)m * 12;
    const double H[3] = {box[4] - box[0], box[5] - box[1], box[6] - box[2]};
    unsigned rem = (unsigned)split[m] & 15u;
    while (rem) {
      const int t = __ffs(rem) - 1;
      rem &= rem - 1;
      const int iv[3] = {t == 0 ? 1 : 0, t <= 1 ? 2 : 1, t <= 2 ? 3 : 2};
      double tri[3][3];
      for (int v = 0; v < 3; ++v)
        for (int k = 0; k < 3; ++k) tri[v][k] = quad[iv[v] * 3 + k] - box[k];
      if (!tri_gate(tri, H)) continue;
      ClippedPoly C;
      if (!item_polygon(box, quad, t, lut, C)) continue;
      const int v0 = (int)(C.poly & 15u);
      const double a[3] = {C.P.c[0][v0], C.P.c[1][v0], C.P.c[2][v0]};
      for (int j = 2; j < C.n; ++j) {
        const int v1 = (int)((C.poly >> (4 * (j - 1))) & 15u), v2 = (int)((C.poly >> (4 * j)) & 15u);
        const double e1[3] = {C.P.c[0][v1] - a[0], C.P.c[1][v1] - a[1], C.P.c[2][v1] - a[2]};
        const double e2[3] = {C.P.c[0][v2] - a[0], C.P.c[1][v2] - a[1], C.P.c[2][v2] - a[2]};
        const double J = fabs((e1[1] * e2[2] - e1[2] * e2[1]) * C.nx + (e1[2] * e2[0] - e1[0] * e2[2]) * C.ny + (e1[0] * e2[1] - e1[1] * e2[0]) * C.nz);
        if (!(0.25 * J * J > tol * tol) || J < 1e-12) continue;       // same thresholds as item_clip_moments
        for (int q = 0; q < c_rule.n; ++q) {
          const double s = c_rule.x[q][0], u = c_rule.x[q][1];
          if (w < w_end) {
            pts[w * 3 + 0] = a[0] + s * e1[0] + u * e2[0] + box[0];
            pts[w * 3 + 1] = a[1] + s * e1[1] + u * e2[1] + box[1];
            pts[w * 3 + 2] = a[2] + s * e1[2] + u * e2[2] + box[2];
            wts[w] = J * c_rule.w[q];
          }
          ++w;
        }
      }
    }
  } else if (D == 3) {
    pair_3d(box, im_verts + (uint64_t)m * 12, split[m], tol, o, [&](const double *x, double wq) {
      if (w < w_end) { pts[w * 3 + 0] = x[0] + box[0]; pts[w * 3 + 1] = x[1] + box[1]; pts[w * 3 + 2] = x[2] + box[2]; wts[w] = wq; }
      ++w;
    });
  } else if (codim0) {
    pair_2d2d(box, im_verts + (uint64_t)m * 8, tol, o, [&](const double *x, double wq) {
      if (w < w_end) { pts[w * 2 + 0] = x[0] + box[0]; pts[w * 2 + 1] = x[1] + box[1]; wts[w] = wq; }
      ++w;
    });
  } else {
    pair_2d(box, im_verts + (uint64_t)m * 4, o, [&](const double *x, double wq) {
      if (w < w_end) { pts[w * 2 + 0] = x[0] + box[0]; pts[w * 2 + 1] = x[1] + box[1]; wts[w] = wq; }
      ++w;
    });
  }
  if (w != w_end) atomicAdd(n_mismatch, 1ULL);   // the emitted count must equal the count stored with the pair
}

__global__ void gather_nq_kernel(uint64_t n_acc, const uint64_t *__restrict__ acc_idx, const uint16_t *__restrict__ nq, uint32_t *__restrict__ out)
{
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_acc) out[p] = nq[acc_idx[p]];
  if (p == n_acc) out[p] = 0;
}

// local coupling vector from caller-supplied quadrature (coupling_utilities.h:238-274):
// xi_q = F^-1(x_q) on the axis-aligned cell, local(i) += phi_i(xi_q) w_q.  A pair that appears in several tuples adds
// up, as the reference's per-tuple distribute_local_to_global does (:285-287); `claim` (one bit per candidate) tells
// the first tuple of a pair from the later ones, so that the pair is counted once.
template <int D>
__global__ void local_from_quadrature_kernel(uint64_t n_pairs, const uint32_t *__restrict__ im, const uint32_t *__restrict__ bg,
                                             const uint64_t *__restrict__ q_off, const double *__restrict__ pts,
                                             const double *__restrict__ wts, const uint64_t *__restrict__ cand_ptr,
                                             const uint32_t *__restrict__ cand_bg, const double *__restrict__ bg_box, uint64_t n_im,
                                             double *__restrict__ sumw_out, double *__restrict__ local_out,
                                             uint8_t *__restrict__ acc_out, uint32_t *__restrict__ claim, unsigned long long *n_missing,
                                             unsigned long long *n_distinct)
{
  constexpr int NL = 1 << D;
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const uint32_t m = im[p], b = bg[p];
  if (m >= n_im) { atomicAdd(n_missing, 1ULL); return; }
  // locate the pair in the candidate list (sorted by background id inside the cell's run)
  uint64_t lo = cand_ptr[m], hi = cand_ptr[m + 1];
  while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (cand_bg[mid] < b) lo = mid + 1; else hi = mid; }
  if (lo >= cand_ptr[m + 1] || cand_bg[lo] != b) { atomicAdd(n_missing, 1ULL); return; }
  const double *box = bg_box + (uint64_t)b * 8;
  PairAcc o;
  o.sumw = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) o.loc[i] = 0;
  double ih[3];
#pragma unroll
  for (int k = 0; k < D; ++k) ih[k] = 1.0 / (box[4 + k] - box[k]);
  for (uint64_t q = q_off[p]; q < q_off[p + 1]; ++q) {
    const double *x = pts + q * D;
    if (D == 3) q1_add_3d(o, (x[0] - box[0]) * ih[0], (x[1] - box[1]) * ih[1], (x[2] - box[2]) * ih[2], wts[q]);
    else q1_add_2d(o, (x[0] - box[0]) * ih[0], (x[1] - box[1]) * ih[1], wts[q]);
  }
  const unsigned bit = 1u << (lo & 31);
  if (!(atomicOr(&claim[lo >> 5], bit) & bit)) { acc_out[lo] = 1; atomicAdd(n_distinct, 1ULL); }
  // the buffers were cleared: a pair given once receives exactly its own sums
  atomicAdd(&sumw_out[lo], o.sumw);
#pragma unroll
  for (int i = 0; i < NL; ++i) atomicAdd(&local_out[lo * NL + i], o.loc[i]);
}

__global__ void flag_to_u64(uint64_t n, const uint8_t *__restrict__ f, uint64_t *__restrict__ out)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = f[i];
  if (i == n) out[i] = 0;
}
__global__ void compact_idx(uint64_t n, const uint8_t *__restrict__ f, const uint64_t *__restrict__ pos, uint64_t *__restrict__ idx)
{
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && f[i]) idx[pos[i]] = i;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
int nm_split_run(nm_ctx *ctx)
{
  const uint64_t n = ctx->n_im;
  NM_TRY(nm_ensure(ctx, ctx->im_split, n + 1));
  NM_TRY(nm_ensure(ctx, ctx->im_measure, (n + 1) * sizeof(double)));
  NM_TRY(nm_ensure(ctx, ctx->counters, 64 * sizeof(uint64_t)));
  unsigned long long *ties = ctx->counters.as<unsigned long long>() + 8;
  NM_CUDA(cudaMemsetAsync(ties, 0, 8, ctx->stream));
  ctx->counts.n_split_ties = 0;
  if (n == 0) return NM_OK;
  if (ctx->spacedim == 3) {
    split_kernel<<<nm_blocks(n, 128), 128, 0, NM_LS(ctx)>>>(n, ctx->im_verts.as<double>(), ctx->im_split.as<uint8_t>(),
                                                              ctx->im_measure.as<double>(), ties);
    NM_CUDA(cudaGetLastError());
    unsigned long long h = 0;
    NM_CUDA(cudaMemcpyAsync(&h, ties, 8, cudaMemcpyDeviceToHost, ctx->stream));
    NM_SYNC(ctx);
    ctx->counts.n_split_ties = h;
  } else if (!ctx->codim0) {   // (codimension 0: the areas were taken when the quads were set, nm_quad2d_prepare)
    segment_measure_kernel<<<nm_blocks(n, 256), 256, 0, NM_LS(ctx)>>>(n, ctx->im_verts.as<double>(), ctx->im_measure.as<double>());
    NM_CUDA(cudaGetLastError());
  }
  return NM_OK;
}

// codimension 0 in 2D: from the quads in ctx->im_poly, the diagonal of every bounding box (ctx->im_verts: what the candidate
// search reads), the areas, and a convexity check
int nm_quad2d_prepare(nm_ctx *ctx)
{
  const uint64_t n = ctx->n_im;
  NM_TRY(nm_ensure(ctx, ctx->im_verts, (n + 1) * 4 * sizeof(double)));
  NM_TRY(nm_ensure(ctx, ctx->im_measure, (n + 1) * sizeof(double)));
  NM_TRY(nm_ensure(ctx, ctx->counters, 64 * sizeof(uint64_t)));
  if (n == 0) return NM_OK;
  unsigned long long *bad = ctx->counters.as<unsigned long long>() + 9;
  NM_CUDA(cudaMemsetAsync(bad, 0, 8, ctx->stream));
  quad2d_prepare_kernel<<<nm_blocks(n, 256), 256, 0, NM_LS(ctx)>>>(n, ctx->im_poly.as<double>(), ctx->im_verts.as<double>(),
                                                                    ctx->im_measure.as<double>(), bad);
  NM_CUDA(cudaGetLastError());
  unsigned long long h = 0;
  NM_CUDA(cudaMemcpyAsync(&h, bad, 8, cudaMemcpyDeviceToHost, ctx->stream));
  NM_SYNC(ctx);
  if (h) return nm_fail(ctx, NM_ERR_UNSUPPORTED, "%llu immersed quad(s) are not convex (or have no area): the codimension-0 path clips convex cells only", h);
  return NM_OK;
}

static int upload_rule(nm_ctx *ctx)
{
  if (!ctx->clip_lut.p) {   // section polygons: cut box edges in cyclic order, per corner sign pattern (built once per context)
    uint32_t lut[256];
    section_lut_build(lut);
    NM_TRY(nm_ensure(ctx, ctx->clip_lut, sizeof(lut)));
    NM_CUDA(cudaMemcpyAsync(ctx->clip_lut.p, lut, sizeof(lut), cudaMemcpyHostToDevice, ctx->stream));
    NM_SYNC(ctx);   // `lut` lives on this stack frame
  }
  Rule r;
  if (!make_rule(ctx->dim1, ctx->degree, r))
    return nm_fail(ctx, NM_ERR_UNSUPPORTED, "QGaussSimplex<%d>(n_points_1D=%u) is not tabulated (supported: %s)", ctx->dim1,
                   ctx->degree, ctx->dim1 == 1 ? "1..8" : "1..3");
  NM_CUDA(cudaMemcpyToSymbolAsync(c_rule, &r, sizeof(Rule), 0, cudaMemcpyHostToDevice, ctx->stream));
  return NM_OK;
}

int nm_clip_run(nm_ctx *ctx)
{
  const int d = ctx->spacedim;
  const uint64_t nc = ctx->counts.n_candidates;
  const int NL = 1 << d;
  NM_TRY(upload_rule(ctx));
  NM_TRY(nm_ensure(ctx, ctx->cand_sumw, (nc + 1) * sizeof(double)));
  NM_TRY(nm_ensure(ctx, ctx->cand_local, (nc + 1) * NL * sizeof(double)));
  NM_TRY(nm_ensure(ctx, ctx->cand_nq, (nc + 1) * sizeof(uint16_t)));
  NM_TRY(nm_ensure(ctx, ctx->scratch_b, (nc + 2) * sizeof(uint64_t)));  // reused later for compaction
  NM_TRY(nm_ensure(ctx, ctx->cand_acc, nc + 1));
  unsigned long long *cnt = ctx->counters.as<unsigned long long>();
  NM_CUDA(cudaMemsetAsync(cnt, 0, 4 * sizeof(unsigned long long), ctx->stream));
  ctx->counts.n_accepted = ctx->counts.n_qpoints = ctx->counts.n_dropped = ctx->counts.n_marginal = 0;
  if (nc == 0) return NM_OK;
  const unsigned T = 128, B = nm_blocks(nc, T);
  KernelTimer kt(ctx, NM_T_K_CLIP);
  // closed-form moments are what a rule exact to degree 3 returns: the reference's degree = 3
  // (QGaussSimplex<2>(3), exact to degree 5) qualifies; lower rules go through the quadrature kernel
  if (d == 3 && ctx->degree == 3 && !ctx->force_quadrature_kernel) {
    static bool attr_set = false;   // the pool needs more than the 48 KB a kernel gets by default
    if (CM_POOL_BYTES && !attr_set) {
      NM_CUDA(cudaFuncSetAttribute(clip_moments3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CM_POOL_BYTES));
      attr_set = true;
    }
    clip_moments3_kernel<<<nm_blocks(nc, CM_WARPS * 32 * CM_CHUNKS), CM_WARPS * 32, CM_POOL_BYTES, NM_LS(ctx)>>>(
        nc, ctx->cand_im.as<uint32_t>(), ctx->cand_bg.as<uint32_t>(), ctx->bg_box.as<double>(), ctx->im_verts.as<double>(),
        ctx->im_split.as<uint8_t>(), ctx->im_measure.as<double>(), ctx->tol, 7u, ctx->cand_sumw.as<double>(),
        ctx->cand_local.as<double>(), ctx->cand_nq.as<uint16_t>(), ctx->cand_acc.as<uint8_t>(), cnt, ctx->clip_lut.as<uint32_t>());
  } else if (d == 3)
    clip_local_kernel<3><<<B, T, 0, NM_LS(ctx)>>>(nc, ctx->cand_im.as<uint32_t>(), ctx->cand_bg.as<uint32_t>(), ctx->bg_box.as<double>(),
                                                   ctx->im_verts.as<double>(), ctx->im_split.as<uint8_t>(), ctx->im_measure.as<double>(),
                                                   ctx->tol, ctx->cand_sumw.as<double>(), ctx->cand_local.as<double>(),
                                                   ctx->cand_nq.as<uint16_t>(), ctx->cand_acc.as<uint8_t>(), cnt, 0);
  else
    clip_local_kernel<2><<<B, T, 0, NM_LS(ctx)>>>(nc, ctx->cand_im.as<uint32_t>(), ctx->cand_bg.as<uint32_t>(), ctx->bg_box.as<double>(),
                                                   ctx->codim0 ? ctx->im_poly.as<double>() : ctx->im_verts.as<double>(), nullptr,
                                                   ctx->im_measure.as<double>(), ctx->tol, ctx->cand_sumw.as<double>(), ctx->cand_local.as<double>(),
                                                   ctx->cand_nq.as<uint16_t>(), ctx->cand_acc.as<uint8_t>(), cnt, ctx->codim0 ? 1 : 0);
  kt.stop();
  NM_CUDA(cudaGetLastError());
  unsigned long long h[4];
  NM_CUDA(cudaMemcpyAsync(h, cnt, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  NM_SYNC(ctx);
  ctx->counts.n_accepted = h[0];
  ctx->counts.n_qpoints = h[1];
  ctx->counts.n_dropped = h[2];
  ctx->counts.n_marginal = h[3];
  return NM_OK;
}

// accepted-candidate index list (ascending) in ctx->scratch_b[nc+2 ...]; returns device pointer
int nm_accepted_index(nm_ctx *ctx, uint64_t **idx_dev)
{
  const uint64_t nc = ctx->counts.n_candidates, na = ctx->counts.n_accepted;
  NM_TRY(nm_ensure(ctx, ctx->scratch_b, (nc + 2) * sizeof(uint64_t)));
  NM_TRY(nm_ensure(ctx, ctx->stage_x, (nc + 2) * sizeof(uint64_t)));
  NM_TRY(nm_ensure(ctx, ctx->a_cursor, (na + 2) * sizeof(uint64_t)));
  uint64_t *flags = ctx->scratch_b.as<uint64_t>(), *pos = ctx->stage_x.as<uint64_t>();
  flag_to_u64<<<nm_blocks(nc + 1, 256), 256, 0, NM_LS(ctx)>>>(nc, ctx->cand_acc.as<uint8_t>(), flags);
  NM_TRY(nm_exclusive_scan_u64(ctx, flags, pos, nc + 1));
  if (nc) compact_idx<<<nm_blocks(nc, 256), 256, 0, NM_LS(ctx)>>>(nc, ctx->cand_acc.as<uint8_t>(), pos, ctx->a_cursor.as<uint64_t>());
  NM_CUDA(cudaGetLastError());
  *idx_dev = ctx->a_cursor.as<uint64_t>();
  return NM_OK;
}

int nm_quadrature_emit(nm_ctx *ctx, uint64_t *q_offset, double *points, double *weights, int where)
{
  const int d = ctx->spacedim;
  const uint64_t na = ctx->counts.n_accepted, nq = ctx->counts.n_qpoints;
  uint64_t *acc_idx = nullptr;
  NM_TRY(nm_accepted_index(ctx, &acc_idx));
  // offsets
  DevBuf &cnt32 = ctx->scratch_a;
  NM_TRY(nm_ensure(ctx, cnt32, (na + 2) * sizeof(uint32_t)));
  DevBuf &offs = ctx->c_rowstart;  // scratch use is safe: matrix build re-creates it
  ctx->matrix_valid = false;
  NM_TRY(nm_ensure(ctx, offs, (na + 2) * sizeof(uint64_t)));
  gather_nq_kernel<<<nm_blocks(na + 1, 256), 256, 0, NM_LS(ctx)>>>(na, acc_idx, ctx->cand_nq.as<uint16_t>(), cnt32.as<uint32_t>());
  NM_TRY(nm_exclusive_scan_u32_to_u64(ctx, cnt32.as<uint32_t>(), offs.as<uint64_t>(), na + 1));
  if (q_offset) NM_TRY(nm_download(ctx, q_offset, offs.p, (na + 1) * sizeof(uint64_t), where));
  if (!points && !weights) return NM_OK;
  double *pts_d = points, *w_d = weights;
  DevBuf &pbuf = ctx->c_val, &wbuf = ctx->c_col;  // scratch use (matrix invalidated above)
  if (where == NM_HOST || !points || !weights) {
    NM_TRY(nm_ensure(ctx, pbuf, (nq + 1) * d * sizeof(double)));
    NM_TRY(nm_ensure(ctx, wbuf, (nq + 1) * sizeof(double)));
    if (where == NM_HOST || !points) pts_d = pbuf.as<double>();
    if (where == NM_HOST || !weights) w_d = wbuf.as<double>();
  }
  unsigned long long *mismatch = ctx->counters.as<unsigned long long>() + 10;
  NM_CUDA(cudaMemsetAsync(mismatch, 0, 8, ctx->stream));
  if (na) {
    NM_TRY(upload_rule(ctx));
    const bool moment_path = d == 3 && ctx->degree == 3 && !ctx->force_quadrature_kernel;
    const unsigned B = nm_blocks(na, 128);
    const uint32_t *ci = ctx->cand_im.as<uint32_t>(), *cb = ctx->cand_bg.as<uint32_t>();
    if (d == 3 && moment_path)
      emit_quadrature_kernel<3, true><<<B, 128, 0, NM_LS(ctx)>>>(na, acc_idx, ci, cb, ctx->bg_box.as<double>(), ctx->im_verts.as<double>(),
                                                                ctx->im_split.as<uint8_t>(), ctx->tol, offs.as<uint64_t>(), pts_d, w_d, mismatch, ctx->clip_lut.as<uint32_t>(), 0);
    else if (d == 3)
      emit_quadrature_kernel<3, false><<<B, 128, 0, NM_LS(ctx)>>>(na, acc_idx, ci, cb, ctx->bg_box.as<double>(), ctx->im_verts.as<double>(),
                                                                 ctx->im_split.as<uint8_t>(), ctx->tol, offs.as<uint64_t>(), pts_d, w_d, mismatch, ctx->clip_lut.as<uint32_t>(), 0);
    else
      emit_quadrature_kernel<2, false><<<B, 128, 0, NM_LS(ctx)>>>(na, acc_idx, ci, cb, ctx->bg_box.as<double>(),
                                                                 ctx->codim0 ? ctx->im_poly.as<double>() : ctx->im_verts.as<double>(), nullptr,
                                                                 ctx->tol, offs.as<uint64_t>(), pts_d, w_d, mismatch, ctx->clip_lut.as<uint32_t>(),
                                                                 ctx->codim0 ? 1 : 0);
    NM_CUDA(cudaGetLastError());
    unsigned long long h_mis = 0;
    NM_CUDA(cudaMemcpyAsync(&h_mis, mismatch, 8, cudaMemcpyDeviceToHost, ctx->stream));
    NM_SYNC(ctx);
    if (h_mis) return nm_fail(ctx, NM_ERR_STATE, "internal error: %llu pairs emitted a different number of quadrature points than counted", h_mis);
  }
  if (where == NM_HOST) {
    if (points) NM_TRY(nm_download(ctx, points, pts_d, nq * d * sizeof(double), NM_HOST));
    if (weights) NM_TRY(nm_download(ctx, weights, w_d, nq * sizeof(double), NM_HOST));
  }
  NM_SYNC(ctx);
  return NM_OK;
}

int nm_local_from_quadrature(nm_ctx *ctx, uint64_t n_pairs, const uint32_t *im, const uint32_t *bg, const uint64_t *q_offset,
                             const double *points, const double *weights, int where)
{
  const int d = ctx->spacedim, NL = 1 << d;
  const uint64_t nc = ctx->counts.n_candidates;
  uint64_t nq = 0;
  // all inputs to the device
  const uint32_t *im_d = im, *bg_d = bg;
  const uint64_t *qo_d = q_offset;
  const double *p_d = points, *w_d = weights;
  if (where == NM_HOST) {
    nq = n_pairs ? q_offset[n_pairs] : 0;
    size_t b0 = n_pairs * 4, b1 = n_pairs * 4, b2 = (n_pairs + 1) * 8, b3 = nq * d * 8, b4 = nq * 8;
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    NM_TRY(nm_ensure(ctx, ctx->h2d_stage, al(b0) + al(b1) + al(b2) + al(b3) + al(b4) + 256));
    char *base = ctx->h2d_stage.as<char>();
    char *p0 = base, *p1 = p0 + al(b0), *p2 = p1 + al(b1), *p3 = p2 + al(b2), *p4 = p3 + al(b3);
    NM_CUDA(cudaMemcpyAsync(p0, im, b0, cudaMemcpyHostToDevice, ctx->stream));
    NM_CUDA(cudaMemcpyAsync(p1, bg, b1, cudaMemcpyHostToDevice, ctx->stream));
    NM_CUDA(cudaMemcpyAsync(p2, q_offset, b2, cudaMemcpyHostToDevice, ctx->stream));
    NM_CUDA(cudaMemcpyAsync(p3, points, b3, cudaMemcpyHostToDevice, ctx->stream));
    NM_CUDA(cudaMemcpyAsync(p4, weights, b4, cudaMemcpyHostToDevice, ctx->stream));
    im_d = (const uint32_t *)p0; bg_d = (const uint32_t *)p1; qo_d = (const uint64_t *)p2; p_d = (const double *)p3; w_d = (const double *)p4;
  }
  NM_TRY(nm_ensure(ctx, ctx->cand_sumw, (nc + 1) * sizeof(double)));
  NM_TRY(nm_ensure(ctx, ctx->cand_local, (nc + 1) * NL * sizeof(double)));
  NM_TRY(nm_ensure(ctx, ctx->cand_acc, nc + 1));
  NM_TRY(nm_ensure(ctx, ctx->scratch_b, (nc / 32 + 2) * sizeof(uint32_t)));   // one claim bit per candidate
  NM_CUDA(cudaMemsetAsync(ctx->cand_acc.p, 0, nc + 1, ctx->stream));
  NM_CUDA(cudaMemsetAsync(ctx->cand_sumw.p, 0, (nc + 1) * sizeof(double), ctx->stream));
  NM_CUDA(cudaMemsetAsync(ctx->cand_local.p, 0, (nc + 1) * NL * sizeof(double), ctx->stream));
  NM_CUDA(cudaMemsetAsync(ctx->scratch_b.p, 0, (nc / 32 + 2) * sizeof(uint32_t), ctx->stream));
  unsigned long long *missing = ctx->counters.as<unsigned long long>() + 9;    // [9] pairs without overlapping boxes, [10] distinct pairs
  NM_CUDA(cudaMemsetAsync(missing, 0, 16, ctx->stream));
  if (n_pairs) {
    if (d == 3)
      local_from_quadrature_kernel<3><<<nm_blocks(n_pairs, 128), 128, 0, NM_LS(ctx)>>>(
          n_pairs, im_d, bg_d, qo_d, p_d, w_d, ctx->cand_ptr.as<uint64_t>(), ctx->cand_bg.as<uint32_t>(), ctx->bg_box.as<double>(), ctx->n_im,
          ctx->cand_sumw.as<double>(), ctx->cand_local.as<double>(), ctx->cand_acc.as<uint8_t>(), ctx->scratch_b.as<uint32_t>(), missing, missing + 1);
    else
      local_from_quadrature_kernel<2><<<nm_blocks(n_pairs, 128), 128, 0, NM_LS(ctx)>>>(
          n_pairs, im_d, bg_d, qo_d, p_d, w_d, ctx->cand_ptr.as<uint64_t>(), ctx->cand_bg.as<uint32_t>(), ctx->bg_box.as<double>(), ctx->n_im,
          ctx->cand_sumw.as<double>(), ctx->cand_local.as<double>(), ctx->cand_acc.as<uint8_t>(), ctx->scratch_b.as<uint32_t>(), missing, missing + 1);
    NM_CUDA(cudaGetLastError());
  }
  unsigned long long h[2] = {0, 0};
  NM_CUDA(cudaMemcpyAsync(h, missing, 16, cudaMemcpyDeviceToHost, ctx->stream));
  NM_SYNC(ctx);
  if (h[0]) return nm_fail(ctx, NM_ERR_INVALID, "%llu of the given pairs do not have overlapping bounding boxes on the grids set in this context", h[0]);
  ctx->counts.n_accepted = h[1];   // distinct pairs
  return NM_OK;
}
