/* TEST INFRASTRUCTURE ONLY -- plain-C FP64 restatement of the reference's coupling path.
 *
 * PARITY UNPINNED: deal.II 9.4.2 / CGAL 5.5.2 / Boost / Trilinos are not available in this
 * environment and the reference ships no tests, golden vectors or stored tables for this
 * path.  This file restates the reference algorithm from its sources;
 * it is pinned by oracle/exact_oracle.py (exact rational arithmetic) and by the
 * known-answer identities in tests/.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product (libnmgpu.so) never does.
 *
 * It deliberately follows the *reference's* decomposition, not the CUDA path's:
 *   - candidate pairs = closed bounding-box overlap        src/coupling_utilities.cc:40-55
 *   - 2D: background quad -> two triangles, segment ∩ triangle, keep 1-D results
 *                                                           src/intersections.cc:358-393
 *   - 3D: hexahedron -> tetrahedra, quad -> Delaunay triangles of its xy-projection,
 *         triangle ∩ tetrahedron, polygon results re-triangulated, squared-area filter
 *                                                           src/intersections.cc:437-509
 *   - simplex -> QGaussSimplex rule, |J| < 1e-12 dropped    src/intersections.cc:693-746
 *   - accept iff sum of weights > tol                       src/coupling_utilities.cc:59-65
 *   - sparsity, keep_constrained_entries                    include/coupling_utilities.h:100-139
 *   - local(i,j) += phi_i psi_j w, scatter through lines    include/coupling_utilities.h:238-289
 *   - y = A x, y = A^T x on the CSR matrix (Ct.vmult / C.Tvmult)
 *                                   examples/lagrange_multiplier/2d3d/lagrange_multiplier.cc:627-643
 * CGAL's Delaunay tetrahedralisation of the 8 cospherical hex vertices is not reproducible
 * without CGAL; the six Kuhn tetrahedra around the main diagonal are used (survey rule R1:
 * any subdivision of the box gives the same integrals up to rounding).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NMO_TIE 0x80

/* ------------------------------------------------------------------------- */
/* exact sign of sums of products of doubles (Shewchuk-style expansions)      */
/* ------------------------------------------------------------------------- */
static inline void two_sum(double a, double b, double *s, double *e)
{
  double x = a + b, bv = x - a, av = x - bv;
  *s = x;
  *e = (a - av) + (b - bv);
}
static inline void two_prod(double a, double b, double *p, double *e)
{
  double x = a * b;
  *p = x;
  *e = fma(a, b, -x);
}
typedef struct { double c[1024]; int n; } expansion;

static void exp_grow(expansion *E, double b)
{ /* GROW-EXPANSION with zero elimination */
  double q = b;
  int m = 0;
  for (int i = 0; i < E->n; ++i) {
    double s, e;
    two_sum(q, E->c[i], &s, &e);
    q = s;
    if (e != 0.0) E->c[m++] = e;
  }
  if (q != 0.0 || m == 0) E->c[m++] = q;
  E->n = m;
}
/* add sign * a*b*c*d exactly */
static void exp_add_prod4(expansion *E, double sg, double a, double b, double c, double d)
{
  double p[2], q[4], r[8];
  two_prod(a, b, &p[0], &p[1]);
  for (int i = 0; i < 2; ++i) two_prod(p[i], c, &q[2 * i], &q[2 * i + 1]);
  for (int i = 0; i < 4; ++i) two_prod(q[i], d, &r[2 * i], &r[2 * i + 1]);
  for (int i = 0; i < 8; ++i) if (r[i] != 0.0) exp_grow(E, sg * r[i]);
}
static void exp_add_prod2(expansion *E, double sg, double a, double b)
{
  double p, e;
  two_prod(a, b, &p, &e);
  if (e != 0.0) exp_grow(E, sg * e);
  if (p != 0.0) exp_grow(E, sg * p);
}
static int exp_sign(const expansion *E)
{
  double top = E->c[E->n - 1];
  return (top > 0) - (top < 0);
}

/* sign of (b-a) x (c-a) */
int nmo_orient2d(const double *a, const double *b, const double *c)
{
  double l = (b[0] - a[0]) * (c[1] - a[1]), r = (b[1] - a[1]) * (c[0] - a[0]);
  double det = l - r, bound = 8.0e-16 * (fabs(l) + fabs(r));
  if (det > bound) return 1;
  if (det < -bound) return -1;
  expansion E; E.n = 0;
  /* ax(by - cy) - bx(ay - cy) + cx(ay - by) */
  exp_add_prod2(&E, 1, a[0], b[1]); exp_add_prod2(&E, -1, a[0], c[1]);
  exp_add_prod2(&E, -1, b[0], a[1]); exp_add_prod2(&E, 1, b[0], c[1]);
  exp_add_prod2(&E, 1, c[0], a[1]); exp_add_prod2(&E, -1, c[0], b[1]);
  if (E.n == 0) return 0;
  return exp_sign(&E);
}

/* sign of the 4x4 in-circle determinant: > 0 iff d strictly inside circle(a,b,c), abc ccw */
int nmo_incircle(const double *a, const double *b, const double *c, const double *d)
{
  double adx = a[0] - d[0], ady = a[1] - d[1], bdx = b[0] - d[0], bdy = b[1] - d[1];
  double cdx = c[0] - d[0], cdy = c[1] - d[1];
  double al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;
  double bc1 = bdx * cdy, bc2 = cdx * bdy, ca1 = cdx * ady, ca2 = adx * cdy, ab1 = adx * bdy, ab2 = bdx * ady;
  double det = al * (bc1 - bc2) + bl * (ca1 - ca2) + cl * (ab1 - ab2);
  double perm = al * (fabs(bc1) + fabs(bc2)) + bl * (fabs(ca1) + fabs(ca2)) + cl * (fabs(ab1) + fabs(ab2));
  /* Shewchuk's static filter (10+96eps)eps * permanent covers the rounded translation too */
  double bound = 2.0e-15 * perm;
  if (det > bound) return 1;
  if (det < -bound) return -1;
  const double *P[4] = {a, b, c, d};
  expansion E; E.n = 0;
  for (int k = 0; k < 4; ++k) {
    const double *p = P[(k == 0) ? 1 : 0], *q = P[(k <= 1) ? 2 : 1], *r = P[(k <= 2) ? 3 : 2];
    double sg = (k & 1) ? -1.0 : 1.0;
    for (int t = 0; t < 2; ++t) {
      double l = P[k][t]; /* l*l * M */
      exp_add_prod4(&E, sg, l, l, p[0], q[1]); exp_add_prod4(&E, -sg, l, l, p[0], r[1]);
      exp_add_prod4(&E, -sg, l, l, q[0], p[1]); exp_add_prod4(&E, sg, l, l, q[0], r[1]);
      exp_add_prod4(&E, sg, l, l, r[0], p[1]); exp_add_prod4(&E, -sg, l, l, r[0], q[1]);
    }
  }
  if (E.n == 0) return 0;
  return exp_sign(&E);
}

static const int TRIPLES[4][3] = {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}};

/* Delaunay faces of the xy-projection of the 4 quad vertices (src/intersections.cc:83-84,464).
 * bit t: triangle TRIPLES[t] is a face; NMO_TIE: depends on CGAL insertion order. */
uint8_t nmo_quad_split_code(const double *q /* [4][3] */)
{
  const double *P[4] = {q, q + 3, q + 6, q + 9};
  for (int i = 0; i < 4; ++i)
    for (int j = i + 1; j < 4; ++j)
      if (P[i][0] == P[j][0] && P[i][1] == P[j][1]) {
        int keep[3], n = 0;
        for (int k = 0; k < 4; ++k) if (k != j) keep[n++] = k;
        const double *a = P[keep[0]], *b = P[keep[1]], *c = P[keep[2]];
        int dup = (a[0] == b[0] && a[1] == b[1]) || (b[0] == c[0] && b[1] == c[1]) || (a[0] == c[0] && a[1] == c[1]);
        if (dup || nmo_orient2d(a, b, c) == 0) return NMO_TIE;
        return (uint8_t)(NMO_TIE | (1 << j));
      }
  uint8_t code = 0;
  int tie = 0;
  for (int t = 0; t < 4; ++t) {
    const double *a = P[TRIPLES[t][0]], *b = P[TRIPLES[t][1]], *c = P[TRIPLES[t][2]];
    int o = nmo_orient2d(a, b, c);
    if (o == 0) continue;
    int ic = nmo_incircle(a, b, c, P[t]) * o;
    if (ic < 0) code |= (uint8_t)(1 << t);
    else if (ic == 0) tie = 1;
  }
  if (tie) return (uint8_t)(NMO_TIE | (1 << 2) | (1 << 1));
  return code;
}

void nmo_quad_split(int64_t n, const double *verts, uint8_t *codes)
{
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) codes[i] = nmo_quad_split_code(verts + 12 * i);
}

/* ------------------------------------------------------------------------- */
/* candidate pairs: sorted-bin join, then exact closed-box test               */
/* ------------------------------------------------------------------------- */
typedef struct { uint64_t key; uint32_t id; } bin_entry;
static int cmp_bin(const void *a, const void *b)
{
  const bin_entry *x = a, *y = b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return (x->id > y->id) - (x->id < y->id);
}
static int cmp_u64(const void *a, const void *b)
{
  uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return (x > y) - (x < y);
}
static inline int64_t bin_of(double x, double org, double inv, int64_t nb)
{
  int64_t i = (int64_t)floor((x - org) * inv);
  return i < 0 ? 0 : (i >= nb ? nb - 1 : i);
}

/* returns count; *out = malloc'ed array of (im<<32 | bg) sorted ascending */
int64_t nmo_candidates(int d, int64_t n_bg, const double *bg_lo, const double *bg_hi, int64_t n_im,
                       const double *im_lo, const double *im_hi, uint64_t **out)
{
  double org[3] = {0, 0, 0}, top[3] = {0, 0, 0}, ext = 0;
  for (int k = 0; k < d; ++k) { org[k] = INFINITY; top[k] = -INFINITY; }
  for (int64_t b = 0; b < n_bg; ++b)
    for (int k = 0; k < d; ++k) {
      double l = bg_lo[b * d + k], h = bg_hi[b * d + k];
      if (l < org[k]) org[k] = l;
      if (h > top[k]) top[k] = h;
      if (h - l > ext) ext = h - l;
    }
  if (n_bg == 0 || n_im == 0) { *out = NULL; return 0; }
  if (ext <= 0) ext = 1;
  int64_t nb[3] = {1, 1, 1};
  double inv = 1.0 / ext;
  for (int k = 0; k < d; ++k) {
    nb[k] = (int64_t)ceil((top[k] - org[k]) * inv) + 1;
    if (nb[k] > (1 << 20)) nb[k] = 1 << 20;
  }
  /* bins of width >= max extent: a box spans at most 3 bins per axis (closed) */
  int64_t cap = n_bg * (d == 2 ? 9 : 27), ne = 0;
  bin_entry *E = malloc(sizeof(bin_entry) * (size_t)cap);
  for (int64_t b = 0; b < n_bg; ++b) {
    int64_t l[3] = {0, 0, 0}, h[3] = {0, 0, 0};
    for (int k = 0; k < d; ++k) {
      l[k] = bin_of(bg_lo[b * d + k], org[k], inv, nb[k]);
      h[k] = bin_of(bg_hi[b * d + k], org[k], inv, nb[k]);
    }
    for (int64_t z = l[2]; z <= h[2]; ++z)
      for (int64_t y = l[1]; y <= h[1]; ++y)
        for (int64_t x = l[0]; x <= h[0]; ++x) {
          E[ne].key = ((uint64_t)z * (uint64_t)nb[1] + (uint64_t)y) * (uint64_t)nb[0] + (uint64_t)x;
          E[ne].id = (uint32_t)b;
          ++ne;
        }
  }
  qsort(E, (size_t)ne, sizeof(bin_entry), cmp_bin);
  int64_t ocap = n_im * 8 + 1024, no = 0;
  uint64_t *O = malloc(sizeof(uint64_t) * (size_t)ocap);
  for (int64_t m = 0; m < n_im; ++m) {
    int64_t l[3] = {0, 0, 0}, h[3] = {0, 0, 0};
    for (int k = 0; k < d; ++k) {
      l[k] = bin_of(im_lo[m * d + k], org[k], inv, nb[k]);
      h[k] = bin_of(im_hi[m * d + k], org[k], inv, nb[k]);
    }
    int64_t start = no;
    for (int64_t z = l[2]; z <= h[2]; ++z)
      for (int64_t y = l[1]; y <= h[1]; ++y)
        for (int64_t x = l[0]; x <= h[0]; ++x) {
          uint64_t key = ((uint64_t)z * (uint64_t)nb[1] + (uint64_t)y) * (uint64_t)nb[0] + (uint64_t)x;
          int64_t a = 0, b = ne;
          while (a < b) { int64_t mid = (a + b) >> 1; if (E[mid].key < key) a = mid + 1; else b = mid; }
          for (; a < ne && E[a].key == key; ++a) {
            uint32_t bg = E[a].id;
            int ok = 1;
            for (int k = 0; k < d; ++k)
              ok &= (bg_lo[(int64_t)bg * d + k] <= im_hi[m * d + k]) && (im_lo[m * d + k] <= bg_hi[(int64_t)bg * d + k]);
            if (!ok) continue;
            if (no == ocap) { ocap *= 2; O = realloc(O, sizeof(uint64_t) * (size_t)ocap); }
            O[no++] = ((uint64_t)m << 32) | bg;
          }
        }
    /* a pair can be seen from several bins: sort + unique this immersed cell's run */
    qsort(O + start, (size_t)(no - start), sizeof(uint64_t), cmp_u64);
    int64_t w = start;
    for (int64_t i = start; i < no; ++i) if (i == start || O[i] != O[i - 1]) O[w++] = O[i];
    no = w;
  }
  free(E);
  *out = O;
  return no;
}

void nmo_free(void *p) { free(p); }

/* ------------------------------------------------------------------------- */
/* reference quadrature rules: QGaussSimplex<dim>(n_points_1D)  (intersections.cc:693-695) */
/* ------------------------------------------------------------------------- */
static int rule_1d(int n, double *x, double *w)
{ /* Gauss-Legendre on [0,1] */
  switch (n) {
  case 1: x[0] = 0.5; w[0] = 1.0; return 1;
  case 2: { double a = 0.5 / sqrt(3.0); x[0] = 0.5 - a; x[1] = 0.5 + a; w[0] = w[1] = 0.5; return 2; }
  case 3: { double a = 0.5 * sqrt(0.6); x[0] = 0.5 - a; x[1] = 0.5; x[2] = 0.5 + a;
            w[0] = w[2] = 5.0 / 18.0; w[1] = 4.0 / 9.0; return 3; }
  case 4: { double a = 0.5 * sqrt(3.0 / 7.0 - 2.0 / 7.0 * sqrt(1.2)), b = 0.5 * sqrt(3.0 / 7.0 + 2.0 / 7.0 * sqrt(1.2));
            double wa = (18.0 + sqrt(30.0)) / 72.0, wb = (18.0 - sqrt(30.0)) / 72.0;
            x[0] = 0.5 - b; x[1] = 0.5 - a; x[2] = 0.5 + a; x[3] = 0.5 + b; w[0] = w[3] = wb; w[1] = w[2] = wa; return 4; }
  }
  if (n >= 5 && n <= 8) { /* Newton on P_n: the rules the drivers ask for at degree p >= 2 (n = 2p + 1, 2d3d/lagrange_multiplier.cc:750-753) */
    for (int i = 0; i < n; ++i) {
      double z = cos(3.14159265358979323846 * (i + 0.75) / (n + 0.5)), pp = 1.0;
      for (int it = 0; it < 100; ++it) {
        double p1 = 1.0, p2 = 0.0;
        for (int j = 0; j < n; ++j) { double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0); }
        pp = n * (z * p1 - p2) / (z * z - 1.0);
        double dz = p1 / pp;
        z -= dz;
        if (fabs(dz) < 1e-16) break;
      }
      { double p1 = 1.0, p2 = 0.0;
        for (int j = 0; j < n; ++j) { double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0); }
        pp = n * (z * p1 - p2) / (z * z - 1.0); }
      x[n - 1 - i] = 0.5 + 0.5 * z;
      w[n - 1 - i] = 1.0 / ((1.0 - z * z) * pp * pp);
    }
    return n;
  }
  return 0;
}
static int rule_tri(int n, double (*x)[2], double *w)
{
  if (n == 1) { x[0][0] = x[0][1] = 1.0 / 3.0; w[0] = 0.5; return 1; }
  if (n == 2) {
    double a = 1.0 / 6.0, b = 2.0 / 3.0;
    x[0][0] = a; x[0][1] = a; x[1][0] = b; x[1][1] = a; x[2][0] = a; x[2][1] = b;
    w[0] = w[1] = w[2] = 1.0 / 6.0; return 3;
  }
  if (n == 3) {
    double s15 = sqrt(15.0), a = (6.0 - s15) / 21.0, b = (6.0 + s15) / 21.0;
    double wa = 0.5 * (155.0 - s15) / 1200.0, wb = 0.5 * (155.0 + s15) / 1200.0;
    x[0][0] = x[0][1] = 1.0 / 3.0; w[0] = 0.5 * 9.0 / 40.0;
    x[1][0] = a; x[1][1] = a; x[2][0] = 1 - 2 * a; x[2][1] = a; x[3][0] = a; x[3][1] = 1 - 2 * a;
    x[4][0] = b; x[4][1] = b; x[5][0] = 1 - 2 * b; x[5][1] = b; x[6][0] = b; x[6][1] = 1 - 2 * b;
    w[1] = w[2] = w[3] = wa; w[4] = w[5] = w[6] = wb; return 7;
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* per-pair geometry                                                          */
/* ------------------------------------------------------------------------- */
typedef struct {
  double sumw; double local[8]; int nq; int dropped;
  double *qp, *qw; /* optional sinks for the quadrature points / weights of the pair (nmo_pairs_quadrature) */
} pair_out;

static inline void emit_point(pair_out *o, int d, const double *x, double w)
{
  if (o->qp) {
    for (int k = 0; k < d; ++k) o->qp[(int64_t)o->nq * d + k] = x[k];
    o->qw[o->nq] = w;
  }
}

static inline void q1_accumulate(int d, const double *lo, const double *hi, const double *x, double w, double *local)
{ /* xi = F^{-1}(x) on an axis-aligned cell; phi_i = prod_k (bit_k(i) ? xi_k : 1 - xi_k)
     (include/coupling_utilities.h:248-274 with FE_Q(1) x FE_DGQ(0)) */
  double xi[3];
  for (int k = 0; k < d; ++k) xi[k] = (x[k] - lo[k]) / (hi[k] - lo[k]);
  for (int i = 0; i < (1 << d); ++i) {
    double v = 1.0;
    for (int k = 0; k < d; ++k) v *= ((i >> k) & 1) ? xi[k] : 1.0 - xi[k];
    local[i] += v * w;
  }
}

/* segment ∩ triangle (2D): clip the segment by the three closed half-planes */
static int seg_tri(const double *p0, const double *p1, const double T[3][2], double *t0, double *t1)
{
  double a = 0.0, b = 1.0;
  double area2 = (T[1][0] - T[0][0]) * (T[2][1] - T[0][1]) - (T[1][1] - T[0][1]) * (T[2][0] - T[0][0]);
  double sg = area2 > 0 ? 1.0 : -1.0;
  for (int e = 0; e < 3; ++e) {
    const double *u = T[e], *v = T[(e + 1) % 3];
    double nx = -(v[1] - u[1]) * sg, ny = (v[0] - u[0]) * sg; /* inward normal */
    double f0 = (p0[0] - u[0]) * nx + (p0[1] - u[1]) * ny;
    double f1 = (p1[0] - u[0]) * nx + (p1[1] - u[1]) * ny;
    if (f0 < 0 && f1 < 0) return 0;
    if (f0 < 0) { double t = f0 / (f0 - f1); if (t > a) a = t; }
    else if (f1 < 0) { double t = f0 / (f0 - f1); if (t < b) b = t; }
  }
  if (!(a < b)) return 0; /* point or empty: not a Segment_2 */
  *t0 = a; *t1 = b;
  return 1;
}

static void pair_2d(const double *lo, const double *hi, const double *seg, int n1d, double tol, pair_out *o)
{
  /* quad in CGAL cyclic order, split by the diagonal (c0,c2): intersections.cc:375-378 */
  double c[4][2] = {{lo[0], lo[1]}, {hi[0], lo[1]}, {hi[0], hi[1]}, {lo[0], hi[1]}};
  double T[2][3][2] = {{{c[0][0], c[0][1]}, {c[1][0], c[1][1]}, {c[2][0], c[2][1]}},
                       {{c[0][0], c[0][1]}, {c[2][0], c[2][1]}, {c[3][0], c[3][1]}}};
  double gx[8], gw[8];
  int ng = rule_1d(n1d, gx, gw);
  const double *p0 = seg, *p1 = seg + 2;
  for (int t = 0; t < 2; ++t) {
    double area = 0.5 * fabs((T[t][1][0] - T[t][0][0]) * (T[t][2][1] - T[t][0][1]) - (T[t][1][1] - T[t][0][1]) * (T[t][2][0] - T[t][0][0]));
    if (!(area > tol)) continue; /* :381 */
    double t0, t1;
    if (!seg_tri(p0, p1, T[t], &t0, &t1)) continue;
    double a[2] = {p0[0] + t0 * (p1[0] - p0[0]), p0[1] + t0 * (p1[1] - p0[1])};
    double b[2] = {p0[0] + t1 * (p1[0] - p0[0]), p0[1] + t1 * (p1[1] - p0[1])};
    double J = hypot(b[0] - a[0], b[1] - a[1]);
    if (J < 1e-12) { o->dropped++; continue; } /* :710 */
    for (int q = 0; q < ng; ++q) {
      double x[2] = {a[0] + gx[q] * (b[0] - a[0]), a[1] + gx[q] * (b[1] - a[1])};
      double w = J * gw[q];
      o->sumw += w;
      emit_point(o, 2, x, w);
      o->nq++;
      q1_accumulate(2, lo, hi, x, w, o->local);
    }
  }
}

/* ------------------------------------------------------------------------- */
/* codimension 0 in 2D: background quad n immersed quad (intersections.cc:307-355, compute_intersection_quad_quad) */
/* ------------------------------------------------------------------------- */
/* The reference intersects the two polygons with CGAL, triangulates the result with a constrained Delaunay triangulation,
 * keeps the triangles of area > tol and puts QGaussSimplex<2> on them (intersections.cc:693-746).  Restated the other way
 * round from the CUDA kernel: here the BOX is clipped by the four edges of the (convex) immersed quad, and the fan starts
 * at the vertex of smallest x.  Any triangulation of the same polygon gives the same integral of the bilinear integrand
 * once the rule is exact for degree 2 (n_points_1D >= 2), rule R1. */
static int clip_by_edge_2d(const double (*in)[2], int n, const double *u, const double *v, double sg, double (*out)[2])
{ /* keep the side to the left of u -> v (sg = +1 for a counter-clockwise quad) */
  const double nx = -(v[1] - u[1]) * sg, ny = (v[0] - u[0]) * sg;
  int m = 0;
  for (int i = 0; i < n; ++i) {
    const double *a = in[i], *b = in[(i + 1) % n];
    const double fa = (a[0] - u[0]) * nx + (a[1] - u[1]) * ny, fb = (b[0] - u[0]) * nx + (b[1] - u[1]) * ny;
    if (fa >= 0) { out[m][0] = a[0]; out[m][1] = a[1]; ++m; }
    if ((fa > 0 && fb < 0) || (fa < 0 && fb > 0)) {
      const double t = fa / (fa - fb);
      out[m][0] = a[0] + t * (b[0] - a[0]); out[m][1] = a[1] + t * (b[1] - a[1]); ++m;
    }
  }
  return m;
}

static void pair_2d2d(const double *lo, const double *hi, const double *quad /* [4][2], deal.II order */, int n1d, double tol, pair_out *o)
{
  const int ring[4] = {0, 1, 3, 2};   /* deal.II vertex order -> around the cell */
  double Q[4][2];
  for (int i = 0; i < 4; ++i) { Q[i][0] = quad[2 * ring[i]]; Q[i][1] = quad[2 * ring[i] + 1]; }
  double a2 = 0;
  for (int i = 0; i < 4; ++i) a2 += Q[i][0] * Q[(i + 1) % 4][1] - Q[(i + 1) % 4][0] * Q[i][1];
  const double sg = a2 > 0 ? 1.0 : -1.0;
  double A[16][2] = {{lo[0], lo[1]}, {hi[0], lo[1]}, {hi[0], hi[1]}, {lo[0], hi[1]}}, B[16][2];
  int n = 4;
  for (int e = 0; e < 4 && n >= 3; ++e) {
    n = clip_by_edge_2d((const double(*)[2])A, n, Q[e], Q[(e + 1) % 4], sg, B);
    memcpy(A, B, sizeof(double) * 2 * (size_t)n);
  }
  if (n < 3) return;
  int i0 = 0;
  for (int i = 1; i < n; ++i) if (A[i][0] < A[i0][0]) i0 = i;
  double tx[16][2], tw[16];
  const int ng = rule_tri(n1d, tx, tw);
  for (int j = 1; j + 1 < n; ++j) {
    const double *a = A[i0], *b = A[(i0 + j) % n], *c = A[(i0 + j + 1) % n];
    const double J = fabs((b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]));
    if (!(0.5 * J > tol)) continue;              /* triangle area > tol   intersections.cc:336-337 */
    if (J < 1e-12) { o->dropped++; continue; }   /* :710 */
    for (int q = 0; q < ng; ++q) {
      const double x[2] = {a[0] + tx[q][0] * (b[0] - a[0]) + tx[q][1] * (c[0] - a[0]), a[1] + tx[q][0] * (b[1] - a[1]) + tx[q][1] * (c[1] - a[1])};
      const double w = J * tw[q];
      o->sumw += w;
      emit_point(o, 2, x, w);
      o->nq++;
      q1_accumulate(2, lo, hi, x, w, o->local);
    }
  }
}

/* nmo_pairs / nmo_pairs_quadrature for <2,2,2>: im_verts = [n][4][2] */
int nmo_pairs_codim0(int64_t n_cand, const uint64_t *cand, const double *bg_lo, const double *bg_hi, const double *im_verts, int n1d,
                     double tol, double *sumw, double *local, int32_t *nq, int64_t *n_dropped)
{
  double tx[16][2], tw[16];
  if (rule_tri(n1d, tx, tw) == 0) return -1;
  int64_t dropped = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : dropped)
  for (int64_t i = 0; i < n_cand; ++i) {
    uint32_t m = (uint32_t)(cand[i] >> 32), b = (uint32_t)cand[i];
    pair_out o;
    memset(&o, 0, sizeof(o));
    pair_2d2d(bg_lo + 2 * (int64_t)b, bg_hi + 2 * (int64_t)b, im_verts + 8 * (int64_t)m, n1d, tol, &o);
    sumw[i] = o.sumw;
    for (int k = 0; k < 4; ++k) local[i * 4 + k] = o.local[k];
    if (nq) nq[i] = o.nq;
    dropped += o.dropped;
  }
  if (n_dropped) *n_dropped = dropped;
  return 0;
}

int nmo_pairs_quadrature_codim0(int64_t n, const uint64_t *pairs, const double *bg_lo, const double *bg_hi, const double *im_verts, int n1d,
                                double tol, const uint64_t *q_off, double *pts, double *wts)
{
  double tx[16][2], tw[16];
  if (rule_tri(n1d, tx, tw) == 0) return -1;
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : bad)
  for (int64_t i = 0; i < n; ++i) {
    uint32_t m = (uint32_t)(pairs[i] >> 32), b = (uint32_t)pairs[i];
    pair_out o;
    memset(&o, 0, sizeof(o));
    o.qp = pts + (int64_t)q_off[i] * 2;
    o.qw = wts + q_off[i];
    pair_2d2d(bg_lo + 2 * (int64_t)b, bg_hi + 2 * (int64_t)b, im_verts + 8 * (int64_t)m, n1d, tol, &o);
    bad += (uint64_t)o.nq != q_off[i + 1] - q_off[i];
  }
  return bad ? -2 : 0;
}

typedef struct { double n[3]; double c; } plane; /* inside: n.x >= c */

static int clip_poly(const double (*in)[3], int n, const plane *P, double (*out)[3])
{
  int m = 0;
  for (int i = 0; i < n; ++i) {
    const double *p = in[i], *q = in[(i + 1) % n];
    double dp = P->n[0] * p[0] + P->n[1] * p[1] + P->n[2] * p[2] - P->c;
    double dq = P->n[0] * q[0] + P->n[1] * q[1] + P->n[2] * q[2] - P->c;
    if (dp >= 0) { out[m][0] = p[0]; out[m][1] = p[1]; out[m][2] = p[2]; ++m; }
    if ((dp > 0 && dq < 0) || (dp < 0 && dq > 0)) {
      double t = dp / (dp - dq);
      for (int k = 0; k < 3; ++k) out[m][k] = p[k] + t * (q[k] - p[k]);
      ++m;
    }
  }
  return m;
}

static void tri_quadrature(const double *lo, const double *hi, const double *a, const double *b, const double *c,
                           int ng, double (*gx)[2], const double *gw, double tol, pair_out *o)
{
  double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
  double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
  double J = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]); /* 2 * area */
  if (!(0.25 * J * J > tol * tol)) return;                    /* squared_area > tol^2, :477,493 */
  if (J < 1e-12) { o->dropped++; return; }                    /* :710 */
  for (int q = 0; q < ng; ++q) {
    double x[3];
    for (int k = 0; k < 3; ++k) x[k] = a[k] + gx[q][0] * e1[k] + gx[q][1] * e2[k];
    double w = J * gw[q];
    o->sumw += w;
    emit_point(o, 3, x, w);
    o->nq++;
    q1_accumulate(3, lo, hi, x, w, o->local);
  }
}

static void pair_3d(const double *lo, const double *hi, const double *quad, uint8_t code, int n1d, double tol, pair_out *o)
{
  static const int perm[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
  double gx[16][2], gw[16];
  int ng = rule_tri(n1d, gx, gw);
  double h[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
  for (int t = 0; t < 4; ++t) {
    if (!((code >> t) & 1)) continue;
    double tri[3][3];
    for (int v = 0; v < 3; ++v) memcpy(tri[v], quad + 3 * TRIPLES[t][v], 3 * sizeof(double));
    for (int s = 0; s < 6; ++s) {
      /* Kuhn tetrahedron: 0 <= u_{p0} <= u_{p1} <= u_{p2} <= 1 with u_k = (x_k - lo_k)/h_k.
         Written on physical coordinates as planes n.x >= c. */
      const int *p = perm[s];
      plane P[4];
      memset(P, 0, sizeof(P));
      P[0].n[p[0]] = 1.0; P[0].c = lo[p[0]];                                   /* u_p0 >= 0 */
      P[1].n[p[1]] = 1.0 / h[p[1]]; P[1].n[p[0]] = -1.0 / h[p[0]];
      P[1].c = lo[p[1]] / h[p[1]] - lo[p[0]] / h[p[0]];                        /* u_p1 >= u_p0 */
      P[2].n[p[2]] = 1.0 / h[p[2]]; P[2].n[p[1]] = -1.0 / h[p[1]];
      P[2].c = lo[p[2]] / h[p[2]] - lo[p[1]] / h[p[1]];                        /* u_p2 >= u_p1 */
      P[3].n[p[2]] = -1.0; P[3].c = -hi[p[2]];                                 /* u_p2 <= 1 */
      double A[16][3], B[16][3];
      int n = 3;
      memcpy(A, tri, sizeof(tri));
      for (int f = 0; f < 4 && n >= 3; ++f) {
        n = clip_poly((const double(*)[3])A, n, &P[f], B);
        memcpy(A, B, sizeof(double) * 3 * (size_t)n);
      }
      if (n < 3) continue;
      /* Triangle_3 result or polygon re-triangulated (fan; the reference uses a Delaunay of
         the polygon's xy-projection -- same union) */
      for (int j = 1; j + 1 < n; ++j) tri_quadrature(lo, hi, A[0], A[j], A[j + 1], ng, gx, gw, tol, o);
    }
  }
}

/* For every candidate (key = im<<32|bg): sum of weights, local vector, #points.
 * Returns 0, or -1 for an unsupported rule. */
int nmo_pairs(int d, int64_t n_cand, const uint64_t *cand, const double *bg_lo, const double *bg_hi,
              const double *im_verts, const uint8_t *codes, int n1d, double tol, double *sumw, double *local,
              int32_t *nq, int64_t *n_dropped)
{
  double tx[16][2], tw[16];
  if (d == 2 ? rule_1d(n1d, tw, tw) == 0 : rule_tri(n1d, tx, tw) == 0) return -1;
  int64_t dropped = 0;
  const int nl = 1 << d;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : dropped)
  for (int64_t i = 0; i < n_cand; ++i) {
    uint32_t m = (uint32_t)(cand[i] >> 32), b = (uint32_t)cand[i];
    pair_out o;
    memset(&o, 0, sizeof(o));
    if (d == 2)
      pair_2d(bg_lo + 2 * (int64_t)b, bg_hi + 2 * (int64_t)b, im_verts + 4 * (int64_t)m, n1d, tol, &o);
    else
      pair_3d(bg_lo + 3 * (int64_t)b, bg_hi + 3 * (int64_t)b, im_verts + 12 * (int64_t)m, codes[m], n1d, tol, &o);
    sumw[i] = o.sumw;
    for (int k = 0; k < nl; ++k) local[i * nl + k] = o.local[k];
    if (nq) nq[i] = o.nq;
    dropped += o.dropped;
  }
  if (n_dropped) *n_dropped = dropped;
  return 0;
}

/* The Quadrature<spacedim> of every given pair, in the reference's own subdivision (2D: the two CDT triangles of the
 * background quad, intersections.cc:375-388; 3D: tetrahedra x Delaunay triangles, :459-503), as
 * collect_quadratures_on_overlapped_grids returns it in its tuples (coupling_utilities.cc:57-66).
 * q_off[n+1] must hold the offsets built from the point counts nmo_pairs reported. */
int nmo_pairs_quadrature(int d, int64_t n, const uint64_t *pairs, const double *bg_lo, const double *bg_hi, const double *im_verts,
                         const uint8_t *codes, int n1d, double tol, const uint64_t *q_off, double *pts, double *wts)
{
  double tx[16][2], tw[16];
  if (d == 2 ? rule_1d(n1d, tw, tw) == 0 : rule_tri(n1d, tx, tw) == 0) return -1;
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : bad)
  for (int64_t i = 0; i < n; ++i) {
    uint32_t m = (uint32_t)(pairs[i] >> 32), b = (uint32_t)pairs[i];
    pair_out o;
    memset(&o, 0, sizeof(o));
    o.qp = pts + (int64_t)q_off[i] * d;
    o.qw = wts + q_off[i];
    if (d == 2)
      pair_2d(bg_lo + 2 * (int64_t)b, bg_hi + 2 * (int64_t)b, im_verts + 4 * (int64_t)m, n1d, tol, &o);
    else
      pair_3d(bg_lo + 3 * (int64_t)b, bg_hi + 3 * (int64_t)b, im_verts + 12 * (int64_t)m, codes[m], n1d, tol, &o);
    bad += (uint64_t)o.nq != q_off[i + 1] - q_off[i];
  }
  return bad ? -2 : 0;
}

/* ------------------------------------------------------------------------- */
/* sparsity + values of the stored (n_space x n_immersed) matrix              */
/* ------------------------------------------------------------------------- */
typedef struct { uint64_t key; double v; } coo;
static int cmp_coo(const void *a, const void *b)
{
  const coo *x = a, *y = b;
  return (x->key > y->key) - (x->key < y->key);
}

/* accepted pairs only (sumw > tol).  line_of[dof] = index into c_ptr or -1.
 * Output CSR (malloc'ed): rowptr[n_space+1], col, val.  Returns nnz. */
int64_t nmo_build_csr(int d, int64_t n_acc, const uint64_t *acc, const double *local, const uint32_t *bg_dofs,
                      const uint32_t *im_dofs, int64_t n_space, const int32_t *line_of, const int64_t *c_ptr,
                      const uint32_t *c_master, const double *c_weight, uint64_t **rowptr_out, uint32_t **col_out,
                      double **val_out)
{
  const int nl = 1 << d;
  int64_t cap = n_acc * nl * 5 + 16, n = 0;
  coo *E = malloc(sizeof(coo) * (size_t)cap);
  for (int64_t p = 0; p < n_acc; ++p) {
    uint32_t m = (uint32_t)(acc[p] >> 32), b = (uint32_t)acc[p];
    uint32_t col = im_dofs[m];
    for (int i = 0; i < nl; ++i) {
      uint32_t r = bg_dofs[(int64_t)b * nl + i];
      double v = local[p * nl + i];
      int32_t ln = line_of ? line_of[r] : -1;
      int64_t need = 2 + (ln >= 0 ? c_ptr[ln + 1] - c_ptr[ln] : 0);
      if (n + need > cap) { cap = cap * 2 + need; E = realloc(E, sizeof(coo) * (size_t)cap); }
      if (ln < 0) {
        E[n].key = ((uint64_t)r << 32) | col; E[n].v = v; ++n;
      } else {
        E[n].key = ((uint64_t)r << 32) | col; E[n].v = 0.0; ++n;  /* keep_constrained_entries: structural zero */
        for (int64_t k = c_ptr[ln]; k < c_ptr[ln + 1]; ++k) {
          E[n].key = ((uint64_t)c_master[k] << 32) | col; E[n].v = c_weight[k] * v; ++n;
        }
      }
    }
  }
  qsort(E, (size_t)n, sizeof(coo), cmp_coo);
  int64_t nnz = 0;
  for (int64_t i = 0; i < n; ++i) if (i == 0 || E[i].key != E[i - 1].key) ++nnz;
  uint64_t *rp = calloc((size_t)n_space + 1, sizeof(uint64_t));
  uint32_t *cl = malloc(sizeof(uint32_t) * (size_t)(nnz ? nnz : 1));
  double *vl = malloc(sizeof(double) * (size_t)(nnz ? nnz : 1));
  int64_t w = -1;
  for (int64_t i = 0; i < n; ++i) {
    if (i == 0 || E[i].key != E[i - 1].key) {
      ++w;
      cl[w] = (uint32_t)E[i].key;
      vl[w] = 0.0;
      rp[(E[i].key >> 32) + 1]++;
    }
    vl[w] += E[i].v;
  }
  for (int64_t r = 0; r < n_space; ++r) rp[r + 1] += rp[r];
  free(E);
  *rowptr_out = rp; *col_out = cl; *val_out = vl;
  return nnz;
}

/* ------------------------------------------------------------------------- */
/* interface-penalisation (Nitsche) terms on a given quadrature               */
/* ------------------------------------------------------------------------- */
/* code/include/coupling_utilities.h:306-398 (matrix) and :400-481 (right-hand side) for FE_Q(1) on axis-aligned
   cells: per tuple local(i,j) += c_q (penalty/h) phi_i phi_j w_q (:366-370), rhs(i) += c_q (penalty/h) phi_i f_q w_q
   (:447-451), h = cell->diameter() (:357,:429), then AffineConstraints::distribute_local_to_global
   [UPSTREAM deal.II 9.4 affine_constraints.templates.h: rows and columns resolved through the lines; for every
   constrained local dof |local(i,i)|, or the mean of |local(k,k)| when that is zero, is added on its diagonal].
   Output: unsorted COO records (duplicates to be summed by the caller) and a dense rhs. */
int64_t nmo_nitsche(int d, int64_t n_pairs, const uint32_t *bg, const uint64_t *q_off, const double *pts, const double *wts,
                    const double *coeff, const double *fval, double penalty, const double *bg_lo, const double *bg_hi,
                    const uint32_t *bg_dofs, const int32_t *line_of, const int64_t *c_ptr, const uint32_t *c_master,
                    const double *c_weight, uint32_t **row_out, uint32_t **col_out, double **val_out, double *rhs)
{
  const int nl = 1 << d;
  int64_t cap = n_pairs * nl * nl + 64, n = 0;
  uint32_t *R = malloc(sizeof(uint32_t) * (size_t)cap), *Cc = malloc(sizeof(uint32_t) * (size_t)cap);
  double *V = malloc(sizeof(double) * (size_t)cap);
  for (int64_t p = 0; p < n_pairs; ++p) {
    const double *lo = bg_lo + (int64_t)bg[p] * d, *hi = bg_hi + (int64_t)bg[p] * d;
    double h = 0, L[8][8], F[8];
    for (int k = 0; k < d; ++k) h += (hi[k] - lo[k]) * (hi[k] - lo[k]);
    h = sqrt(h);
    memset(L, 0, sizeof L);
    memset(F, 0, sizeof F);
    for (uint64_t q = q_off[p]; q < q_off[p + 1]; ++q) {
      double xi[3], phi[8];
      for (int k = 0; k < d; ++k) xi[k] = (pts[q * d + k] - lo[k]) / (hi[k] - lo[k]);
      for (int i = 0; i < nl; ++i) {
        double v = 1.0;
        for (int k = 0; k < d; ++k) v *= ((i >> k) & 1) ? xi[k] : 1.0 - xi[k];
        phi[i] = v;
      }
      const double c = coeff ? coeff[q] : 1.0;
      for (int i = 0; i < nl; ++i) {
        for (int j = 0; j < nl; ++j) L[i][j] += c * (penalty / h) * phi[i] * phi[j] * wts[q];
        if (fval) F[i] += c * (penalty / h) * phi[i] * fval[q] * wts[q];
      }
    }
    const uint32_t *dof = bg_dofs + (int64_t)bg[p] * nl;
    double avg = 0;
    for (int i = 0; i < nl; ++i) avg += fabs(L[i][i]);
    avg /= nl;
    for (int i = 0; i < nl; ++i) {
      const int32_t li = line_of ? line_of[dof[i]] : -1;
      const int64_t ib = li < 0 ? 0 : c_ptr[li], ie = li < 0 ? 1 : c_ptr[li + 1];
      for (int64_t a = ib; a < ie; ++a) {
        const uint32_t r = li < 0 ? dof[i] : c_master[a];
        const double wr = li < 0 ? 1.0 : c_weight[a];
        if (rhs && fval) rhs[r] += wr * F[i];
        for (int j = 0; j < nl; ++j) {
          const int32_t lj = line_of ? line_of[dof[j]] : -1;
          const int64_t jb = lj < 0 ? 0 : c_ptr[lj], je = lj < 0 ? 1 : c_ptr[lj + 1];
          for (int64_t b = jb; b < je; ++b) {
            if (n + 2 > cap) {
              cap *= 2;
              R = realloc(R, sizeof(uint32_t) * (size_t)cap); Cc = realloc(Cc, sizeof(uint32_t) * (size_t)cap);
              V = realloc(V, sizeof(double) * (size_t)cap);
            }
            R[n] = r; Cc[n] = lj < 0 ? dof[j] : c_master[b];
            V[n] = wr * (lj < 0 ? 1.0 : c_weight[b]) * L[i][j];
            ++n;
          }
        }
      }
      if (li >= 0) {
        if (n + 2 > cap) {
          cap *= 2;
          R = realloc(R, sizeof(uint32_t) * (size_t)cap); Cc = realloc(Cc, sizeof(uint32_t) * (size_t)cap);
          V = realloc(V, sizeof(double) * (size_t)cap);
        }
        R[n] = dof[i]; Cc[n] = dof[i]; V[n] = fabs(L[i][i]) != 0.0 ? fabs(L[i][i]) : avg; ++n;
      }
    }
  }
  *row_out = R; *col_out = Cc; *val_out = V;
  return n;
}

/* y = A x (transpose = 0) or y = A^T x (transpose = 1); y must be zero-length-correct */
void nmo_spmv(int transpose, int64_t n_rows, int64_t n_cols, const uint64_t *rowptr, const uint32_t *col,
              const double *val, const double *x, double *y)
{
  if (!transpose) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_rows; ++r) {
      double s = 0.0;
      for (uint64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) s += val[k] * x[col[k]];
      y[r] = s;
    }
  } else {
    for (int64_t c = 0; c < n_cols; ++c) y[c] = 0.0;
    for (int64_t r = 0; r < n_rows; ++r)
      for (uint64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) y[col[k]] += val[k] * x[r];
  }
}

int nmo_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void nmo_set_num_threads(int n)
{
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
