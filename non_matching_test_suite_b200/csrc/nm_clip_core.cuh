// Geometry core of the 3D clip kernels: triangle-vs-box gate, clipped polygon, closed-form moments.  Pure functions of
// their arguments, compiled for the device by nm_clip.cu and for the HOST by tests/cpp/clip_core_host.cpp, which checks
// them pair by pair against an independent CPU restatement (reference: code/src/intersections.cc:437-509, 693-746).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define NMC_FN __device__ __forceinline__
#define NMC_RCP(x) __drcp_rn(x)
#define NMC_RSQRT(x) rsqrt(x)
#define NMC_FFS64(x) __ffsll((long long)(x))
#define NMC_POPC64(x) __popcll(x)
#define NMC_FFS32(x) __ffs(x)
#define NMC_UNROLL _Pragma("unroll")
#else
#define NMC_UNROLL
#ifdef __CUDACC__
#define NMC_FN __host__ __device__ inline
#else
#define NMC_FN inline
#endif
#define NMC_RCP(x) (1.0 / (x))
#define NMC_RSQRT(x) (1.0 / sqrt(x))
#define NMC_FFS64(x) __builtin_ffsll((long long)(x))
#define NMC_POPC64(x) __builtin_popcountll(x)
#define NMC_FFS32(x) __builtin_ffs((int)(x))
#endif

// ---- build switches ----
// 0: clip the triangle by the box planes (the kernel's algorithm); 1: section algorithm (measured on B200 at sphere
// cycle 8: 22.6 ms against 14.4 ms -- kept as a cross-check of the polygon construction, tests/test_clip_core_cpu.py)
#ifndef NM_CLIP_SECTION
#define NM_CLIP_SECTION 0
#endif
// outcode layout of the box-plane clip: 1 = position-major 64-bit list (item_clip_polygon_v1), 2 = plane-major 32-bit words
#ifndef NM_CLIP_BITS
#define NM_CLIP_BITS 2
#endif
// pool slot of a new vertex, see item_clip_polygon_v2 (1 is what a nine-slot pool needs)
#ifndef NM_CLIP_SLOTS
#define NM_CLIP_SLOTS 0
#endif
// plane distances of the four edge end points: 0 = selected from the loaded coordinates, 1 = fetched by (axis, vertex) address
#ifndef NM_CLIP_AXISLOAD
#define NM_CLIP_AXISLOAD 0
#endif

namespace nmclip {

// Separating-axis test triangle vs box [0,H] (box normals, triangle normal, 9 edge cross
// products).  Conservative on equality (touching = overlap); used only to keep triangles that
// cannot contribute out of the clipping stage.
template <bool CROSS_AXES>
NMC_FN bool tri_box_overlap(const double (*v)[3], const double *H)
{
  const double h[3] = {0.5 * H[0], 0.5 * H[1], 0.5 * H[2]};
  double p[3][3];
NMC_UNROLL
  for (int i = 0; i < 3; ++i)
NMC_UNROLL
    for (int k = 0; k < 3; ++k) p[i][k] = v[i][k] - h[k];
  bool sep = false;
NMC_UNROLL
  for (int k = 0; k < 3; ++k) {
    // min > h or max < -h, written as comparisons: an FP64 min/max costs ~8 instructions on sm_100 (DSETP + selects)
    sep |= (p[0][k] > h[k] && p[1][k] > h[k] && p[2][k] > h[k]) || (p[0][k] < -h[k] && p[1][k] < -h[k] && p[2][k] < -h[k]);
  }
  const double f[3][3] = {{p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]},
                          {p[2][0] - p[1][0], p[2][1] - p[1][1], p[2][2] - p[1][2]},
                          {p[0][0] - p[2][0], p[0][1] - p[2][1], p[0][2] - p[2][2]}};
  {
    const double nx = f[0][1] * f[1][2] - f[0][2] * f[1][1], ny = f[0][2] * f[1][0] - f[0][0] * f[1][2], nz = f[0][0] * f[1][1] - f[0][1] * f[1][0];
    const double d = nx * p[0][0] + ny * p[0][1] + nz * p[0][2];
    const double r = h[0] * fabs(nx) + h[1] * fabs(ny) + h[2] * fabs(nz);
    sep |= fabs(d) > r;
  }
  if (CROSS_AXES) {
  // axes e_i x f_j: the two end points of edge j project to the same value, so two projections
  // (edge start, opposite vertex) bound the triangle
NMC_UNROLL
  for (int j = 0; j < 3; ++j) {
    const int ia = j, ib = (j + 2) % 3;
    {  // e_x x f_j = (0, -fz, fy)
      const double a = -f[j][2], b = f[j][1];
      const double q0 = a * p[ia][1] + b * p[ia][2], q1 = a * p[ib][1] + b * p[ib][2];
      const double r = h[1] * fabs(a) + h[2] * fabs(b);
      sep |= (q0 > r && q1 > r) || (q0 < -r && q1 < -r);
    }
    {  // e_y x f_j = (fz, 0, -fx)
      const double a = f[j][2], b = -f[j][0];
      const double q0 = a * p[ia][0] + b * p[ia][2], q1 = a * p[ib][0] + b * p[ib][2];
      const double r = h[0] * fabs(a) + h[2] * fabs(b);
      sep |= (q0 > r && q1 > r) || (q0 < -r && q1 < -r);
    }
    {  // e_z x f_j = (-fy, fx, 0)
      const double a = -f[j][1], b = f[j][0];
      const double q0 = a * p[ia][0] + b * p[ia][1], q1 = a * p[ib][0] + b * p[ib][1];
      const double r = h[0] * fabs(a) + h[1] * fabs(b);
      sep |= (q0 > r && q1 > r) || (q0 < -r && q1 < -r);
    }
  }
  }
  return !sep;
}


struct ItemOut {
  double area2;   // sum of J = 2 * area over kept sub-triangles
  double m[7];    // A-weighted sums for u, v, w, uv, uw, vw, uvw (constants applied at the end)
  unsigned nfan, dropped;
};

// Stage-2 work of one (candidate, triangle) item.
//
// The polygon is never copied: vertices live in a pool (3 originals + at most 2 per plane <= 15,
// local memory); the polygon is a list of 4-bit pool indices packed in one 64-bit register, and a
// second register holds, position by position, the 6-bit Cohen-Sutherland outcode of each vertex
// (one "inside" bit per box plane, computed once when the vertex is created).  Classifying the
// polygon against a plane is a mask of that register, the inside run is found with ffs on the
// strided bits, and the new polygon is spliced with shifts -- no loop over vertices and no memory
// traffic except for the (at most two) edge/plane crossings.
struct VPool {
  double c[3][16];
  NMC_FN double get(int k, int i) const { return c[k][i]; }
  NMC_FN void set(int k, int i, double v) { c[k][i] = v; }
};

NMC_FN unsigned outcode(double x, double y, double z, const double *H)
{
  return (x >= 0.0 ? 1u : 0u) | (x <= H[0] ? 2u : 0u) | (y >= 0.0 ? 4u : 0u) | (y <= H[1] ? 8u : 0u) | (z >= 0.0 ? 16u : 0u) |
         (z <= H[2] ? 32u : 0u);
}

// Clipped polygon of one (cell, triangle) item: pool vertices in cell-local coordinates, index list,
// vertex count, unit normal of the parent triangle.  Shared by the moment kernel and the quadrature
// emission so that both see exactly the same fan triangles (same count of points per pair).
template <class Pool> struct ClippedPolyT {
  Pool P;
  unsigned long long poly;
  int n;
  double nx, ny, nz;
};
typedef ClippedPolyT<VPool> ClippedPoly;

NMC_FN bool item_clip_polygon_v1(const double *box, const double *quad, int t, ClippedPoly &C)
{
  VPool &P = C.P;
  const double lo[3] = {box[0], box[1], box[2]};
  const double H[3] = {box[4] - lo[0], box[5] - lo[1], box[6] - lo[2]};
  const int iv[3] = {t == 0 ? 1 : 0, t <= 1 ? 2 : 1, t <= 2 ? 3 : 2};
  unsigned long long code = 0;         // 6 bits per polygon position
NMC_UNROLL
  for (int v = 0; v < 3; ++v) {
    const double x = quad[iv[v] * 3 + 0] - lo[0], y = quad[iv[v] * 3 + 1] - lo[1], z = quad[iv[v] * 3 + 2] - lo[2];
    P.c[0][v] = x; P.c[1][v] = y; P.c[2][v] = z;
    code |= (unsigned long long)outcode(x, y, z, H) << (6 * v);
  }
  // unit normal of the parent triangle: J of a coplanar sub-triangle = |cross . n|
  double nx, ny, nz;
  {
    const double e1[3] = {P.c[0][1] - P.c[0][0], P.c[1][1] - P.c[1][0], P.c[2][1] - P.c[2][0]};
    const double e2[3] = {P.c[0][2] - P.c[0][0], P.c[1][2] - P.c[1][0], P.c[2][2] - P.c[2][0]};
    nx = e1[1] * e2[2] - e1[2] * e2[1]; ny = e1[2] * e2[0] - e1[0] * e2[2]; nz = e1[0] * e2[1] - e1[1] * e2[0];
    const double n2 = nx * nx + ny * ny + nz * nz;
    if (!(n2 > 0.0)) return false;
    const double inv = NMC_RSQRT(n2);
    nx *= inv; ny *= inv; nz *= inv;
  }
  unsigned long long poly = 0x210ULL;  // pool indices, 4 bits per polygon position
  int n = 3, nv = 3;
  constexpr unsigned long long M6 = 0x0041041041041041ULL;  // bit 0 of every 6-bit group
  // Every lane walks only the planes that cut ITS polygon (lowest first): the trip count of the warp is the largest
  // number of cutting planes of a lane, not six, and no lane idles through planes it does not need.  A plane that has
  // been applied is all-inside afterwards (new vertices carry `forced`), so it is never picked again.
#pragma unroll 1
  while (true) {
    // AND of the outcodes of all n positions (positions >= n read as inside)
    const unsigned long long cf = code | ~((1ULL << (6 * n)) - 1ULL);
    const unsigned long long a1 = cf & (cf >> 30);            // groups 0..4 &= groups 5..9
    const unsigned long long a2 = a1 & (a1 >> 12);            // group 0 &= 2, group 1 &= 3
    const unsigned all_in = (unsigned)(a2 & (a2 >> 6) & (a1 >> 24)) & 63u;
    if (all_in == 63u) break;
    const int p = NMC_FFS32(~all_in & 63u) - 1;
    const unsigned long long full6 = M6 & ((1ULL << (6 * n)) - 1ULL);
    const unsigned long long in6 = (code >> p) & full6;
    if (in6 == 0ULL) return false;
    const int axis = p >> 1;
    const double sgn = (p & 1) ? -1.0 : 1.0, bound = (p & 1) ? H[axis] : 0.0;
    const unsigned long long prev6 = ((in6 << 6) | (in6 >> (6 * (n - 1)))) & full6;       // position i: vertex i-1 inside
    const int s6 = NMC_FFS64(in6 & ~prev6) - 1;                                  // run start (bit index)
    const int s = NMC_POPC64(M6 & ((1ULL << s6) - 1ULL));
    const unsigned long long rot6 = ((in6 >> s6) | (in6 << (6 * n - s6))) & full6;
    const int l6 = NMC_FFS64(~rot6 & full6) - 1;                                 // run length (bit index), < 6n
    const int len = NMC_POPC64(M6 & ((1ULL << l6) - 1ULL));
    // both lists rotated so that the run comes first
    const unsigned long long pr = (poly >> (4 * s)) | (poly << (4 * (n - s)));
    const unsigned long long cr = (code >> s6) | (code << (6 * n - s6));
    const unsigned forced = (2u << p) - 1u;   // planes already applied count as inside for a new vertex
    unsigned long long np = 0, nc = 0;
    int m = 0;
    // The pool lives in local memory: fetch the four end points of the entering and the leaving edge in ONE round of
    // loads (all indices come from the bit lists), then work in registers.
    const int ve_p = (int)((poly >> (4 * (s == 0 ? n - 1 : s - 1))) & 15u), ve_q = (int)(pr & 15u);        // outside -> first inside
    const int vl_p = (int)((pr >> (4 * (len - 1))) & 15u), vl_q = (int)((pr >> (4 * len)) & 15u);           // last inside -> outside
    double ep[3], eq[3], lp[3], lq[3];
NMC_UNROLL
    for (int k = 0; k < 3; ++k) { ep[k] = P.c[k][ve_p]; eq[k] = P.c[k][ve_q]; lp[k] = P.c[k][vl_p]; lq[k] = P.c[k][vl_q]; }
    const double e_dp = sgn * ((axis == 0 ? ep[0] : axis == 1 ? ep[1] : ep[2]) - bound);
    const double e_dq = sgn * ((axis == 0 ? eq[0] : axis == 1 ? eq[1] : eq[2]) - bound);
    const double l_dp = sgn * ((axis == 0 ? lp[0] : axis == 1 ? lp[1] : lp[2]) - bound);
    const double l_dq = sgn * ((axis == 0 ? lq[0] : axis == 1 ? lq[1] : lq[2]) - bound);
    if (e_dq > 0.0) {  // entering edge
      const double tt = e_dp * NMC_RCP(e_dp - e_dq);
      double x[3];
NMC_UNROLL
      for (int k = 0; k < 3; ++k) x[k] = ep[k] + tt * (eq[k] - ep[k]);
      if (axis == 0) x[0] = bound; else if (axis == 1) x[1] = bound; else x[2] = bound;  // on the plane by construction
NMC_UNROLL
      for (int k = 0; k < 3; ++k) P.c[k][nv] = x[k];
      np = (unsigned long long)nv;
      nc = (unsigned long long)(outcode(x[0], x[1], x[2], H) | forced);
      m = 1;
      ++nv;
    }
    np |= (pr & ((1ULL << (4 * len)) - 1ULL)) << (4 * m);
    nc |= (cr & ((1ULL << (6 * len)) - 1ULL)) << (6 * m);
    m += len;
    if (l_dp > 0.0) {  // leaving edge
      const double tt = l_dp * NMC_RCP(l_dp - l_dq);
      double x[3];
NMC_UNROLL
      for (int k = 0; k < 3; ++k) x[k] = lp[k] + tt * (lq[k] - lp[k]);
      if (axis == 0) x[0] = bound; else if (axis == 1) x[1] = bound; else x[2] = bound;
NMC_UNROLL
      for (int k = 0; k < 3; ++k) P.c[k][nv] = x[k];
      np |= (unsigned long long)nv << (4 * m);
      nc |= (unsigned long long)(outcode(x[0], x[1], x[2], H) | forced) << (6 * m);
      ++m;
      ++nv;
    }
    poly = np;
    code = nc;
    n = m;
    if (n < 3) return false;
  }
  C.poly = poly; C.n = n; C.nx = nx; C.ny = ny; C.nz = nz;
  return true;
}

// The same clip with the outcodes kept PLANE-major in two 32-bit words: word A holds planes 0..2 (x >= 0, x <= Hx,
// y >= 0), word B planes 3..5 (y <= Hy, z >= 0, z <= Hz); plane q of a word owns bits [10 q, 10 q + n), bit i = polygon
// position i is inside that plane (n <= 9).  The inside mask of the plane being applied is then one bit field: run start
// and run length are single ffs instructions on 32-bit values, and the three fields of a word are rotated / spliced
// together (multiplication by 0x00100401 replicates a mask into the fields).  Same planes in the same order, same run,
// same vertices as item_clip_polygon_v1 -- the two are compared candidate by candidate on the host
// (tests/test_clip_core_cpu.py).
NMC_FN void outcode2(double x, double y, double z, const double *H, unsigned &a, unsigned &b)
{
  a = (x >= 0.0 ? 1u : 0u) | (x <= H[0] ? 1u << 10 : 0u) | (y >= 0.0 ? 1u << 20 : 0u);
  b = (y <= H[1] ? 1u : 0u) | (z >= 0.0 ? 1u << 10 : 0u) | (z <= H[2] ? 1u << 20 : 0u);
}

template <class Pool> NMC_FN bool item_clip_polygon_v2(const double *box, const double *quad, int t, ClippedPolyT<Pool> &C)
{
  Pool &P = C.P;
  constexpr unsigned REP = 0x00100401u;   // bit 0 of the three fields of a word
  const double lo[3] = {box[0], box[1], box[2]};
  const double H[3] = {box[4] - lo[0], box[5] - lo[1], box[6] - lo[2]};
  const int iv[3] = {t == 0 ? 1 : 0, t <= 1 ? 2 : 1, t <= 2 ? 3 : 2};
  unsigned A = 0, B = 0;
NMC_UNROLL
  for (int v = 0; v < 3; ++v) {
    const double x = quad[iv[v] * 3 + 0] - lo[0], y = quad[iv[v] * 3 + 1] - lo[1], z = quad[iv[v] * 3 + 2] - lo[2];
    P.set(0, v, x); P.set(1, v, y); P.set(2, v, z);
    unsigned a, b;
    outcode2(x, y, z, H, a, b);
    A |= a << v; B |= b << v;
  }
  double nx, ny, nz;
  {
    const double e1[3] = {P.get(0, 1) - P.get(0, 0), P.get(1, 1) - P.get(1, 0), P.get(2, 1) - P.get(2, 0)};
    const double e2[3] = {P.get(0, 2) - P.get(0, 0), P.get(1, 2) - P.get(1, 0), P.get(2, 2) - P.get(2, 0)};
    nx = e1[1] * e2[2] - e1[2] * e2[1]; ny = e1[2] * e2[0] - e1[0] * e2[2]; nz = e1[0] * e2[1] - e1[1] * e2[0];
    const double n2 = nx * nx + ny * ny + nz * nz;
    if (!(n2 > 0.0)) return false;
    const double inv = NMC_RSQRT(n2);
    nx *= inv; ny *= inv; nz *= inv;
  }
  unsigned long long poly = 0x210ULL;  // pool indices, 4 bits per polygon position
  int n = 3, nv = 3;
#pragma unroll 1
  while (true) {
    const unsigned full = (1u << n) - 1u, fullr = full * REP;
    const unsigned tA = A ^ fullr, tB = B ^ fullr;   // set bits = positions outside a plane
    if ((tA | tB) == 0u) break;
    const bool inA = tA != 0u;
    const int bpos = NMC_FFS32(inA ? tA : tB) - 1;
    const int q = (bpos * 205) >> 11;                 // bpos / 10: field of the lowest plane that cuts
    const int p = q + (inA ? 0 : 3);
    const unsigned f = ((inA ? A : B) >> (10 * q)) & full;   // inside mask of plane p over the n positions
    if (f == 0u) return false;
    const int axis = p >> 1;
    const double sgn = (p & 1) ? -1.0 : 1.0, bound = (p & 1) ? H[axis] : 0.0;
    const unsigned prev = ((f << 1) | (f >> (n - 1))) & full;   // position i: vertex i-1 inside
    const int s = NMC_FFS32(f & ~prev) - 1;                      // run start
    const unsigned rot = ((f >> s) | (f << (n - s))) & full;
    const int len = NMC_FFS32(~rot) - 1;                         // run length (the plane cuts: rot != full)
    // all lists rotated so that the run comes first
    const unsigned long long pr = (poly >> (4 * s)) | (poly << (4 * (n - s)));
    const unsigned lowm = ((1u << (n - s)) - 1u) * REP;
    const unsigned Ar = ((A >> s) & lowm) | ((A << (n - s)) & (fullr ^ lowm));
    const unsigned Br = ((B >> s) & lowm) | ((B << (n - s)) & (fullr ^ lowm));
    // planes already applied count as inside for a new vertex, and so does the opposite plane of this axis (the vertex
    // is put ON plane p: coordinate 0 <= H resp. H >= 0)
    const int pf = p | 1;
    const unsigned fm = REP & ((2u << (10 * (pf < 3 ? pf : pf - 3))) - 1u);
    const unsigned forcedA = pf < 3 ? fm : REP, forcedB = pf < 3 ? 0u : fm;
    const unsigned keep = ((1u << len) - 1u) * REP;
    unsigned long long np = 0;
    unsigned nA = 0, nB = 0;
    int m = 0;
    const int ve_p = (int)((poly >> (4 * (s == 0 ? n - 1 : s - 1))) & 15u), ve_q = (int)(pr & 15u);        // outside -> first inside
    const int vl_p = (int)((pr >> (4 * (len - 1))) & 15u), vl_q = (int)((pr >> (4 * len)) & 15u);           // last inside -> outside
    double ep[3], eq[3], lp[3], lq[3];
NMC_UNROLL
    for (int k = 0; k < 3; ++k) { ep[k] = P.get(k, ve_p); eq[k] = P.get(k, ve_q); lp[k] = P.get(k, vl_p); lq[k] = P.get(k, vl_q); }
#if NM_CLIP_AXISLOAD
    // the coordinate the plane acts on, fetched by its (axis, vertex) address instead of selected from the three loaded ones
    const double e_dp = sgn * (P.get(axis, ve_p) - bound), e_dq = sgn * (P.get(axis, ve_q) - bound);
    const double l_dp = sgn * (P.get(axis, vl_p) - bound), l_dq = sgn * (P.get(axis, vl_q) - bound);
#else
    const double e_dp = sgn * ((axis == 0 ? ep[0] : axis == 1 ? ep[1] : ep[2]) - bound);
    const double e_dq = sgn * ((axis == 0 ? eq[0] : axis == 1 ? eq[1] : eq[2]) - bound);
    const double l_dp = sgn * ((axis == 0 ? lp[0] : axis == 1 ? lp[1] : lp[2]) - bound);
    const double l_dq = sgn * ((axis == 0 ? lq[0] : axis == 1 ? lq[1] : lq[2]) - bound);
#endif
    // Pool slot of a new vertex (NM_CLIP_SLOTS).  0: the next unused one (3 + 2 per plane <= 15).  1: the slot of a vertex
    // this plane cuts off -- the entering vertex takes that of the outside end of its edge (ve_p), the leaving one that of
    // vl_q, or a fresh slot when both edges leave from the SAME outside vertex (exactly one vertex cut off: the only case
    // in which the polygon grows; at most six such steps: nine slots).  2: slots 3 + 2p, 4 + 2p of plane p.
    const bool entered = e_dq > 0.0;
    if (entered) {  // entering edge
      const double tt = e_dp * NMC_RCP(e_dp - e_dq);
      double x[3];
NMC_UNROLL
      for (int k = 0; k < 3; ++k) x[k] = ep[k] + tt * (eq[k] - ep[k]);
NMC_UNROLL
#if NM_CLIP_SLOTS == 1
      const int slot = ve_p;
#elif NM_CLIP_SLOTS == 2
      const int slot = 3 + 2 * p;
#else
      const int slot = nv++;
#endif
      for (int k = 0; k < 3; ++k) P.set(k, slot, x[k]);
      P.set(axis, slot, bound);                        // on the plane by construction
      unsigned a, b;
      outcode2(x[0], x[1], x[2], H, a, b);
      np = (unsigned long long)slot;
      nA = a | forcedA; nB = b | forcedB;
      m = 1;
    }
    np |= (pr & ((1ULL << (4 * len)) - 1ULL)) << (4 * m);
    nA |= (Ar & keep) << m;
    nB |= (Br & keep) << m;
    m += len;
    if (l_dp > 0.0) {  // leaving edge
      const double tt = l_dp * NMC_RCP(l_dp - l_dq);
      double x[3];
NMC_UNROLL
      for (int k = 0; k < 3; ++k) x[k] = lp[k] + tt * (lq[k] - lp[k]);
NMC_UNROLL
#if NM_CLIP_SLOTS == 1
      const int slot = (vl_q != ve_p || !entered) ? vl_q : nv++;
#elif NM_CLIP_SLOTS == 2
      const int slot = 4 + 2 * p;
#else
      const int slot = nv++;
#endif
      for (int k = 0; k < 3; ++k) P.set(k, slot, x[k]);
      P.set(axis, slot, bound);
      unsigned a, b;
      outcode2(x[0], x[1], x[2], H, a, b);
      np |= (unsigned long long)slot << (4 * m);
      nA |= (a | forcedA) << m; nB |= (b | forcedB) << m;
      ++m;
    }
    poly = np;
    A = nA; B = nB;
    n = m;
    if (n < 3) return false;
  }
  C.poly = poly; C.n = n; C.nx = nx; C.ny = ny; C.nz = nz;
  return true;
}

// integrals of 1, u, v, w, uv, uw, vw, uvw over the polygon (fan from its first vertex), the reference's thresholds
// applied per fan triangle
template <class Pool> NMC_FN void poly_moments(const ClippedPolyT<Pool> &C, const double *box, double tol, ItemOut &o)
{
  const Pool &P = C.P;
  const unsigned long long poly = C.poly;
  const int n = C.n;
  const double nx = C.nx, ny = C.ny, nz = C.nz;
  const double H[3] = {box[4] - box[0], box[5] - box[1], box[6] - box[2]};
  const double ih[3] = {NMC_RCP(H[0]), NMC_RCP(H[1]), NMC_RCP(H[2])};
  const int v0 = (int)(poly & 15u), v1 = (int)((poly >> 4) & 15u);
  const double ax = P.get(0, v0), ay = P.get(1, v0), az = P.get(2, v0);
  const double f0 = ax * ih[0], g0 = ay * ih[1], h0 = az * ih[2];
  double px = P.get(0, v1), py = P.get(1, v1), pz = P.get(2, v1);
  double f1 = px * ih[0], g1 = py * ih[1], h1 = pz * ih[2];
  for (int j = 2; j < n; ++j) {
    const int vj = (int)((poly >> (4 * j)) & 15u);
    const double qx = P.get(0, vj), qy = P.get(1, vj), qz = P.get(2, vj);
    const double f2 = qx * ih[0], g2 = qy * ih[1], h2 = qz * ih[2];
    const double e1x = px - ax, e1y = py - ay, e1z = pz - az, e2x = qx - ax, e2y = qy - ay, e2z = qz - az;
    const double J = fabs((e1y * e2z - e1z * e2y) * nx + (e1z * e2x - e1x * e2z) * ny + (e1x * e2y - e1y * e2x) * nz);
    if (0.25 * J * J > tol * tol) {                      // squared_area > tol^2     intersections.cc:477,493
      if (J < 1e-12) {                                   // intersections.cc:710
        o.dropped++;
      } else {
        const double Sf = f0 + f1 + f2, Sg = g0 + g1 + g2, Sh = h0 + h1 + h2;
        const double fg = f0 * g0 + f1 * g1 + f2 * g2, fh = f0 * h0 + f1 * h1 + f2 * h2, gh = g0 * h0 + g1 * h1 + g2 * h2;
        const double fgh = f0 * g0 * h0 + f1 * g1 * h1 + f2 * g2 * h2;
        o.area2 += J;
        o.m[0] += J * Sf; o.m[1] += J * Sg; o.m[2] += J * Sh;
        o.m[3] += J * (Sf * Sg + fg); o.m[4] += J * (Sf * Sh + fh); o.m[5] += J * (Sg * Sh + gh);
        o.m[6] += J * (Sf * Sg * Sh + Sf * gh + Sg * fh + Sh * fg + 2.0 * fgh);
        o.nfan++;
      }
    }
    px = qx; py = qy; pz = qz; f1 = f2; g1 = g2; h1 = h2;
  }
}


// ---------------------------------------------------------------------------------------
// Section algorithm: (plane of the triangle  n  box) clipped by the three triangle edges
// ---------------------------------------------------------------------------------------
// The same set as clipping the triangle by the six box planes, built the other way round: the plane of the triangle
// cuts the box in a convex polygon of 3..6 vertices, each ON A BOX EDGE (two coordinates are box bounds exactly, one is
// interpolated); that polygon is then clipped by at most three half-planes (the triangle's edges) instead of the
// triangle by up to six.  Which box edges are cut, and in which cyclic order, depends only on the signs of the plane at
// the eight corners: a 256-entry table (section_lut_build).
//
// LUT entry: bits 0..2 = number of vertices (0 = no polygon), then 4 bits per vertex = box edge id.
// Box edge id = axis * 4 + j; the edge runs along `axis` from corner a(axis, j) (its `axis` bit clear) to a | 1 << axis:
//   axis 0: a = (j & 1) << 1 | (j >> 1) << 2      axis 1: a = (j & 1) | (j >> 1) << 2      axis 2: a = j
NMC_FN int section_edge_corner(int axis, int j)
{
  return axis == 0 ? (((j & 1) << 1) | ((j >> 1) << 2)) : (axis == 1 ? ((j & 1) | ((j >> 1) << 2)) : j);
}

// host side: fills lut[256]
inline void section_lut_build(uint32_t *lut)
{
  for (int mask = 0; mask < 256; ++mask) {
    lut[mask] = 0;
    // crossed edges
    int crossed[12], nc = 0;
    for (int e = 0; e < 12; ++e) {
      const int axis = e >> 2, a = section_edge_corner(axis, e & 3), b = a | (1 << axis);
      crossed[e] = ((mask >> a) & 1) != ((mask >> b) & 1);
      nc += crossed[e];
    }
    if (nc < 3 || nc > 6) continue;
    // faces of an edge: the two coordinates it keeps fixed.  face id = k * 2 + value of coordinate k
    auto on_face = [&](int e, int f) {
      const int axis = e >> 2, a = section_edge_corner(axis, e & 3), k = f >> 1, v = f & 1;
      return k != axis && ((a >> k) & 1) == v;
    };
    bool valid = true;
    for (int f = 0; f < 6 && valid; ++f) {
      int cnt = 0;
      for (int e = 0; e < 12; ++e) cnt += crossed[e] && on_face(e, f);
      valid = cnt == 0 || cnt == 2;
    }
    if (!valid) continue;
    // walk the cycle: consecutive polygon vertices lie on crossed edges of a common face
    int order[6], n = 0, prev = -1, cur = -1;
    for (int e = 0; e < 12; ++e) if (crossed[e]) { cur = e; break; }
    const int start = cur;
    while (n < 6) {
      order[n++] = cur;
      int next = -1;
      for (int f = 0; f < 6 && next < 0; ++f) {
        if (!on_face(cur, f)) continue;
        for (int e = 0; e < 12; ++e)
          if (e != cur && e != prev && crossed[e] && on_face(e, f)) { next = e; break; }
      }
      if (next < 0 || next == start) break;
      prev = cur;
      cur = next;
    }
    if (n != nc) continue;   // not one cycle through all crossed edges: no planar section has this sign pattern
    uint32_t v = (uint32_t)n;
    for (int i = 0; i < n; ++i) v |= (uint32_t)order[i] << (3 + 4 * i);
    lut[mask] = v;
  }
}

// Polygon of one (cell, triangle) item by the section algorithm; same output contract as item_clip_polygon.
NMC_FN bool item_section_polygon(const double *box, const double *quad, int t, const uint32_t *lut, ClippedPoly &C)
{
  VPool &P = C.P;
  const double H[3] = {box[4] - box[0], box[5] - box[1], box[6] - box[2]};
  const int iv[3] = {t == 0 ? 1 : 0, t <= 1 ? 2 : 1, t <= 2 ? 3 : 2};
  double V[3][3];
NMC_UNROLL
  for (int v = 0; v < 3; ++v)
NMC_UNROLL
    for (int k = 0; k < 3; ++k) V[v][k] = quad[iv[v] * 3 + k] - box[k];
  // plane of the triangle: d(x) = n . (x - V0)
  const double e1[3] = {V[1][0] - V[0][0], V[1][1] - V[0][1], V[1][2] - V[0][2]};
  const double e2[3] = {V[2][0] - V[0][0], V[2][1] - V[0][1], V[2][2] - V[0][2]};
  const double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
  const double n2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
  if (!(n2 > 0.0)) return false;
  const double inv_norm = NMC_RSQRT(n2);
  const double d000 = -(n[0] * V[0][0] + n[1] * V[0][1] + n[2] * V[0][2]);
  const double dA[3] = {n[0] * H[0], n[1] * H[1], n[2] * H[2]};   // change of d along a whole box edge of each axis
  // sign pattern at the eight corners (closed box: a corner ON the plane sides with whichever reading gives a polygon)
  unsigned ge = 0, gt = 0;
NMC_UNROLL
  for (int c = 0; c < 8; ++c) {
    const double d = ((d000 + ((c & 1) ? dA[0] : 0.0)) + ((c & 2) ? dA[1] : 0.0)) + ((c & 4) ? dA[2] : 0.0);
    ge |= (d >= 0.0 ? 1u : 0u) << c;
    gt |= (d > 0.0 ? 1u : 0u) << c;
  }
  const unsigned mask = ge == 255u ? gt : ge;
  const uint32_t entry = lut[mask];
  const int cnt = (int)(entry & 7u);
  if (cnt < 3) return false;
  // half-planes of the triangle's edges in its plane: inside iff m_j . x >= c_j, m_j = n x (V_{j+1} - V_j)
  double mj[3][3], cj[3];
NMC_UNROLL
  for (int j = 0; j < 3; ++j) {
    const int j1 = j == 2 ? 0 : j + 1;
    const double f[3] = {V[j1][0] - V[j][0], V[j1][1] - V[j][1], V[j1][2] - V[j][2]};
    mj[j][0] = n[1] * f[2] - n[2] * f[1];
    mj[j][1] = n[2] * f[0] - n[0] * f[2];
    mj[j][2] = n[0] * f[1] - n[1] * f[0];
    cj[j] = mj[j][0] * V[j][0] + mj[j][1] * V[j][1] + mj[j][2] * V[j][2];
  }
  const double rA[3] = {dA[0] != 0.0 ? NMC_RCP(dA[0]) : 0.0, dA[1] != 0.0 ? NMC_RCP(dA[1]) : 0.0, dA[2] != 0.0 ? NMC_RCP(dA[2]) : 0.0};
  unsigned long long poly = 0, code = 0;   // 4-bit pool indices / 6-bit inside codes per position (bits 3..5 always set)
  for (int i = 0; i < cnt; ++i) {
    const int e = (int)((entry >> (3 + 4 * i)) & 15u), axis = e >> 2, a = section_edge_corner(axis, e & 3);
    const double da = ((d000 + ((a & 1) ? dA[0] : 0.0)) + ((a & 2) ? dA[1] : 0.0)) + ((a & 4) ? dA[2] : 0.0);
    double tt = -da * (axis == 0 ? rA[0] : (axis == 1 ? rA[1] : rA[2]));
    tt = tt < 0.0 ? 0.0 : (tt > 1.0 ? 1.0 : tt);
    double x[3] = {(a & 1) ? H[0] : 0.0, (a & 2) ? H[1] : 0.0, (a & 4) ? H[2] : 0.0};
    if (axis == 0) x[0] = tt * H[0]; else if (axis == 1) x[1] = tt * H[1]; else x[2] = tt * H[2];
NMC_UNROLL
    for (int k = 0; k < 3; ++k) P.c[k][i] = x[k];
    unsigned oc = 0x38u;
NMC_UNROLL
    for (int j = 0; j < 3; ++j) oc |= (mj[j][0] * x[0] + mj[j][1] * x[1] + mj[j][2] * x[2] >= cj[j] ? 1u : 0u) << j;
    poly |= (unsigned long long)i << (4 * i);
    code |= (unsigned long long)oc << (6 * i);
  }
  int np_ = cnt, nv = cnt;
  constexpr unsigned long long M6 = 0x0041041041041041ULL;  // bit 0 of every 6-bit group
  // clip by the triangle edges that cut the polygon (the bit-list machinery of item_clip_polygon, general planes)
  while (true) {
    const int nn = np_;
    const unsigned long long cf = code | ~((1ULL << (6 * nn)) - 1ULL);
    const unsigned long long a1 = cf & (cf >> 30);
    const unsigned long long a2 = a1 & (a1 >> 12);
    const unsigned all_in = (unsigned)(a2 & (a2 >> 6) & (a1 >> 24)) & 63u;
    if (all_in == 63u) break;
    const int p = NMC_FFS32(~all_in & 63u) - 1;   // 0..2
    const unsigned long long full6 = M6 & ((1ULL << (6 * nn)) - 1ULL);
    const unsigned long long in6 = (code >> p) & full6;
    if (in6 == 0ULL) return false;
    const unsigned long long prev6 = ((in6 << 6) | (in6 >> (6 * (nn - 1)))) & full6;
    const int s6 = NMC_FFS64(in6 & ~prev6) - 1;
    const int s = NMC_POPC64(M6 & ((1ULL << s6) - 1ULL));
    const unsigned long long rot6 = ((in6 >> s6) | (in6 << (6 * nn - s6))) & full6;
    const int l6 = NMC_FFS64(~rot6 & full6) - 1;
    const int len = NMC_POPC64(M6 & ((1ULL << l6) - 1ULL));
    const unsigned long long pr = (poly >> (4 * s)) | (poly << (4 * (nn - s)));
    const unsigned long long cr = (code >> s6) | (code << (6 * nn - s6));
    const unsigned forced = (2u << p) - 1u;
    unsigned long long npoly = 0, ncode = 0;
    int m = 0;
    const int ve_p = (int)((poly >> (4 * (s == 0 ? nn - 1 : s - 1))) & 15u), ve_q = (int)(pr & 15u);
    const int vl_p = (int)((pr >> (4 * (len - 1))) & 15u), vl_q = (int)((pr >> (4 * len)) & 15u);
    double ep[3], eq[3], lp[3], lq[3];
NMC_UNROLL
    for (int k = 0; k < 3; ++k) { ep[k] = P.c[k][ve_p]; eq[k] = P.c[k][ve_q]; lp[k] = P.c[k][vl_p]; lq[k] = P.c[k][vl_q]; }
    const double m0 = p == 0 ? mj[0][0] : (p == 1 ? mj[1][0] : mj[2][0]), m1 = p == 0 ? mj[0][1] : (p == 1 ? mj[1][1] : mj[2][1]),
                 m2 = p == 0 ? mj[0][2] : (p == 1 ? mj[1][2] : mj[2][2]), cc = p == 0 ? cj[0] : (p == 1 ? cj[1] : cj[2]);
    const double e_dp = m0 * ep[0] + m1 * ep[1] + m2 * ep[2] - cc, e_dq = m0 * eq[0] + m1 * eq[1] + m2 * eq[2] - cc;
    const double l_dp = m0 * lp[0] + m1 * lp[1] + m2 * lp[2] - cc, l_dq = m0 * lq[0] + m1 * lq[1] + m2 * lq[2] - cc;
    if (e_dq > 0.0 && e_dp < 0.0) {  // entering edge
      const double tt = e_dp * NMC_RCP(e_dp - e_dq);
      double x[3];
NMC_UNROLL
      for (int k = 0; k < 3; ++k) x[k] = ep[k] + tt * (eq[k] - ep[k]);
NMC_UNROLL
      for (int k = 0; k < 3; ++k) P.c[k][nv] = x[k];
      unsigned oc = 0x38u | forced;
NMC_UNROLL
      for (int j = 0; j < 3; ++j) oc |= (mj[j][0] * x[0] + mj[j][1] * x[1] + mj[j][2] * x[2] >= cj[j] ? 1u : 0u) << j;
      npoly = (unsigned long long)nv;
      ncode = (unsigned long long)oc;
      m = 1;
      ++nv;
    }
    npoly |= (pr & ((1ULL << (4 * len)) - 1ULL)) << (4 * m);
    ncode |= (cr & ((1ULL << (6 * len)) - 1ULL)) << (6 * m);
    m += len;
    if (l_dp > 0.0 && l_dq < 0.0) {  // leaving edge
      const double tt = l_dp * NMC_RCP(l_dp - l_dq);
      double x[3];
NMC_UNROLL
      for (int k = 0; k < 3; ++k) x[k] = lp[k] + tt * (lq[k] - lp[k]);
NMC_UNROLL
      for (int k = 0; k < 3; ++k) P.c[k][nv] = x[k];
      unsigned oc = 0x38u | forced;
NMC_UNROLL
      for (int j = 0; j < 3; ++j) oc |= (mj[j][0] * x[0] + mj[j][1] * x[1] + mj[j][2] * x[2] >= cj[j] ? 1u : 0u) << j;
      npoly |= (unsigned long long)nv << (4 * m);
      ncode |= (unsigned long long)oc << (6 * m);
      ++m;
      ++nv;
    }
    poly = npoly;
    code = ncode;
    np_ = m;
    if (np_ < 3) return false;
  }
  C.poly = poly; C.n = np_;
  C.nx = n[0] * inv_norm; C.ny = n[1] * inv_norm; C.nz = n[2] * inv_norm;
  return true;
}

// ---------------------------------------------------------------------------------------
// what the kernels call: one switch for the gate, the polygon and (through them) the quadrature emission
// ---------------------------------------------------------------------------------------
// stage-1 gate: can triangle `tri` (cell-local coordinates) meet the box [0, H] at all?  Conservative.
NMC_FN bool tri_gate(const double (*tri)[3], const double *H)
{
#if NM_CLIP_SECTION
  return tri_box_overlap<false>(tri, H);   // box axes + triangle plane; the section polygon settles the rest exactly
#else
  return tri_box_overlap<true>(tri, H);
#endif
}

NMC_FN bool item_polygon(const double *box, const double *quad, int t, const uint32_t *lut, ClippedPoly &C)
{
#if NM_CLIP_SECTION
  return item_section_polygon(box, quad, t, lut, C);
#else
  (void)lut;
#if NM_CLIP_BITS == 2
  return item_clip_polygon_v2(box, quad, t, C);
#else
  return item_clip_polygon_v1(box, quad, t, C);
#endif
#endif
}

NMC_FN void item_moments(const double *box, const double *quad, int t, double tol, const uint32_t *lut, ItemOut &o)
{
  ClippedPoly C;
  if (!item_polygon(box, quad, t, lut, C)) return;
  poly_moments(C, box, tol, o);
}

// the same with the vertex pool supplied by the caller (shared memory in clip_moments3_kernel); plane-major clip only
template <class Pool> NMC_FN void item_moments_pool(const double *box, const double *quad, int t, double tol, ClippedPolyT<Pool> &C, ItemOut &o)
{
  if (!item_clip_polygon_v2(box, quad, t, C)) return;
  poly_moments(C, box, tol, o);
}

// accumulated item sums of one candidate -> its measure and the eight Q1 x P0 local entries
NMC_FN void moments_to_local(const double *tot /* area2, m[7] */, double &sumw, double *l)
{
  sumw = 0.5 * tot[0];
  const double M1 = sumw, Mu = tot[1] * (1.0 / 6.0), Mv = tot[2] * (1.0 / 6.0), Mw = tot[3] * (1.0 / 6.0);
  const double Muv = tot[4] * (1.0 / 24.0), Muw = tot[5] * (1.0 / 24.0), Mvw = tot[6] * (1.0 / 24.0), Muvw = tot[7] * (1.0 / 120.0);
  // phi_i = prod_k (bit_k(i) ? u_k : 1 - u_k)
  l[7] = Muvw; l[3] = Muv - Muvw; l[5] = Muw - Muvw; l[6] = Mvw - Muvw;
  l[1] = Mu - Muv - l[5]; l[2] = Mv - Muv - l[6]; l[4] = Mw - Muw - l[6];
  l[0] = M1 - (l[1] + l[2] + l[3] + l[4] + l[5] + l[6] + l[7]);
}

}  // namespace nmclip
