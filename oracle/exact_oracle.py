"""TEST INFRASTRUCTURE ONLY -- exact-rational oracle for the coupling path.

PARITY UNPINNED: the reference (deal.II 9.4.2 + CGAL 5.5.2) cannot be built or run in
this environment and ships no tests, golden vectors or stored tables for this path
(SURVEY.md section 4 / 8(c)).  This file restates the *mathematics* the reference
computes; it is pinned only by the known-answer identities in tests/ (sum of weights =
immersed measure, partition of unity, linear reproduction, refinement invariance).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

What is restated (file:line relative to /root/reference/code):
  * candidate pairs: closed bounding-box overlap, Boost `bgi::intersects`
        src/coupling_utilities.cc:53-55
  * quad -> triangles: finite faces of the Delaunay triangulation of the four vertices
    projected to xy (`Projection_traits_xy_3`)           src/intersections.cc:83-84,464
  * exact intersection of each immersed simplex with the (axis-aligned) background cell;
    lower-dimensional results are discarded               src/intersections.cc:358-393,437-509
  * integrals  local(i) = int phi_i * 1  over the intersection (what the quadrature of
    src/intersections.cc:693-746 evaluates exactly for Q1 x P0, rule R1 of the survey)
        include/coupling_utilities.h:258-274
  * acceptance  sum of weights > tol                     src/coupling_utilities.cc:59-65
  * sparsity with keep_constrained_entries = true        include/coupling_utilities.h:127-129
  * scatter through constraint lines                     include/coupling_utilities.h:285-287

All geometry is done with `fractions.Fraction` on the exact values of the input doubles;
the single irrational factor (segment length / triangle normal length) is taken in
50-digit `mpmath` arithmetic and rounded to double at the very end.
"""
from __future__ import annotations

from fractions import Fraction as Fr
from typing import Dict, List, Sequence, Tuple

import mpmath

mpmath.mp.dps = 50

TRIPLES = ((1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2))  # triple t omits vertex t
TIE = 0x80


# --------------------------------------------------------------------------- #
# exact predicates and the quad split (rule R2)
# --------------------------------------------------------------------------- #
def orient2d(a, b, c) -> Fr:
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])


def incircle(a, b, c, d) -> Fr:
    """> 0 iff d lies strictly inside the circle through a,b,c given counter-clockwise."""
    rows = []
    for p in (a, b, c):
        x, y = p[0] - d[0], p[1] - d[1]
        rows.append((x, y, x * x + y * y))
    (a0, a1, a2), (b0, b1, b2), (c0, c1, c2) = rows
    return a0 * (b1 * c2 - b2 * c1) - a1 * (b0 * c2 - b2 * c0) + a2 * (b0 * c1 - b1 * c0)


def quad_split_code(quad) -> int:
    """Bit t set <=> triangle TRIPLES[t] is a finite Delaunay face of the xy-projection.
    Bit 7 (TIE) is set when the answer depends on CGAL's insertion order: four cocircular
    points or points that coincide in projection (a deterministic choice is made)."""
    P = [(Fr(float(v[0])), Fr(float(v[1]))) for v in quad]
    # coincident projections: CGAL keeps one of them, which one depends on spatial_sort
    for i in range(4):
        for j in range(i + 1, 4):
            if P[i] == P[j]:
                keep = [k for k in range(4) if k != j]
                # if more coincide, or the rest is collinear, nothing is left
                a, b, c = (P[k] for k in keep)
                if a == b or b == c or a == c or orient2d(a, b, c) == 0:
                    return TIE
                return TIE | (1 << j)
    code = 0
    tie = False
    for t, (i, j, k) in enumerate(TRIPLES):
        o = orient2d(P[i], P[j], P[k])
        if o == 0:
            continue
        ic = incircle(P[i], P[j], P[k], P[t])
        if o < 0:
            ic = -ic
        if ic < 0:
            code |= 1 << t
        elif ic == 0:
            tie = True
    if tie:
        # cocircular: lexicographic convention = diagonal (v0, v3): triangles (0,1,3), (0,2,3)
        return TIE | (1 << 2) | (1 << 1)
    return code


def split_triangles(code: int) -> List[Tuple[int, int, int]]:
    return [TRIPLES[t] for t in range(4) if (code >> t) & 1]


# --------------------------------------------------------------------------- #
# exact clipping
# --------------------------------------------------------------------------- #
def _clip_poly_halfspace(poly, axis, bound, keep_greater):
    """Sutherland-Hodgman step in exact arithmetic (closed half-space)."""
    out = []
    n = len(poly)
    for i in range(n):
        p, q = poly[i], poly[(i + 1) % n]
        dp = (p[axis] - bound) if keep_greater else (bound - p[axis])
        dq = (q[axis] - bound) if keep_greater else (bound - q[axis])
        if dp >= 0:
            out.append(p)
        if (dp > 0 and dq < 0) or (dp < 0 and dq > 0):
            t = dp / (dp - dq)
            out.append(tuple(p[k] + t * (q[k] - p[k]) for k in range(len(p))))
    return out


def clip_triangle_box(tri, lo, hi):
    poly = [tuple(v) for v in tri]
    for axis in range(3):
        poly = _clip_poly_halfspace(poly, axis, lo[axis], True)
        if len(poly) < 3:
            return []
        poly = _clip_poly_halfspace(poly, axis, hi[axis], False)
        if len(poly) < 3:
            return []
    return poly


def clip_segment_box(p0, p1, lo, hi):
    """Liang-Barsky in exact arithmetic; returns (t0, t1) or None (also None for a point)."""
    t0, t1 = Fr(0), Fr(1)
    for k in range(2):
        d = p1[k] - p0[k]
        if d == 0:
            if p0[k] < lo[k] or p0[k] > hi[k]:
                return None
        else:
            ta, tb = (lo[k] - p0[k]) / d, (hi[k] - p0[k]) / d
            if ta > tb:
                ta, tb = tb, ta
            t0, t1 = max(t0, ta), min(t1, tb)
            if t0 > t1:
                return None
    if t0 >= t1:
        return None
    return t0, t1


# --------------------------------------------------------------------------- #
# exact integration of the Q1 shape functions
# --------------------------------------------------------------------------- #
def _shape(i, xi):
    v = Fr(1)
    for k, x in enumerate(xi):
        v *= x if (i >> k) & 1 else (1 - x)
    return v


# degree-3 exact rules with rational nodes (integrands here have degree <= 3)
_SIMPSON = ((Fr(0), Fr(1, 6)), (Fr(1, 2), Fr(4, 6)), (Fr(1), Fr(1, 6)))
_TRI4 = (((Fr(1, 3), Fr(1, 3)), Fr(-27, 96)), ((Fr(1, 5), Fr(1, 5)), Fr(25, 96)),
         ((Fr(3, 5), Fr(1, 5)), Fr(25, 96)), ((Fr(1, 5), Fr(3, 5)), Fr(25, 96)))


def pair_integrals_2d(lo, hi, seg):
    """(sum_w, local[4]) for a background rectangle and an immersed segment; doubles in, doubles out."""
    lo = [Fr(float(x)) for x in lo]; hi = [Fr(float(x)) for x in hi]
    p0 = [Fr(float(x)) for x in seg[0]]; p1 = [Fr(float(x)) for x in seg[1]]
    clip = clip_segment_box(p0, p1, lo, hi)
    if clip is None:
        return 0.0, [0.0] * 4
    t0, t1 = clip
    d = [p1[k] - p0[k] for k in range(2)]
    length = mpmath.sqrt(mpmath.mpf(d[0].numerator) / d[0].denominator * mpmath.mpf(d[0].numerator) / d[0].denominator
                         + mpmath.mpf(d[1].numerator) / d[1].denominator * mpmath.mpf(d[1].numerator) / d[1].denominator)
    h = [hi[k] - lo[k] for k in range(2)]
    acc = [Fr(0)] * 4
    for s, w in _SIMPSON:
        t = t0 + s * (t1 - t0)
        xi = [(p0[k] + t * d[k] - lo[k]) / h[k] for k in range(2)]
        for i in range(4):
            acc[i] += w * _shape(i, xi)
    scale = length * mpmath.mpf((t1 - t0).numerator) / (t1 - t0).denominator
    local = [float(scale * mpmath.mpf(a.numerator) / a.denominator) for a in acc]
    return float(scale), local


def pair_integrals_2d2d(lo, hi, quad):
    """(sum_w, local[4]) for a background rectangle and a convex immersed quad in its plane (the reference's <2,2,2>
    instantiation, intersections.cc:307-355): quad clipped by the four lines of the rectangle, then the bilinear shape
    functions integrated exactly over the fan triangles with a rational rule exact to degree 3; everything in rationals,
    the result rounded once."""
    lo = [Fr(float(x)) for x in lo]; hi = [Fr(float(x)) for x in hi]
    poly = [tuple(Fr(float(c)) for c in quad[i]) for i in (0, 1, 3, 2)]        # deal.II vertex order -> around the cell
    for axis in range(2):
        poly = _clip_poly_halfspace(poly, axis, lo[axis], True)
        if len(poly) < 3:
            return 0.0, [0.0] * 4
        poly = _clip_poly_halfspace(poly, axis, hi[axis], False)
        if len(poly) < 3:
            return 0.0, [0.0] * 4
    h = [hi[k] - lo[k] for k in range(2)]
    tot, acc = Fr(0), [Fr(0)] * 4
    a = poly[0]
    for j in range(1, len(poly) - 1):
        b, c = poly[j], poly[j + 1]
        J = abs((b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]))       # 2 * area, exact
        if J == 0:
            continue
        tot += J / 2
        for (s, t), w in _TRI4:                                                     # weights sum to 1/2: w * J = weight on the triangle
            x = [a[k] + s * (b[k] - a[k]) + t * (c[k] - a[k]) for k in range(2)]
            xi = [(x[k] - lo[k]) / h[k] for k in range(2)]
            for i in range(4):
                acc[i] += w * J * _shape(i, xi)
    return float(tot), [float(v) for v in acc]


# Boole's rule: rational nodes, exact to degree 5 (phi_i phi_j has degree 4 along a segment)
_BOOLE = ((Fr(0), Fr(7, 90)), (Fr(1, 4), Fr(32, 90)), (Fr(1, 2), Fr(12, 90)), (Fr(3, 4), Fr(32, 90)), (Fr(1), Fr(7, 90)))


def nitsche_block_2d(lo, hi, seg):
    """Exact  int phi_i phi_j ds  over (segment n rectangle), 4x4, doubles in / doubles out: the local block of
    assemble_nitsche_with_exact_intersections (code/include/coupling_utilities.h:366-370) for coefficient = 1 and
    penalty / h = 1, i.e. what the reference's 3-point rule evaluates exactly on each of its sub-segments."""
    lo = [Fr(float(x)) for x in lo]; hi = [Fr(float(x)) for x in hi]
    p0 = [Fr(float(x)) for x in seg[0]]; p1 = [Fr(float(x)) for x in seg[1]]
    clip = clip_segment_box(p0, p1, lo, hi)
    if clip is None:
        return [[0.0] * 4 for _ in range(4)]
    t0, t1 = clip
    d = [p1[k] - p0[k] for k in range(2)]
    length = mpmath.sqrt(_mpf(d[0]) * _mpf(d[0]) + _mpf(d[1]) * _mpf(d[1]))
    h = [hi[k] - lo[k] for k in range(2)]
    acc = [[Fr(0)] * 4 for _ in range(4)]
    for s, w in _BOOLE:
        t = t0 + s * (t1 - t0)
        xi = [(p0[k] + t * d[k] - lo[k]) / h[k] for k in range(2)]
        phi = [_shape(i, xi) for i in range(4)]
        for i in range(4):
            for j in range(4):
                acc[i][j] += w * phi[i] * phi[j]
    scale = length * _mpf(t1 - t0)
    return [[float(scale * _mpf(a)) for a in row] for row in acc]


def _cross(u, v):
    return (u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0])


def _mpf(fr: Fr):
    return mpmath.mpf(fr.numerator) / fr.denominator


def triangle_integrals_3d(lo, hi, tri):
    """Exact (as mpf) measure and int phi_i over (triangle ∩ box)."""
    poly = clip_triangle_box(tri, lo, hi)
    if len(poly) < 3:
        return mpmath.mpf(0), [mpmath.mpf(0)] * 8
    e1 = [tri[1][k] - tri[0][k] for k in range(3)]
    e2 = [tri[2][k] - tri[0][k] for k in range(3)]
    N = _cross(e1, e2)
    n2 = N[0] * N[0] + N[1] * N[1] + N[2] * N[2]
    if n2 == 0:
        return mpmath.mpf(0), [mpmath.mpf(0)] * 8
    ax = max(range(3), key=lambda k: abs(N[k]))
    nlen = mpmath.sqrt(_mpf(n2))
    h = [hi[k] - lo[k] for k in range(3)]
    area_r = Fr(0)            # rational: sum over fan of (n_fan[ax] / N[ax]) ; times |N| = 2 * area
    acc = [Fr(0)] * 8
    for j in range(1, len(poly) - 1):
        a, b, c = poly[0], poly[j], poly[j + 1]
        f1 = [b[k] - a[k] for k in range(3)]
        f2 = [c[k] - a[k] for k in range(3)]
        ratio = _cross(f1, f2)[ax] / N[ax]          # (2*area_fan) / |N|, same orientation => >= 0
        if ratio == 0:
            continue
        area_r += ratio
        for (s, t), w in _TRI4:
            x = [a[k] + s * f1[k] + t * f2[k] for k in range(3)]
            xi = [(x[k] - lo[k]) / h[k] for k in range(3)]
            for i in range(8):
                acc[i] += ratio * w * _shape(i, xi)
    # weights of _TRI4 sum to 1/2 => integral = |N| * ratio * sum(w f); measure = |N| * ratio / 2
    return nlen * _mpf(area_r) / 2, [nlen * _mpf(a) for a in acc]


def pair_integrals_3d(lo, hi, quad, code=None):
    """(sum_w, local[8]) for a background box and an immersed quad (deal.II vertex order)."""
    lo = [Fr(float(x)) for x in lo]; hi = [Fr(float(x)) for x in hi]
    Q = [[Fr(float(x)) for x in v] for v in quad]
    if code is None:
        code = quad_split_code(quad)
    tot = mpmath.mpf(0)
    loc = [mpmath.mpf(0)] * 8
    for (i, j, k) in split_triangles(code):
        m, l = triangle_integrals_3d(lo, hi, (Q[i], Q[j], Q[k]))
        tot += m
        loc = [loc[q] + l[q] for q in range(8)]
    return float(tot), [float(v) for v in loc]


# --------------------------------------------------------------------------- #
# whole-path oracle on small grids
# --------------------------------------------------------------------------- #
def candidates(bg_lo, bg_hi, im_lo, im_hi) -> List[Tuple[int, int]]:
    """Brute-force closed-box overlap on raw doubles; sorted by (immersed, background)."""
    out = []
    d = len(bg_lo[0])
    for m in range(len(im_lo)):
        for b in range(len(bg_lo)):
            if all(bg_lo[b][k] <= im_hi[m][k] and im_lo[m][k] <= bg_hi[b][k] for k in range(d)):
                out.append((m, b))
    return out


def assemble(bg, im, tol: float = 1e-15, cand=None):
    """Exact restatement of the three reference entry points on a (small) grid pair.

    bg, im: objects with the fields of grids.BackgroundGrid / grids.ImmersedGrid (CPU tensors).
    Returns dict with candidates, accepted pairs (+ sum_w), the sparsity set of the stored
    n_space x n_immersed matrix and its values {(row, col): value}.
    """
    d = bg.spacedim
    lo, hi = bg.lo.tolist(), bg.hi.tolist()
    V = im.verts.tolist()
    ilo, ihi = (t.tolist() for t in im.boxes())
    bdofs = bg.dofs.tolist()
    idofs = im.dofs.tolist()
    lines = constraint_lines(bg)
    if cand is None:
        cand = candidates(lo, hi, ilo, ihi)
    codes = [quad_split_code(V[m]) for m in range(len(V))] if d == 3 else None
    accepted, sumw, pattern, values = [], [], set(), {}
    for (m, b) in cand:
        if d == 2 and im.dim == 2:
            s, loc = pair_integrals_2d2d(lo[b], hi[b], V[m])
        elif d == 2:
            s, loc = pair_integrals_2d(lo[b], hi[b], V[m])
        else:
            s, loc = pair_integrals_3d(lo[b], hi[b], V[m], codes[m])
        if not s > tol:
            continue
        accepted.append((m, b)); sumw.append(s)
        col = idofs[m][0]
        for i, r in enumerate(bdofs[b]):
            if r in lines:
                pattern.add((r, col))                 # keep_constrained_entries
                for (mm, w) in lines[r]:
                    pattern.add((mm, col))
                    if loc[i] != 0.0:
                        values[(mm, col)] = values.get((mm, col), 0.0) + w * loc[i]
            else:
                pattern.add((r, col))
                if loc[i] != 0.0:
                    values[(r, col)] = values.get((r, col), 0.0) + loc[i]
    return dict(candidates=cand, accepted=accepted, sum_w=sumw, pattern=pattern, values=values, codes=codes)


def constraint_lines(bg) -> Dict[int, List[Tuple[int, float]]]:
    cd, cp = bg.c_dof.tolist(), bg.c_ptr.tolist()
    cm, cw = bg.c_master.tolist(), bg.c_weight.tolist()
    return {cd[i]: [(cm[k], cw[k]) for k in range(cp[i], cp[i + 1])] for i in range(len(cd))}
