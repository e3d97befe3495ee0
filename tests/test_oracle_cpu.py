"""CPU tests: the oracles against each other, against the golden fixtures and against the
known-answer identities that pin the path without deal.II/CGAL."""
import os
from fractions import Fraction as Fr

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from non_matching_test_suite_b200.grids import flower_points, make_config
from oracle import exact_oracle as eo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _csr(ref, bg, im):
    return sp.csr_matrix((ref["val"], ref["col"].astype(np.int64), ref["rowptr"].astype(np.int64)), shape=(bg.n_dofs, im.n_dofs))


@pytest.mark.parametrize("name,cycle", [("disk", 0), ("disk", 1), ("flower", 0), ("sphere", 0)])
def test_c_oracle_matches_golden(c_oracle, name, cycle):
    """Golden = exact-rational integrals (tests/golden/make_golden.py); C oracle = FP64 restatement that
    follows the reference's decomposition (tets, diagonal-split quads, 7-point rule)."""
    g = np.load(os.path.join(GOLD, "%s_c%d.npz" % (name, cycle)))
    bg, im = make_config(name, cycle, band_only=False)
    assert (bg.n_cells, im.n_cells, bg.n_dofs) == (int(g["n_bg"]), int(g["n_im"]), int(g["n_space_dofs"]))
    ref = c_oracle.assemble(bg, im, n1d=3, tol=1e-15)
    assert np.array_equal(ref["candidates"], g["candidates"])
    assert np.array_equal(ref["accepted"], g["accepted"])
    assert np.array_equal(ref["rowptr"], g["rowptr"])
    assert np.array_equal(ref["col"], g["col"])
    assert np.linalg.norm(ref["val"] - g["val"]) <= 1e-13 * np.linalg.norm(g["val"])
    # a small clipped piece carries the rounding of its parent simplex: relative to itself it is loose
    assert np.allclose(ref["sum_w"], g["sum_w"], rtol=1e-9, atol=0)
    assert abs(ref["sum_w"].sum() - g["sum_w"].sum()) <= 1e-13 * g["sum_w"].sum()
    if bg.spacedim == 3:
        assert np.array_equal(ref["codes"], g["codes"])


def test_exact_oracle_small_sphere_subset(c_oracle):
    """Exact-rational oracle re-run on a slice of the sphere (the full run is what make_golden.py stores)."""
    bg, im = make_config("sphere", 0, band_only=False)
    ref = c_oracle.assemble(bg, im)
    cand = [(int(k >> 32), int(k & 0xffffffff)) for k in ref["candidates"][:200]]
    r = eo.assemble(bg, im, cand=cand)
    got = {(m, b): s for (m, b), s in zip(r["accepted"], r["sum_w"])}
    for k, s in zip(ref["accepted"], ref["sum_w"]):
        key = (int(k >> 32), int(k & 0xffffffff))
        if key in got:
            assert abs(got[key] - s) <= 1e-13 * s
    assert len(got) > 50


@pytest.mark.parametrize("name,cycle", [("disk", 0), ("disk", 2), ("flower", 1), ("sphere", 1)])
def test_known_answer_identities(c_oracle, name, cycle):
    bg, im = make_config(name, cycle, band_only=False)
    ref = c_oracle.assemble(bg, im)
    A = _csr(ref, bg, im)
    im_cells = (ref["accepted"] >> np.uint64(32)).astype(np.int64)
    per_cell = np.bincount(im_cells, weights=ref["sum_w"], minlength=im.n_cells)
    # 1. sum of weights = immersed measure (per cell), independent of the background mesh
    if bg.spacedim == 2:
        measure = im.measure().numpy()
    else:
        measure = np.zeros(im.n_cells)
        v = im.verts
        for m, c in enumerate(ref["codes"]):
            for (i, j, k) in eo.split_triangles(int(c)):
                measure[m] += 0.5 * float(torch.linalg.cross(v[m, j] - v[m, i], v[m, k] - v[m, i]).norm())
    assert np.allclose(per_cell, measure, rtol=1e-11, atol=0)
    # 2. partition of unity: hanging-node weights sum to one, Dirichlet rows never touch a cut cell
    assert np.allclose(np.asarray(A.sum(axis=0)).ravel(), per_cell, rtol=1e-11, atol=1e-15)
    # 3. linear reproduction: sum_i x_i A[i,j] = int_{K_j} x   (Q1 reproduces linears, also through the lines)
    X = bg.dof_coords.numpy()
    lhs = A.T @ X
    nl = 1 << bg.spacedim
    local = ref["local"]
    dofs = bg.dofs.numpy().astype(np.int64)[(ref["accepted"] & np.uint64(0xffffffff)).astype(np.int64)]
    rhs = np.zeros_like(lhs)
    for k in range(bg.spacedim):
        np.add.at(rhs[:, k], im_cells, (local * X[dofs, k]).sum(axis=1))
    assert np.allclose(lhs, rhs, rtol=1e-10, atol=1e-14)
    # constrained rows are in the pattern and structurally zero
    cd = bg.c_dof.numpy().astype(np.int64)
    assert abs(A[cd]).sum() == 0.0
    if (name, cycle) == ("disk", 0):     # hanging nodes sit on cut cells at the first cycle (hard part 4)
        assert np.diff(ref["rowptr"].astype(np.int64))[cd].sum() > 0


def test_refinement_invariance(c_oracle):
    """C on a background mesh and on its uniform refinement give the same integrals of every coarse shape function."""
    bg0, im = make_config("disk", 1, band_only=False)
    bg1, _ = make_config("disk", 2, band_only=False)
    from non_matching_test_suite_b200.grids import circle_grid
    im = circle_grid(64)
    A0 = _csr(c_oracle.assemble(bg0, im), bg0, im)
    A1 = _csr(c_oracle.assemble(bg1, im), bg1, im)
    # columns sums with x-weights agree (moments of the immersed cells do not depend on the background)
    for k in range(2):
        m0 = A0.T @ bg0.dof_coords.numpy()[:, k]
        m1 = A1.T @ bg1.dof_coords.numpy()[:, k]
        assert np.allclose(m0, m1, rtol=1e-11, atol=1e-15)


def test_candidates_equal_brute_force(c_oracle):
    for name in ("disk", "sphere"):
        bg, im = make_config(name, 0, band_only=False)
        ilo, ihi = im.boxes()
        keys = c_oracle.candidates(bg.lo, bg.hi, ilo, ihi)
        brute = eo.candidates(bg.lo.tolist(), bg.hi.tolist(), ilo.tolist(), ihi.tolist())
        assert np.array_equal(keys, np.array([(m << 32) | b for m, b in brute], dtype=np.uint64))


def test_exact_predicates_on_near_degenerate_input(c_oracle):
    """C oracle's filtered/expansion predicates against exact rationals, including exactly cocircular
    and 1-ulp perturbed configurations."""
    rng = np.random.default_rng(0)
    n_zero = 0
    for trial in range(300):
        c = rng.random(2)
        r = 0.1 + rng.random()
        th = np.sort(rng.random(4) * 2 * np.pi)
        P = np.stack([c[0] + r * np.cos(th), c[1] + r * np.sin(th)], axis=1)
        if trial % 3 == 0:      # exactly cocircular lattice points
            P = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]]) * (1 + trial) / 64 + 0.125
        if trial % 3 == 1:      # perturb one coordinate by one ulp
            P[3, 0] = np.nextafter(P[3, 0], 2.0)
        a, b, c_, d = P
        ex = eo.incircle(*[(Fr(float(p[0])), Fr(float(p[1]))) for p in (a, b, c_, d)])
        want = (ex > 0) - (ex < 0)
        assert c_oracle.incircle(a, b, c_, d) == want
        n_zero += want == 0
        eo2 = eo.orient2d(*[(Fr(float(p[0])), Fr(float(p[1]))) for p in (a, b, c_)])
        assert c_oracle.orient2d(a, b, c_) == (eo2 > 0) - (eo2 < 0)
    assert n_zero >= 50
    # collinear triples
    for t in (0.25, 1.0 / 3.0, 0.7):
        a, b = np.array([0.1, 0.2]), np.array([0.9, 0.6])
        c_ = a + t * (b - a)
        ex = eo.orient2d(*[(Fr(float(p[0])), Fr(float(p[1]))) for p in (a, b, c_)])
        assert c_oracle.orient2d(a, b, c_) == (ex > 0) - (ex < 0)


def test_quad_split_codes_match_exact(c_oracle):
    bg, im = make_config("sphere", 1, band_only=False)
    codes = c_oracle.quad_split(im.verts)
    V = im.verts.tolist()
    for m in range(0, im.n_cells, 3):
        assert int(codes[m]) == eo.quad_split_code(V[m])
    # degenerate inputs: square (cocircular tie), three collinear, coincident projection, interior point
    sq = [[0, 0, 0], [1, 0, 0.3], [0, 1, 0.1], [1, 1, 0]]
    col = [[0, 0, 0], [1, 0, 0], [2, 0, 1], [0.5, 1, 0]]
    dup = [[0, 0, 0], [0, 0, 1], [1, 0, 0], [0, 1, 0]]
    inner = [[0, 0, 0], [4, 0, 0], [0, 4, 0], [1, 1, 2]]
    for q in (sq, col, dup, inner):
        assert int(c_oracle.quad_split(torch.tensor([q], dtype=torch.float64))[0]) == eo.quad_split_code(q)
    assert eo.quad_split_code(sq) & eo.TIE
    assert bin(eo.quad_split_code(inner) & 15).count("1") == 3


def test_unsupported_rule_is_an_error(c_oracle):
    bg, im = make_config("disk", 0, band_only=False)
    with pytest.raises(ValueError):
        c_oracle.assemble(bg, im, n1d=9)


def test_flower_points_match_reference_file():
    path = "/root/reference/code/grids/flower_interface.vtk"
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    lines = open(path).read().split("\n")
    i = [k for k, l in enumerate(lines) if l.startswith("POINTS")][0]
    P = np.array([[float(x) for x in lines[i + 1 + k].split()[:2]] for k in range(256)])
    assert np.array_equal(P, flower_points().numpy())


@pytest.mark.parametrize("d", [2, 3])
def test_degenerate_touching_cases(c_oracle, d):
    """Rules R3/R5 on hand-made cases: the FP64 oracle agrees with exact rational arithmetic about which pairs
    exist, and about the double counting of simplices lying in a shared face."""
    from degenerate_cases import case_2d, case_3d
    bg, im = case_2d() if d == 2 else case_3d()
    ex = eo.assemble(bg, im, tol=1e-15)
    ref = c_oracle.assemble(bg, im, n1d=3, tol=1e-15)
    ea = np.array([(m << 32) | b for m, b in ex["accepted"]], dtype=np.uint64)
    assert np.array_equal(ea, ref["accepted"])
    assert np.allclose(ex["sum_w"], ref["sum_w"], rtol=1e-12, atol=1e-15)
    per_cell = np.bincount((ref["accepted"] >> np.uint64(32)).astype(np.int64), weights=ref["sum_w"], minlength=im.n_cells)
    if d == 2:
        meas = im.measure().numpy()
        assert abs(per_cell[0] - 2 * meas[0]) < 1e-14          # in a shared edge: counted for both neighbours
        assert abs(per_cell[3] - 2 * meas[3]) < 1e-14          # along a grid line
        assert abs(per_cell[1] - meas[1]) < 1e-14 and abs(per_cell[4] - meas[4]) < 1e-14
        assert abs(per_cell[5] - meas[5]) < 1e-14              # on the domain boundary: one neighbour only
    else:
        assert abs(per_cell[0] - 2 * 0.36) < 1e-14             # quad in the shared face z = 1
        assert abs(per_cell[2] - 4.0) < 1e-13                  # flat quad of area 4, interior of the cells
        assert per_cell[3] == 0.0                              # vertical quad: its xy-projection is a line, no Delaunay face
        assert ref["codes"][3] & 0x0f == 0


def test_random_quads_and_boxes_c_oracle_against_exact(c_oracle):
    """Pair level, random general-position data: skewed (non-planar) quads of random size against random anisotropic
    boxes -- the FP64 C restatement (tetrahedra x Delaunay triangles, 7-point rule) against exact rational clipping and
    integration, including the split code and the local Q1 vector."""
    rng = np.random.default_rng(20240607)
    n = 60
    lo = rng.uniform(-1, 1, (n, 3)); hi = lo + rng.uniform(0.05, 1.0, (n, 3))
    quads = np.zeros((n, 4, 3))
    for i in range(n):
        c = 0.5 * (lo[i] + hi[i]) + rng.uniform(-0.3, 0.3, 3) * (hi[i] - lo[i])
        e1, e2 = rng.standard_normal(3), rng.standard_normal(3)
        s = rng.uniform(0.2, 1.5) * (hi[i] - lo[i]).max()
        for v, (a, b) in enumerate(((-1, -1), (1, -1), (-1, 1), (1, 1))):      # deal.II vertex order
            quads[i, v] = c + 0.5 * s * (a * e1 + b * e2) + 0.1 * s * rng.standard_normal(3)
    codes = c_oracle.quad_split(quads)
    cand = (np.arange(n, dtype=np.uint64) << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    sumw, local, nq, dropped = c_oracle.pairs(3, cand, lo, hi, quads, codes, 3, 1e-15)
    hits = 0
    for i in range(n):
        assert eo.quad_split_code(quads[i]) == int(codes[i])
        s_ex, l_ex = eo.pair_integrals_3d(lo[i], hi[i], quads[i])
        assert abs(sumw[i] - s_ex) <= 1e-13 * max(s_ex, 1e-3)
        assert np.abs(local[i] - np.array(l_ex)).max() <= 1e-13 * max(s_ex, 1e-3)
        hits += s_ex > 0
    assert hits > n // 2
