"""Pins oracle/fe_oracle.py (general tensor-product Lagrange pair on the coupling path) without a GPU: it reduces to the
Q1 x P0 oracle, satisfies partition of unity, does not depend on the order the support points are given in, and in 2D
the rules the drivers use (n = 2p + 1 points, 2d3d/lagrange_multiplier.cc:750-753) make the matrix independent of how
the intersections are subdivided."""
import numpy as np
import pytest

from non_matching_test_suite_b200.grids import make_config, circle_grid, sphere_grid
from oracle import fe_oracle as fo
from fe_helpers import uniform_background, continuous_dofs, discontinuous_dofs


def _pairs(acc):
    return (acc >> np.uint64(32)).astype(np.int64), (acc & np.uint64(0xffffffff)).astype(np.int64)


@pytest.mark.parametrize("name,cycle", [("disk", 1), ("sphere", 0)])
def test_q1_p0_reduces_to_the_base_oracle(c_oracle, name, cycle):
    bg, im = make_config(name, cycle, band_only=False)
    r = c_oracle.assemble_with_quadrature(bg, im, 3, 1e-15)
    pim, pbg = _pairs(r["accepted"])
    d = bg.spacedim
    A = fo.coupling_matrix(bg.lo.numpy(), bg.hi.numpy(), bg.dofs.numpy(), bg.n_dofs, bg.c_dof.numpy(), bg.c_ptr.numpy(), bg.c_master.numpy(),
                           bg.c_weight.numpy(), im.verts.numpy(), im.dofs.numpy(), im.n_dofs, pim, pbg, r["q_offset"], r["points"], r["weights"],
                           fo.unit_support_points(d, 1), fo.unit_support_points(d - 1, 0))
    assert np.array_equal(A.indptr.astype(np.uint64), r["rowptr"]) and np.array_equal(A.indices.astype(np.uint32), r["col"])
    assert np.linalg.norm(A.data - r["val"]) <= 1e-13 * np.linalg.norm(r["val"])


@pytest.mark.parametrize("d,p,q", [(2, 2, 1), (2, 3, 2), (3, 2, 1)])
def test_partition_of_unity_and_dof_order(c_oracle, d, p, q):
    from non_matching_test_suite_b200.grids import BackgroundGrid
    import torch
    lo, hi = uniform_background(d, 8 if d == 2 else 6)
    im = circle_grid(24) if d == 2 else sphere_grid(3)
    e = torch.empty(0, dtype=torch.int32)
    bg = BackgroundGrid(d, torch.from_numpy(lo), torch.from_numpy(hi), torch.zeros(lo.shape[0], 1 << d, dtype=torch.int32), 1,
                        torch.zeros(lo.shape[0], dtype=torch.int32), e, torch.zeros(1, dtype=torch.int64), e, torch.empty(0, dtype=torch.float64))
    n1d = min(2 * p + 1, 8) if d == 2 else 3
    r = c_oracle.assemble_with_quadrature(bg, im, n1d, 1e-15)
    pim, pbg = _pairs(r["accepted"])
    none = (np.empty(0, np.uint32), np.zeros(1, np.uint64), np.empty(0, np.uint32), np.empty(0))
    mats = []
    for order in ("lexicographic", "shuffled"):
        sb, si = fo.unit_support_points(d, p, order, 1), fo.unit_support_points(d - 1, q, order, 2)
        bd, ns = continuous_dofs(lo, hi, sb)
        idf, ni = discontinuous_dofs(im.n_cells, si.shape[0])
        A = fo.coupling_matrix(lo, hi, bd, ns, *none, im.verts.numpy(), idf, ni, pim, pbg, r["q_offset"], r["points"], r["weights"], sb, si)
        mats.append((A, bd, idf, sb, si))
    A = mats[0][0]
    # partition of unity of both bases: the sum of all entries is the immersed measure
    assert abs(A.sum() - r["sum_w"].sum()) <= 1e-12 * r["sum_w"].sum()
    # the two dof orders describe the same operator: compare through the physical support points
    (A0, bd0, id0, sb0, si0), (A1, bd1, id1, sb1, si1) = mats
    perm_b = np.array([int(np.argmin(np.abs(sb1 - s).sum(1))) for s in sb0])
    perm_i = np.array([int(np.argmin(np.abs(si1 - s).sum(1))) for s in si0]) if q else np.zeros(1, dtype=int)
    row_map = np.zeros(A0.shape[0], dtype=np.int64); row_map[bd0.reshape(-1)] = bd1[:, perm_b].reshape(-1)
    col_map = np.zeros(A0.shape[1], dtype=np.int64); col_map[id0.reshape(-1)] = id1[:, perm_i].reshape(-1)
    C0 = A0.tocoo()
    B = A1.tocsr()[row_map[C0.row], col_map[C0.col]]
    assert np.abs(np.asarray(B).ravel() - C0.data).max() <= 1e-14


@pytest.mark.parametrize("p,q", [(1, 1), (2, 0), (2, 2), (3, 1)])
def test_2d_rules_make_the_matrix_subdivision_independent(c_oracle, p, q):
    """Reference subdivision (segments cut at the CDT diagonal of the background quad, oracle) against ONE sub-segment per
    pair (Liang-Barsky, numpy): the integrand has degree 2p + q <= 2 (2p + 1) - 1 along a straight segment."""
    from non_matching_test_suite_b200.grids import BackgroundGrid
    import torch
    lo, hi = uniform_background(2, 10)
    im = circle_grid(40)
    e = torch.empty(0, dtype=torch.int32)
    bg = BackgroundGrid(2, torch.from_numpy(lo), torch.from_numpy(hi), torch.zeros(lo.shape[0], 4, dtype=torch.int32), 1,
                        torch.zeros(lo.shape[0], dtype=torch.int32), e, torch.zeros(1, dtype=torch.int64), e, torch.empty(0, dtype=torch.float64))
    n1d = 2 * p + 1
    r = c_oracle.assemble_with_quadrature(bg, im, n1d, 1e-15)
    pim, pbg = _pairs(r["accepted"])
    sb, si = fo.unit_support_points(2, p), fo.unit_support_points(1, q)
    bd, ns = continuous_dofs(lo, hi, sb)
    idf, ni = discontinuous_dofs(im.n_cells, si.shape[0])
    none = (np.empty(0, np.uint32), np.zeros(1, np.uint64), np.empty(0, np.uint32), np.empty(0))
    A = fo.coupling_matrix(lo, hi, bd, ns, *none, im.verts.numpy(), idf, ni, pim, pbg, r["q_offset"], r["points"], r["weights"], sb, si)
    # one clipped sub-segment per pair
    gx, gw = np.polynomial.legendre.leggauss(n1d)
    gx, gw = 0.5 + 0.5 * gx, 0.5 * gw
    V = im.verts.numpy()
    off, pts, wts = [0], [], []
    for m, b in zip(pim, pbg):
        p0, p1 = V[m, 0], V[m, 1]
        t0, t1 = 0.0, 1.0
        for k in range(2):
            dk = p1[k] - p0[k]
            if dk != 0:
                ta, tb = sorted(((lo[b, k] - p0[k]) / dk, (hi[b, k] - p0[k]) / dk))
                t0, t1 = max(t0, ta), min(t1, tb)
        a, c = p0 + t0 * (p1 - p0), p0 + t1 * (p1 - p0)
        pts.append(a[None] + gx[:, None] * (c - a)[None]); wts.append(gw * np.linalg.norm(c - a)); off.append(off[-1] + n1d)
    B = fo.coupling_matrix(lo, hi, bd, ns, *none, V, idf, ni, pim, pbg, np.array(off), np.concatenate(pts), np.concatenate(wts), sb, si)
    assert (A != B).nnz == 0 or abs(A - B).max() <= 1e-13 * abs(A).max()
