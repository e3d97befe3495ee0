"""Codimension 0 in 2D -- the reference's <2,2,2> instantiation (code/src/intersections.cc:307-355, quad n quad; explicit
instantiation code/src/coupling_utilities.cc:72-102) -- on the CPU restatements: the C oracle (box clipped by the edges of
the immersed quad, fan from the left-most vertex, QGaussSimplex<2>) against the exact-rational oracle (quad clipped by the
box, exact integrals), and the identities any correct result satisfies."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from non_matching_test_suite_b200.grids import ImmersedGrid, make_config, solid_grid_2d
from oracle import exact_oracle as eo


def _matrix(r, bg, im):
    return sp.csr_matrix((r["val"], r["col"].astype(np.int64), r["rowptr"].astype(np.int64)), shape=(bg.n_dofs, im.n_dofs))


@pytest.mark.parametrize("n", [3, 5])
def test_c_oracle_equals_exact_oracle(c_oracle, n):
    bg, _ = make_config("disk", 0, band_only=False)
    im = solid_grid_2d(n)
    r = c_oracle.assemble(bg, im)
    e = eo.assemble(bg, im)
    acc = [(int(k >> np.uint64(32)), int(k & np.uint64(0xffffffff))) for k in r["accepted"]]
    assert acc == e["accepted"]
    assert np.abs(np.array(e["sum_w"]) - r["sum_w"]).max() <= 1e-16
    A = _matrix(r, bg, im).tocoo()
    assert {(int(i), int(j)) for i, j in zip(A.row, A.col)} == e["pattern"]
    assert max(abs(e["values"].get((int(i), int(j)), 0.0) - v) for i, j, v in zip(A.row, A.col, A.data)) <= 1e-16


@pytest.mark.parametrize("cycle,n", [(1, 8), (3, 40)])
def test_identities(c_oracle, cycle, n):
    bg, _ = make_config("disk", cycle, band_only=False, dirichlet=False)
    im = solid_grid_2d(n)
    r = c_oracle.assemble_with_quadrature(bg, im)
    area = im.measure().numpy()
    m = (r["accepted"] >> np.uint64(32)).astype(np.int64)
    slack = 5e-13 * r["dropped"]          # triangles under the reference's |J| < 1e-12 are left out (intersections.cc:710)
    # the pieces of an immersed cell add up to the cell; the shape functions to one; linear functions are reproduced
    assert np.abs(np.bincount(m, weights=r["sum_w"], minlength=im.n_cells) - area).max() <= 1e-15 + slack
    assert np.abs(r["local"].sum(1) - r["sum_w"]).max() <= 1e-16
    A = _matrix(r, bg, im)
    x = bg.dof_coords.numpy()
    T = _full_T(bg)
    cell_x_moment = np.bincount(np.repeat(m, r["nq"]), weights=r["weights"] * r["points"][:, 0], minlength=im.n_cells)
    assert np.abs(A.T @ (T @ x[:, 0]) - cell_x_moment).max() <= 1e-15          # C applied to the nodal values of x = integral of x
    assert abs(r["weights"].sum() - area.sum()) <= 1e-14 + slack and (r["weights"] > 0).all()
    # every rule that is exact for the bilinear integrand gives the same matrix
    r2 = c_oracle.assemble(bg, im, n1d=2)
    assert np.array_equal(r2["col"], r["col"]) and np.linalg.norm(r2["val"] - r["val"]) <= 1e-13 * np.linalg.norm(r["val"])


def _full_T(bg):
    """hanging nodes replaced by the interpolation of their masters: applied to the nodal values of a linear function it
    is the identity, and A^T (T x) is what the constrained matrix integrates"""
    from oracle import lm_poisson as lp
    return lp.reduction_matrix(bg)[0]


def _two_fine_cells_in_a_row(bg):
    """(h, x, y): lower-left corner of a finest-level cell whose right neighbour is a finest-level cell too"""
    size = (bg.hi - bg.lo)[:, 0]
    h = float(size.min())
    fine = {(round(float(x) / h), round(float(y) / h)) for x, y in bg.lo[size == size.min()].tolist()}
    i, j = next((i, j) for (i, j) in sorted(fine) if (i + 1, j) in fine)
    return h, i * h, j * h


def test_grid_aligned_quads_touching_cells_are_not_pairs(c_oracle):
    """Immersed quads whose edges lie ON grid lines: the neighbours across the line share only an edge (measure zero) and
    must not be accepted (the reference keeps 2D results only)."""
    bg, _ = make_config("disk", 0, band_only=False)
    h, x0, y0 = _two_fine_cells_in_a_row(bg)
    v = torch.tensor([[[x0, y0], [x0 + 2 * h, y0], [x0, y0 + h], [x0 + 2 * h, y0 + h]]], dtype=torch.float64)
    im = ImmersedGrid(2, 2, v, torch.zeros(1, 1, dtype=torch.int32), 1)
    r = c_oracle.assemble(bg, im)
    assert len(r["candidates"]) > len(r["accepted"]) and abs(r["sum_w"].sum() - 2 * h * h) <= 1e-16
    fine = (bg.hi - bg.lo)[:, 0].numpy()[(r["accepted"] & np.uint64(0xffffffff)).astype(np.int64)]
    assert np.allclose(r["sum_w"], fine * fine)                               # whole cells, nothing else
