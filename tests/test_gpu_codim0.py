"""Codimension 0 in 2D on the GPU -- the reference's <2,2,2> instantiation (code/src/intersections.cc:307-355; explicit
instantiation code/src/coupling_utilities.cc:72-102): collect -> sparsity -> mass matrix for FE_Q(1) x FE_DGQ(0) through
the C ABI against the C oracle (candidates, pairs, pattern bit-exact; values 1e-12) and the exact-rational oracle."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from non_matching_test_suite_b200 import _lib, coupling
from non_matching_test_suite_b200.grids import ImmersedGrid, make_config, solid_grid_2d
from oracle import exact_oracle as eo

pytestmark = pytest.mark.gpu


def _keys(im, bg):
    return (im.numpy().astype(np.uint64) << np.uint64(32)) | bg.numpy().astype(np.uint64)


@pytest.mark.parametrize("cycle,n,band", [(0, 5, False), (2, 16, False), (4, 100, False), (6, 400, True)])
def test_parity_against_c_oracle(c_oracle, cycle, n, band):
    bg, _ = make_config("disk", cycle, band_only=False)
    im = solid_grid_2d(n)
    ref = c_oracle.assemble(bg, im)
    ctx = coupling.make_context(bg, im, boxes=not band)              # both layouts of the background
    counts = ctx.intersect(3, 1e-15)
    ci, cb = ctx.candidates()
    assert np.array_equal(_keys(ci, cb), ref["candidates"])
    pi, pb, sw = ctx.pairs()
    assert np.array_equal(_keys(pi, pb), ref["accepted"])
    # triangles under the reference's |J| < 1e-12 (area < 5e-13) are left out per piece of the subdivision -- the fan here,
    # the fan from another vertex in the oracle: the only admissible difference, bounded by their number
    slack = 5e-13 * (counts["n_dropped"] + ref["dropped"])
    assert np.abs(sw.numpy() - ref["sum_w"]).max() <= 1e-13 * np.abs(ref["sum_w"]).max() + slack
    assert counts["n_accepted"] == len(ref["accepted"]) and counts["n_marginal"] == 0
    rp, col, val = ctx.csr()
    assert np.array_equal(rp.numpy().astype(np.uint64), ref["rowptr"]) and np.array_equal(col.numpy().astype(np.uint32), ref["col"])
    assert np.linalg.norm(val.numpy() - ref["val"]) <= 1e-12 * np.linalg.norm(ref["val"]) + slack
    # the area of the body, cell by cell
    area = im.measure().numpy()
    got = np.bincount(pi.numpy().astype(np.int64), weights=sw.numpy(), minlength=im.n_cells)
    assert np.abs(got - area).max() <= 1e-14 + slack
    # products
    x = torch.rand(im.n_dofs, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    A = sp.csr_matrix((ref["val"], ref["col"].astype(np.int64), ref["rowptr"].astype(np.int64)), shape=(bg.n_dofs, im.n_dofs))
    assert np.abs(ctx.spmv(x).numpy() - A @ x.numpy()).max() <= 1e-14 + slack
    ctx.close()


def test_against_the_exact_oracle():
    bg, _ = make_config("disk", 0, band_only=False)
    im = solid_grid_2d(5)
    e = eo.assemble(bg, im)
    ctx = coupling.make_context(bg, im)
    ctx.intersect(3, 1e-15)
    pi, pb, sw = ctx.pairs()
    assert list(zip(pi.tolist(), pb.tolist())) == e["accepted"]
    assert np.abs(sw.numpy() - np.array(e["sum_w"])).max() <= 1e-16
    rp, col, val = ctx.csr()
    A = sp.csr_matrix((val.numpy(), col.numpy(), rp.numpy()), shape=(bg.n_dofs, im.n_dofs)).tocoo()
    assert {(int(i), int(j)) for i, j in zip(A.row, A.col)} == e["pattern"]
    assert max(abs(e["values"].get((int(i), int(j)), 0.0) - v) for i, j, v in zip(A.row, A.col, A.data)) <= 1e-16
    ctx.close()


@pytest.mark.parametrize("degree", [1, 2, 3])
def test_quadrature_round_trip(c_oracle, degree):
    """The points the collect call hands out, fed back through the from-quadrature entry point (the reference's
    signature), rebuild the matrix; their weights add up to the area; rules 2 and 3 are exact for the integrand."""
    bg, _ = make_config("disk", 2, band_only=False)
    im = solid_grid_2d(24)
    ctx = coupling.make_context(bg, im, boxes=True)
    counts = ctx.intersect(degree, 1e-15)
    fused = ctx.csr()
    q_off, pts, w = ctx.quadrature()
    assert int(q_off[-1]) == counts["n_qpoints"] and pts.shape == (counts["n_qpoints"], 2)
    assert abs(float(w.sum()) - float(im.measure().sum())) <= 1e-13 and bool((w > 0).all())
    pi, pb, _ = ctx.pairs()
    order = torch.randperm(len(pi), generator=torch.Generator().manual_seed(degree))
    cnt = (q_off[1:] - q_off[:-1])[order]
    off2 = torch.cat([torch.zeros(1, dtype=q_off.dtype), torch.cumsum(cnt, 0)])
    idx = torch.cat([torch.arange(int(q_off[k]), int(q_off[k + 1])) for k in order.tolist()])
    ctx.assemble_from_quadrature(pi[order].contiguous(), pb[order].contiguous(), off2.contiguous(), pts[idx].contiguous(), w[idx].contiguous())
    again = ctx.csr()
    assert torch.equal(again[0], fused[0]) and torch.equal(again[1], fused[1])
    assert np.linalg.norm((again[2] - fused[2]).numpy()) <= 1e-13 * np.linalg.norm(fused[2].numpy())
    if degree >= 2:
        ref = c_oracle.assemble(bg, im, n1d=3)
        assert np.linalg.norm(fused[2].numpy() - ref["val"]) <= 1e-12 * np.linalg.norm(ref["val"])
    ctx.close()


def _two_fine_cells_in_a_row(bg):
    """(h, x, y): lower-left corner of a finest-level cell whose right neighbour is a finest-level cell too"""
    size = (bg.hi - bg.lo)[:, 0]
    h = float(size.min())
    fine = {(round(float(x) / h), round(float(y) / h)) for x, y in bg.lo[size == size.min()].tolist()}
    i, j = next((i, j) for (i, j) in sorted(fine) if (i + 1, j) in fine)
    return h, i * h, j * h


def test_grid_aligned_quads_and_errors(c_oracle):
    bg, _ = make_config("disk", 0, band_only=False)
    h, x0, y0 = _two_fine_cells_in_a_row(bg)
    v = torch.tensor([[[x0, y0], [x0 + 2 * h, y0], [x0, y0 + h], [x0 + 2 * h, y0 + h]]], dtype=torch.float64)
    im = ImmersedGrid(2, 2, v, torch.zeros(1, 1, dtype=torch.int32), 1)
    ref = c_oracle.assemble(bg, im)
    ctx = coupling.make_context(bg, im)
    counts = ctx.intersect(3, 1e-15)
    pi, pb, sw = ctx.pairs()
    assert np.array_equal(_keys(pi, pb), ref["accepted"]) and counts["n_candidates"] > counts["n_accepted"]
    assert abs(float(sw.sum()) - 2 * h * h) <= 1e-16
    # a quad that turns both ways is refused, not approximated
    bad = v.clone()
    bad[0, 3] = torch.tensor([x0 + 0.5 * h, y0 + 0.3 * h])
    with pytest.raises(_lib.NmError) as e:
        ctx.set_immersed(2, 2, bad, torch.zeros(1, 1, dtype=torch.int32), 1)
    assert e.value.code == 3 and "not convex" in str(e.value)
    # general elements and the single-call host entry point state that they do not take this case
    ctx.set_immersed(2, 2, v, torch.zeros(1, 1, dtype=torch.int32), 1)
    with pytest.raises(_lib.NmError) as e:
        ctx.set_finite_elements((1, np.array([[0, 0], [1, 0], [0, 1], [1, 1]], dtype=float)), bg.dofs.to(torch.int32).contiguous(), bg.n_dofs,
                                (0, np.array([[0.5, 0.5]])), torch.zeros(1, 1, dtype=torch.int32), 1)
    assert e.value.code == 3
    # a volume in 3D is outside the implemented scope
    with pytest.raises(_lib.NmError) as e:
        ctx.set_immersed(3, 3, torch.zeros(1, 8, 3, dtype=torch.float64), torch.zeros(1, 1, dtype=torch.int32), 1)
    assert e.value.code == 3
    ctx.close()
