"""Graded backgrounds (grids.graded_config: hanging nodes on cut cells all along the interface) on the GPU: the constraint
expansion of the single-pass merge, of the general merge and of the Nitsche terms against the C oracle -- pattern bit-exact
(kept constrained rows included), values to 1e-12 -- and at scale through the column-sum identity."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from non_matching_test_suite_b200 import coupling
from non_matching_test_suite_b200.grids import graded_config

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,cycle,stripe,general", [("disk", 3, 1, False), ("flower", 2, 4, False), ("sphere", 1, 1, False),
                                                        ("sphere", 2, 8, False), ("sphere", 1, 1, True)])
def test_parity_on_graded_grids(monkeypatch, c_oracle, name, cycle, stripe, general):
    if general:
        monkeypatch.setenv("NMGPU_FORCE_GENERAL_MERGE", "1")
    bg, im = graded_config(name, cycle, stripe=stripe, band_only=False)
    ref = c_oracle.assemble(bg, im)
    cut = np.unique((ref["accepted"] & np.uint64(0xffffffff)).astype(np.int64))
    on_cut = np.intersect1d(np.unique(bg.dofs.numpy()[cut].ravel()), bg.c_dof.numpy())
    assert len(on_cut) >= 4
    ctx = coupling.make_context(bg, im, boxes=True)
    counts = ctx.intersect(3, 1e-15)
    assert counts["n_accepted"] == len(ref["accepted"])
    rp, col, val = ctx.csr()
    assert ctx.info("merge_path") == (1 if general else 0)
    assert np.array_equal(rp.numpy().astype(np.uint64), ref["rowptr"]) and np.array_equal(col.numpy().astype(np.uint32), ref["col"])
    assert np.linalg.norm(val.numpy() - ref["val"]) <= 1e-12 * np.linalg.norm(ref["val"])
    A = sp.csr_matrix((val.numpy(), col.numpy(), rp.numpy()), shape=(bg.n_dofs, im.n_dofs))
    assert A[on_cut].nnz > 0 and np.abs(A[on_cut].data).max() == 0.0          # constrained rows: kept, zero
    ctx.close()


def test_nitsche_on_a_graded_grid(c_oracle):
    """The interface-penalisation terms with constrained dofs on cut cells (the branch the round-1 timing never ran)."""
    bg, im = graded_config("sphere", 1, stripe=1, band_only=False)
    info = coupling.collect_quadratures_on_overlapped_grids(bg, im, 3, 1e-15)
    coef, f = 2.0 + info.points[:, 0], 1.0 + info.points[:, 1] ** 2
    A_ref, b_ref = c_oracle.nitsche(bg, info.space_cell, info.q_offset, info.points, info.weights, coef, f, 10.0)
    (rp, col, val), rhs = info.ctx.assemble_nitsche(info.space_cell.contiguous(), info.q_offset.contiguous(), info.points.contiguous(),
                                                    info.weights.contiguous(), coefficient=coef.contiguous(), rhs_values=f.contiguous(), penalty=10.0)
    G = sp.csr_matrix((val.numpy(), col.numpy(), rp.numpy()), shape=(bg.n_dofs, bg.n_dofs))
    assert abs(G - A_ref).max() <= 1e-12 * abs(A_ref).max()
    assert np.linalg.norm(rhs.numpy() - b_ref) <= 1e-12 * np.linalg.norm(b_ref)
    info.ctx.close()


def test_column_sums_at_scale():
    """sphere cycle 5, graded, band-only (1.5e5 quads): the weights of every constraint line add up to one, so the columns of
    the stored matrix still sum to the measure the immersed cell cuts out, whatever was distributed to masters."""
    bg, im = graded_config("sphere", 5, stripe=1, band_only=True, device="cuda")
    ctx = coupling.make_context(bg, im, boxes=True, on_device=True)
    counts = ctx.intersect(3, 1e-15)
    pi, _, sw = ctx.pairs(device="cuda")
    per_cell = torch.zeros(im.n_cells, dtype=torch.float64, device="cuda").index_add_(0, pi.long(), sw)
    rp, col, val = ctx.csr(device="cuda")
    colsum = torch.zeros(im.n_dofs, dtype=torch.float64, device="cuda").index_add_(0, col.long(), val)
    assert float((colsum[im.dofs[:, 0].long()] - per_cell).abs().max()) <= 1e-15
    assert counts["n_accepted"] > 1.0e6 and int(bg.c_dof.numel()) > 1000 and ctx.info("merge_path") == 0
    ctx.close()
