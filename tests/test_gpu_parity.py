"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical grids.

Bars (BASELINE.json north_star): candidate-pair set and CSR sparsity bit-exact; matrix
entries within 1e-12 relative Frobenius norm (tolerance written below); accepted-pair set
bit-exact except pairs below the 1e-12 relative-measure threshold, which are reported.
"""
import numpy as np
import pytest
import torch

from non_matching_test_suite_b200 import coupling
from non_matching_test_suite_b200.grids import make_config

pytestmark = pytest.mark.gpu

REL_FROB_TOL = 1e-12

CASES = [("disk", 0), ("disk", 2), ("disk", 4), ("flower", 0), ("flower", 2), ("sphere", 0), ("sphere", 1)]


def _keys(im, bg):
    return (im.numpy().astype(np.uint64) << np.uint64(32)) | bg.numpy().astype(np.uint32).astype(np.uint64)


@pytest.mark.parametrize("name,cycle", CASES)
@pytest.mark.parametrize("boxes", [False, True])
def test_parity_against_c_oracle(c_oracle, name, cycle, boxes):
    bg, im = make_config(name, cycle, band_only=False)
    ref = c_oracle.assemble(bg, im, n1d=3, tol=1e-15)
    ctx = coupling.make_context(bg, im, boxes=boxes)
    counts = ctx.intersect(3, 1e-15)
    # candidates: bit-exact set, same (immersed, background) order
    cim, cbg = ctx.candidates()
    assert np.array_equal(_keys(cim, cbg), ref["candidates"])
    # accepted pairs
    aim, abg, sw = ctx.pairs()
    assert counts["n_marginal"] == 0
    assert np.array_equal(_keys(aim, abg), ref["accepted"])
    # per-pair measure: sub-simplices come from a different subdivision than the oracle's, so tiny
    # clipped pieces carry cancellation error relative to the parent simplex, not to themselves
    rel = np.abs(sw.numpy() - ref["sum_w"]) / np.maximum(np.abs(ref["sum_w"]), 1e-300)
    assert rel.max() <= 1e-9, rel.max()
    assert abs(sw.numpy().sum() - ref["sum_w"].sum()) <= 1e-13 * ref["sum_w"].sum()
    if bg.spacedim == 3:
        assert np.array_equal(ctx.split_codes().numpy(), ref["codes"])
    # sparsity: bit-exact CSR; values: relative Frobenius norm
    rp, col, val = ctx.csr(transpose=False)
    assert np.array_equal(rp.numpy().astype(np.uint64), ref["rowptr"])
    assert np.array_equal(col.numpy().astype(np.uint32), ref["col"])
    err = np.linalg.norm(val.numpy() - ref["val"]) / np.linalg.norm(ref["val"])
    assert err <= REL_FROB_TOL, err
    # transposed CSR is the same matrix
    rpt, colt, valt = ctx.csr(transpose=True)
    import scipy.sparse as sp
    A = sp.csr_matrix((val.numpy(), col.numpy(), rp.numpy()), shape=(bg.n_dofs, im.n_dofs))
    At = sp.csr_matrix((valt.numpy(), colt.numpy(), rpt.numpy()), shape=(im.n_dofs, bg.n_dofs))
    assert abs(A - At.T).max() == 0
    # SpMV both ways, host and device vectors
    g = torch.Generator().manual_seed(0)
    x = torch.rand(im.n_dofs, dtype=torch.float64, generator=g) * 2 - 1
    u = torch.rand(bg.n_dofs, dtype=torch.float64, generator=g) * 2 - 1
    y = ctx.spmv(x, transpose=False)
    z = ctx.spmv(u, transpose=True)
    assert np.allclose(y.numpy(), A @ x.numpy(), rtol=1e-13, atol=1e-15)
    assert np.allclose(z.numpy(), A.T @ u.numpy(), rtol=1e-13, atol=1e-15)
    yd = ctx.spmv(x.cuda(), transpose=False).cpu()
    zd = ctx.spmv(u.cuda(), transpose=True).cpu()
    assert torch.equal(yd, y) and torch.equal(zd, z)
    ctx.close()


@pytest.mark.parametrize("name,cycle", [("disk", 1), ("sphere", 0)])
def test_quadrature_roundtrip(c_oracle, name, cycle):
    """collect -> (points, weights) -> mass matrix from the given quadrature == fused path."""
    bg, im = make_config(name, cycle, band_only=False)
    info = coupling.collect_quadratures_on_overlapped_grids(bg, im, 3, 1e-15)
    off, pts, w = info.q_offset, info.points, info.weights
    assert int(off[-1]) == info.counts["n_qpoints"] == w.shape[0]
    # weights of a pair sum to its measure; points lie inside the (closed) background cell
    seg = torch.repeat_interleave(torch.arange(len(info)), off[1:] - off[:-1])
    sums = torch.zeros(len(info), dtype=torch.float64).index_add_(0, seg, w)
    assert torch.allclose(sums, info.sum_w, rtol=1e-13, atol=0)
    lo, hi = bg.lo[info.space_cell.long()][seg], bg.hi[info.space_cell.long()][seg]
    assert bool(((pts >= lo - 1e-14) & (pts <= hi + 1e-14)).all())
    rp0, col0 = coupling.create_coupling_sparsity_pattern_with_exact_intersections(info, bg, im)
    M = coupling.create_coupling_mass_matrix_with_exact_intersections(bg, im, info, use_given_quadrature=True)
    rp1, col1, val1 = M.csr()
    assert torch.equal(rp0, rp1) and torch.equal(col0, col1)
    M2 = coupling.create_coupling_mass_matrix_with_exact_intersections(bg, im, info, use_given_quadrature=False)
    rp2, col2, val2 = M2.csr()
    assert torch.equal(rp1, rp2) and torch.equal(col1, col2)
    err = (val1 - val2).norm() / val2.norm()
    assert err <= REL_FROB_TOL, err
    info.ctx.close()
