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


@pytest.mark.parametrize("name,cycle", [("disk", 1), ("sphere", 0), ("sphere", 4), ("flower", 6)])
def test_quadrature_roundtrip(c_oracle, name, cycle):
    """collect -> (points, weights) -> mass matrix from the given quadrature == fused path."""
    bg, im = make_config(name, cycle, band_only=cycle > 2)
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


GOLD = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name,cycle", [("disk", 0), ("disk", 1), ("flower", 0), ("sphere", 0)])
def test_against_golden_fixtures(name, cycle):
    """Committed exact-rational fixtures (tests/golden/make_golden.py): candidates, accepted pairs and CSR
    pattern bit-exact, values within 1e-12 relative Frobenius norm."""
    g = np.load(__import__("os").path.join(GOLD, "%s_c%d.npz" % (name, cycle)))
    bg, im = make_config(name, cycle, band_only=False)
    ctx = coupling.make_context(bg, im)
    ctx.intersect(3, 1e-15)
    assert np.array_equal(_keys(*ctx.candidates()), g["candidates"])
    aim, abg, sw = ctx.pairs()
    assert np.array_equal(_keys(aim, abg), g["accepted"])
    rp, col, val = ctx.csr()
    assert np.array_equal(rp.numpy().astype(np.uint64), g["rowptr"])
    assert np.array_equal(col.numpy().astype(np.uint32), g["col"])
    assert np.linalg.norm(val.numpy() - g["val"]) <= REL_FROB_TOL * np.linalg.norm(g["val"])
    ctx.close()


def test_quadrature_kernel_equals_moment_kernel(monkeypatch):
    """The per-point kernel (reference rule, kept for rules below degree 3 and for emission) and the
    closed-form moment kernel integrate the same thing."""
    bg, im = make_config("sphere", 1, band_only=False)
    ctx = coupling.make_context(bg, im)
    ctx.intersect(3, 1e-15)
    rp, col, val = ctx.csr()
    ctx.close()
    monkeypatch.setenv("NMGPU_FORCE_QUADRATURE", "1")
    ctx2 = coupling.make_context(bg, im)
    c2 = ctx2.intersect(3, 1e-15)
    rp2, col2, val2 = ctx2.csr()
    assert torch.equal(rp, rp2) and torch.equal(col, col2)
    assert (val - val2).norm() <= REL_FROB_TOL * val.norm()
    ctx2.close()


def test_error_behaviour():
    from non_matching_test_suite_b200 import _lib
    bg, im = make_config("disk", 0, band_only=False)
    ctx = coupling.make_context(bg, im)
    with pytest.raises(_lib.NmError) as e:       # reference: AssertThrow(degree > 0)
        ctx.intersect(0, 1e-15)
    assert e.value.code == 1 and "Invalid quadrature degree" in str(e.value)
    with pytest.raises(_lib.NmError) as e:       # QGaussSimplex<1>(9) is not tabulated
        ctx.intersect(9, 1e-15)
    assert e.value.code == 3
    fresh = _lib.Context(0)
    with pytest.raises(_lib.NmError) as e:       # call order
        fresh.build_sparsity()
    assert e.value.code == 4
    # a sheared background cell is outside the implemented scope and must be refused, not approximated
    v = bg.vertices().clone()
    v[3, 1, 0] += 0.01
    with pytest.raises(_lib.NmError) as e:
        fresh.set_background(2, verts=v.contiguous(), dofs=bg.dofs.contiguous(), n_dofs=bg.n_dofs)
    assert e.value.code == 3
    # NM_DEVICE_INPLACE / NM_HOST_ASYNC are accepted only by the calls nmgpu.h names: anywhere else they are an error, not a
    # pointer of the wrong kind
    import ctypes as C
    n_cand = ctx.intersect(3, 1e-15)["n_candidates"]
    bi, bb = torch.empty(n_cand, dtype=torch.int32), torch.empty(n_cand, dtype=torch.int32)
    for bad_where in (_lib.NM_DEVICE_INPLACE, _lib.NM_HOST_ASYNC, 7):
        rc = ctx.L.nm_get_candidates(ctx.h, C.c_void_p(bi.data_ptr()), C.c_void_p(bb.data_ptr()), bad_where)
        assert rc == 1 and b"where must be" in ctx.L.nm_last_error(ctx.h)
    fresh.close(); ctx.close()


def test_empty_and_disjoint_inputs():
    bg, im = make_config("disk", 0, band_only=False)
    far = im.slice(0, 4)
    far = type(far)(far.dim, far.spacedim, far.verts + 10.0, far.dofs, far.n_dofs)
    ctx = coupling.make_context(bg, far)
    c = ctx.intersect(3, 1e-15)
    assert c["n_candidates"] == 0 and c["n_accepted"] == 0
    assert ctx.build_sparsity() == 0
    y = ctx.spmv(torch.ones(im.n_dofs, dtype=torch.float64))
    assert float(y.abs().sum()) == 0.0
    ctx.close()
    none = im.slice(0, 0)
    ctx = coupling.make_context(bg, none)
    c = ctx.intersect(3, 1e-15)
    assert c["n_candidates"] == 0 and ctx.build_sparsity() == 0
    ctx.close()


@pytest.mark.parametrize("degree", [1, 2, 4])
def test_other_rules_run(degree):
    """n_points_1D other than 3: 2D rules 1..4 are tabulated; in 3D 1..3 (4 is refused)."""
    bg, im = make_config("disk", 1, band_only=False)
    ctx = coupling.make_context(bg, im)
    c = ctx.intersect(degree, 1e-15)
    assert c["n_qpoints"] == degree * c["n_accepted"]   # one sub-segment per pair (Liang-Barsky), `degree` points each
    rp, col, val = ctx.csr()
    ctx3 = coupling.make_context(bg, im)
    ctx3.intersect(3, 1e-15)
    _, col3, val3 = ctx3.csr()
    assert torch.equal(col, col3)
    if degree >= 2:                                       # Gauss n>=2 is exact for the quadratic integrand
        assert (val - val3).norm() <= REL_FROB_TOL * val3.norm()
    ctx.close(); ctx3.close()


def test_full_size_properties():
    """Size-independent properties on a refined sphere (band only): sum of weights = immersed measure,
    partition of unity through the constraint lines, C and C^T products consistent."""
    bg, im = make_config("sphere", 5, band_only=True, device="cuda")
    ctx = coupling.make_context(bg, im, on_device=True)
    c = ctx.intersect(3, 1e-15)
    assert c["n_accepted"] > 900000 and c["n_marginal"] == 0
    aim, abg, sw = ctx.pairs(device="cuda")
    per_cell = torch.zeros(im.n_cells, dtype=torch.float64, device="cuda").index_add_(0, aim.long(), sw)
    # measure of the split quads
    from oracle import exact_oracle as eo
    codes = ctx.split_codes().cuda()
    v = im.verts
    meas = torch.zeros(im.n_cells, dtype=torch.float64, device="cuda")
    for t, (i, j, k) in enumerate(eo.TRIPLES):
        on = ((codes >> t) & 1).bool()
        meas += on * 0.5 * torch.linalg.cross(v[:, j] - v[:, i], v[:, k] - v[:, i]).norm(dim=1)
    # sub-simplices below |J| = 1e-12 are dropped as in the reference: allow exactly that much
    assert float((per_cell - meas).abs().max()) <= 1e-12 * max(1, c["n_dropped"]) + 1e-13 * float(meas.max())
    ones_u = torch.ones(bg.n_dofs, dtype=torch.float64, device="cuda")
    z = ctx.spmv(ones_u, transpose=True)                  # column sums = per-cell measure (partition of unity)
    assert torch.allclose(z, per_cell, rtol=1e-11, atol=1e-18)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(im.n_dofs, dtype=torch.float64, device="cuda", generator=g)
    u = torch.rand(bg.n_dofs, dtype=torch.float64, device="cuda", generator=g)
    lhs = torch.dot(ctx.spmv(x, transpose=False), u)      # <A x, u> = <x, A^T u>
    rhs = torch.dot(x, ctx.spmv(u, transpose=True))
    assert abs(float(lhs - rhs)) <= 1e-12 * abs(float(lhs))
    ctx.close()


def test_adjust_grids_flagging():
    """nm_flag_overlapped_background == brute-force closed-box overlap (the drivers' adjust_grids loop)."""
    bg, im = make_config("flower", 0, band_only=False)
    ctx = coupling.make_context(bg, im)
    flags = ctx.flag_overlapped_background(bg.n_cells)
    blo, bhi = im.boxes()
    ov = ((bg.lo[:, None, :] <= bhi[None]) & (blo[None] <= bg.hi[:, None, :])).all(-1).any(1)
    assert torch.equal(flags.bool(), ov)
    ctx.close()


def _random_problem(d, n_bg, n_im, seed):
    """Background = arbitrary axis-aligned boxes (not a tree: sizes over 1.5 decades, overlapping, unaligned),
    every box with its own 2^d DoFs; immersed = random small segments / planar-ish quads."""
    from non_matching_test_suite_b200.grids import BackgroundGrid, ImmersedGrid
    g = torch.Generator().manual_seed(seed)
    size = 10.0 ** (-2.0 + 1.5 * torch.rand(n_bg, 1, dtype=torch.float64, generator=g)) * (0.5 + torch.rand(n_bg, d, dtype=torch.float64, generator=g))
    lo = torch.rand(n_bg, d, dtype=torch.float64, generator=g) * 0.9
    hi = lo + size
    dofs = torch.arange(n_bg * (1 << d), dtype=torch.int32).reshape(n_bg, 1 << d)
    e = torch.empty(0, dtype=torch.int32)
    bg = BackgroundGrid(d, lo, hi, dofs, n_bg * (1 << d), torch.zeros(n_bg, dtype=torch.int32), e, torch.zeros(1, dtype=torch.int64), e,
                        torch.empty(0, dtype=torch.float64))
    c = torch.rand(n_im, 1, d, dtype=torch.float64, generator=g)
    if d == 2:
        verts = c + 0.08 * (torch.rand(n_im, 2, 2, dtype=torch.float64, generator=g) - 0.5)
    else:
        u = 0.08 * (torch.rand(n_im, 1, 3, dtype=torch.float64, generator=g) - 0.5)
        v = 0.08 * (torch.rand(n_im, 1, 3, dtype=torch.float64, generator=g) - 0.5)
        w = 0.004 * (torch.rand(n_im, 1, 3, dtype=torch.float64, generator=g) - 0.5)     # slightly non-planar
        verts = torch.cat([c, c + u, c + v, c + u + v + w], dim=1)
    im = ImmersedGrid(d - 1, d, verts.contiguous(), torch.arange(n_im, dtype=torch.int32)[:, None], n_im)
    return bg, im


@pytest.mark.parametrize("d", [2, 3])
def test_arbitrary_boxes_against_oracle(c_oracle, d):
    """The grid hash makes no assumption about the background being a tree."""
    bg, im = _random_problem(d, 3000, 400, seed=d)
    ref = c_oracle.assemble(bg, im, n1d=3, tol=1e-15)
    ctx = coupling.make_context(bg, im, boxes=True)
    counts = ctx.intersect(3, 1e-15)
    assert counts["n_candidates"] > 2000
    assert np.array_equal(_keys(*ctx.candidates()), ref["candidates"])
    aim, abg, sw = ctx.pairs()
    assert np.array_equal(_keys(aim, abg), ref["accepted"])
    if d == 3:
        assert np.array_equal(ctx.split_codes().numpy(), ref["codes"])
    rp, col, val = ctx.csr()
    assert np.array_equal(rp.numpy().astype(np.uint64), ref["rowptr"])
    assert np.array_equal(col.numpy().astype(np.uint32), ref["col"])
    assert np.linalg.norm(val.numpy() - ref["val"]) <= REL_FROB_TOL * np.linalg.norm(ref["val"])
    ctx.close()


def test_big_immersed_cells_take_the_general_merge_path(c_oracle):
    """Immersed cells much larger than the background cells: hundreds of pairs per immersed cell overflow the
    fast merge (192 distinct DoFs) and must come out identical through the two-pass path."""
    from non_matching_test_suite_b200.grids import ImmersedGrid, circle_grid
    bg, _ = make_config("disk", 4, band_only=False)
    im = circle_grid(8)                                                      # 8 long segments across a fine band
    ref = c_oracle.assemble(bg, im)
    ctx = coupling.make_context(bg, im)
    c = ctx.intersect(3, 1e-15)
    assert c["n_accepted"] / im.n_cells > 120
    rp, col, val = ctx.csr()
    assert np.array_equal(col.numpy().astype(np.uint32), ref["col"])
    assert np.array_equal(rp.numpy().astype(np.uint64), ref["rowptr"])
    assert np.linalg.norm(val.numpy() - ref["val"]) <= REL_FROB_TOL * np.linalg.norm(ref["val"])
    ctx.close()


@pytest.mark.parametrize("name,cycle", [("sphere", 3), ("sphere", 4), ("flower", 5), ("disk", 7)])
def test_parity_at_oracle_sized_refinements(c_oracle, name, cycle):
    """Band-only refinements the C oracle still finishes in seconds (up to ~2.5e5 pairs): same bars as above.
    Sub-simplices under the reference's absolute |J| < 1e-12 threshold depend on the subdivision (tets vs box
    clipping); they are counted on both sides and bound the only admissible difference."""
    bg, im = make_config(name, cycle, band_only=True)
    ref = c_oracle.assemble(bg, im, n1d=3, tol=1e-15)
    ctx = coupling.make_context(bg, im, boxes=True)
    counts = ctx.intersect(3, 1e-15)
    assert np.array_equal(_keys(*ctx.candidates()), ref["candidates"])
    aim, abg, sw = ctx.pairs()
    assert counts["n_marginal"] == 0
    assert np.array_equal(_keys(aim, abg), ref["accepted"])
    rp, col, val = ctx.csr()
    assert np.array_equal(rp.numpy().astype(np.uint64), ref["rowptr"])
    assert np.array_equal(col.numpy().astype(np.uint32), ref["col"])
    dropped = counts["n_dropped"] + ref["dropped"]
    err = np.linalg.norm(val.numpy() - ref["val"]) / np.linalg.norm(ref["val"])
    # each dropped sliver is worth at most 1e-12 of measure against entries of size |K|/8
    allowed = REL_FROB_TOL + dropped * 1e-12 / np.linalg.norm(ref["val"])
    assert err <= allowed, (err, allowed, dropped)
    ctx.close()


@pytest.mark.parametrize("d", [2, 3])
def test_degenerate_touching_cases_gpu(c_oracle, d):
    """Same hand-made degenerate cases as tests/test_oracle_cpu.py, GPU against the oracle."""
    from degenerate_cases import case_2d, case_3d
    bg, im = case_2d() if d == 2 else case_3d()
    ref = c_oracle.assemble(bg, im, n1d=3, tol=1e-15)
    ctx = coupling.make_context(bg, im)
    ctx.intersect(3, 1e-15)
    assert np.array_equal(_keys(*ctx.candidates()), ref["candidates"])
    aim, abg, sw = ctx.pairs()
    assert np.array_equal(_keys(aim, abg), ref["accepted"])
    assert np.allclose(sw.numpy(), ref["sum_w"], rtol=1e-12, atol=1e-15)
    if d == 3:
        assert np.array_equal(ctx.split_codes().numpy(), ref["codes"])
    rp, col, val = ctx.csr()
    assert np.array_equal(col.numpy().astype(np.uint32), ref["col"])
    assert np.linalg.norm(val.numpy() - ref["val"]) <= REL_FROB_TOL * np.linalg.norm(ref["val"])
    ctx.close()


def test_spmv_push_single_gpu():
    """nm_spmv_push (Ct.vmult fused with its cross-GPU reduction) on one GPU: the row results are ADDED into
    every target vector; with two local targets both receive the product on top of what they held.
    The multi-GPU run over NVLink peer / multicast memory is tests/multi_gpu_check.py (needs >= 2 GPUs)."""
    bg, im = make_config("sphere", 1, band_only=False)
    ctx = coupling.make_context(bg, im)
    ctx.intersect(3, 1e-15)
    ctx.build_sparsity()
    x = torch.rand(im.n_dofs, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    y_ref = ctx.spmv(x, transpose=False)
    y1 = torch.zeros(bg.n_dofs, dtype=torch.float64, device="cuda")
    y2 = torch.ones(bg.n_dofs, dtype=torch.float64, device="cuda")
    ctx.spmv_push(x, [y1.data_ptr(), y2.data_ptr()])
    assert torch.equal(y1, y_ref)
    assert torch.allclose(y2, y_ref + 1.0, rtol=0, atol=1e-15)
    ctx.close()


def test_index_range_limit_is_reported():
    """A background whose extent spans more than 2^19 finest cells per axis (3D) cannot be indexed by the packed
    grid key: the library must say so instead of aliasing keys."""
    from non_matching_test_suite_b200 import _lib
    from non_matching_test_suite_b200.grids import BackgroundGrid
    lo = torch.tensor([[0.0, 0.0, 0.0], [0.9, 0.9, 0.9]], dtype=torch.float64)
    hi = torch.tensor([[1e-7, 1e-7, 1e-7], [1.0, 1.0, 1.0]], dtype=torch.float64)
    dofs = torch.arange(16, dtype=torch.int32).reshape(2, 8)
    e = torch.empty(0, dtype=torch.int32)
    bg = BackgroundGrid(3, lo, hi, dofs, 16, torch.zeros(2, dtype=torch.int32), e, torch.zeros(1, dtype=torch.int64), e,
                        torch.empty(0, dtype=torch.float64))
    _, im = make_config("sphere", 0, band_only=False)
    ctx = coupling.make_context(bg, im, boxes=True)
    with pytest.raises(_lib.NmError) as err:
        ctx.intersect(3, 1e-15)
    assert err.value.code == 3 and "2^19" in str(err.value)
    ctx.close()


@pytest.mark.parametrize("name,cycle", [("disk", 2), ("sphere", 1)])
@pytest.mark.parametrize("general_merge", [False, True])
def test_compact_row_mode_is_the_same_matrix(monkeypatch, name, cycle, general_merge):
    """Compact-row storage of the stored orientation (used when a shard touches a small part of the global dofs,
    forced here with NMGPU_COMPACT_ROWS=1): identical CSR, identical products, identical fused push."""
    bg, im = make_config(name, cycle, band_only=False)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(im.n_dofs, dtype=torch.float64, generator=g) * 2 - 1
    u = torch.rand(bg.n_dofs, dtype=torch.float64, generator=g) * 2 - 1
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("NMGPU_COMPACT_ROWS", mode)
        if general_merge:
            monkeypatch.setenv("NMGPU_FORCE_GENERAL_MERGE", "1")
        ctx = coupling.make_context(bg, im, boxes=True)
        ctx.intersect(3, 1e-15)
        rp, col, val = ctx.csr(transpose=False)
        rpd, cold, vald = ctx.csr(transpose=False, device="cuda")
        assert torch.equal(rpd.cpu(), rp) and torch.equal(cold.cpu(), col) and torch.equal(vald.cpu(), val)
        y = ctx.spmv(x, transpose=False)
        yd = ctx.spmv(x.cuda(), transpose=False)
        assert torch.equal(yd.cpu(), y)
        z = ctx.spmv(u, transpose=True)
        t1 = torch.zeros(bg.n_dofs, dtype=torch.float64, device="cuda")
        ctx.spmv_push(x.cuda(), [t1.data_ptr()])
        assert torch.equal(t1.cpu(), y)
        out[mode] = (rp, col, val, y, z)
        ctx.close()
    for a, b in zip(out["0"], out["1"]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("name,cycle", [("disk", 2), ("sphere", 1)])
@pytest.mark.parametrize("pinned", [False, True])
def test_assemble_host_equals_separate_calls(name, cycle, pinned):
    """nm_assemble_host (uploads overlapped with the phases) leaves the context in the same state as the separate
    calls: same counts, same CSR bitwise, same products; the context can be reused for another problem."""
    from non_matching_test_suite_b200 import _lib
    bg, im = make_config(name, cycle, band_only=False)
    ref = coupling.make_context(bg, im, boxes=True)
    c_ref = ref.intersect(3, 1e-15)
    rp, col, val = ref.csr(transpose=False)
    pin = (lambda t: t.contiguous().pin_memory()) if pinned else (lambda t: t.contiguous())
    i32 = lambda t: t.to(torch.int32)
    args = (bg.spacedim, pin(bg.lo), pin(bg.hi), pin(i32(bg.dofs)), bg.n_dofs, pin(i32(bg.c_dof)), pin(bg.c_ptr.to(torch.int64)),
            pin(i32(bg.c_master)), pin(bg.c_weight), im.dim, pin(im.verts), pin(i32(im.dofs)), im.n_dofs)
    ctx = _lib.Context(0)
    for _ in range(2):          # second round: buffers are reused while nothing of the first is in flight
        c = ctx.assemble_host(*args, 3, 1e-15)
        assert c == c_ref
        rp2, col2, val2 = ctx.csr(transpose=False)
        assert torch.equal(rp2, rp) and torch.equal(col2, col) and torch.equal(val2, val)
    x = torch.rand(im.n_dofs, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    assert torch.equal(ctx.spmv(x), ref.spmv(x))
    # pairs / quadrature are available as after nm_intersect
    a, b = ctx.pairs(), ref.pairs()
    assert all(torch.equal(p, q) for p, q in zip(a, b))
    with pytest.raises(_lib.NmError) as e:
        ctx.assemble_host(*args, 0, 1e-15)
    assert e.value.code == 1 and "Invalid quadrature degree" in str(e.value)
    ctx.close(); ref.close()


def test_spmv_push_owned_single_gpu():
    """nm_spmv_push_owned with three pretend ranks on one GPU: each dof's row result lands only in the vector of the
    range that owns it; the last range takes the remainder."""
    from non_matching_test_suite_b200.distributed import owned_dof_range, rows_per_rank
    bg, im = make_config("sphere", 1, band_only=False)
    ctx = coupling.make_context(bg, im)
    ctx.intersect(3, 1e-15)
    ctx.build_sparsity()
    x = torch.rand(im.n_dofs, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    y_ref = ctx.spmv(x, transpose=False)
    world = 3
    ys = [torch.zeros(bg.n_dofs, dtype=torch.float64, device="cuda") for _ in range(world)]
    ctx.spmv_push_owned(x, [y.data_ptr() for y in ys], rows_per_rank(bg.n_dofs, world))
    for r, y in enumerate(ys):
        lo, hi = owned_dof_range(bg.n_dofs, r, world)
        assert torch.equal(y[lo:hi], y_ref[lo:hi])
        mask = torch.ones(bg.n_dofs, dtype=torch.bool, device="cuda"); mask[lo:hi] = False
        assert float(y[mask].abs().sum()) == 0.0
    ctx.close()


@pytest.mark.parametrize("name,cycle", [("disk", 2), ("sphere", 1)])
def test_mass_matrix_from_the_reference_subdivision_quadrature(c_oracle, name, cycle):
    """create_coupling_mass_matrix_with_exact_intersections fed with a FOREIGN cells_and_quads vector: the oracle's
    quadrature in the reference's own subdivision (CDT triangles / tetrahedra x Delaunay triangles: 2-3x the points of
    our collect, different order inside a pair), and in a shuffled tuple order.  The matrix is the oracle's: pattern
    bit-exact, values 1e-12 (rule R1: degree <= 3 integrand, both subdivisions are exact)."""
    bg, im = make_config(name, cycle, band_only=False)
    ref = c_oracle.assemble_with_quadrature(bg, im, 3, 1e-15)
    n = ref["accepted"].shape[0]
    perm = np.random.default_rng(7).permutation(n)
    qo = ref["q_offset"].astype(np.int64)
    cnt = np.diff(qo)[perm]
    off = np.concatenate([[0], np.cumsum(cnt)])
    idx = np.concatenate([np.arange(qo[p], qo[p + 1]) for p in perm])
    im_ids = (ref["accepted"] >> np.uint64(32)).astype(np.int32)[perm]
    bg_ids = (ref["accepted"] & np.uint64(0xffffffff)).astype(np.int32)[perm]
    ctx = coupling.make_context(bg, im, boxes=True)
    ctx.assemble_from_quadrature(torch.from_numpy(im_ids), torch.from_numpy(bg_ids), torch.from_numpy(off),
                                 torch.from_numpy(np.ascontiguousarray(ref["points"][idx])), torch.from_numpy(np.ascontiguousarray(ref["weights"][idx])))
    rp, col, val = ctx.csr(transpose=False)
    assert np.array_equal(rp.numpy().astype(np.uint64), ref["rowptr"])
    assert np.array_equal(col.numpy().astype(np.uint32), ref["col"])
    err = np.linalg.norm(val.numpy() - ref["val"]) / np.linalg.norm(ref["val"])
    assert err <= REL_FROB_TOL, err
    ctx.close()
