"""GPU parity of the higher-degree path (nm_set_finite_elements / nm_assemble_fe: FE_Q(p) background, FE_DGQ(q) immersed;
reference code/include/coupling_utilities.h:238-289) against oracle/fe_oracle.py.

Bars: pattern bit-exact; values within 1e-12 relative Frobenius norm wherever the comparison is on identical quadrature
points or the rule is exact for the integrand (every 2D case with the drivers' n = 2p + 1 points; Q1 x DGQ1 on planar
parallelogram cells in 3D).  On warped quads in 3D the reference's own result depends on its subdivision (the 7-point
rule is not exact for psi(eta(x))); there the library's own quadrature is only required to agree to the quadrature error."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from non_matching_test_suite_b200 import _lib, coupling
from non_matching_test_suite_b200.grids import BackgroundGrid, ImmersedGrid, make_config, circle_grid, sphere_grid
from oracle import fe_oracle as fo
from fe_helpers import uniform_background, continuous_dofs, discontinuous_dofs, random_constraint_lines, t

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _pairs(acc):
    return (acc >> np.uint64(32)).astype(np.int64), (acc & np.uint64(0xffffffff)).astype(np.int64)


def _bg(lo, hi, d):
    e = torch.empty(0, dtype=torch.int32)
    return BackgroundGrid(d, torch.from_numpy(lo), torch.from_numpy(hi), torch.zeros(lo.shape[0], 1 << d, dtype=torch.int32), 1,
                          torch.zeros(lo.shape[0], dtype=torch.int32), e, torch.zeros(1, dtype=torch.int64), e, torch.empty(0, dtype=torch.float64))


def _gpu_matrix(ctx, n_space, n_im):
    rp, col, val = ctx.csr()
    return sp.csr_matrix((val.numpy(), col.numpy(), rp.numpy()), shape=(n_space, n_im))


def _same(A, B, tol=TOL):
    A.sort_indices(); B.sort_indices()
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)      # pattern incl. kept zeros
    assert np.linalg.norm(A.data - B.data) <= tol * np.linalg.norm(B.data)


@pytest.mark.parametrize("name,cycle", [("disk", 1), ("flower", 0), ("sphere", 0), ("sphere", 1)])
def test_q1_p0_through_the_general_path_is_the_same_matrix(name, cycle):
    """(p, q) = (1, 0) with the vertex dofs and hanging-node lines of the configs: the general path must rebuild the matrix
    of the optimised path (same pattern, values to rounding)."""
    bg, im = make_config(name, cycle, band_only=False)
    d = bg.spacedim
    ctx = coupling.make_context(bg, im, boxes=True)
    ctx.intersect(3, 1e-15)
    A = _gpu_matrix(ctx, bg.n_dofs, im.n_dofs)
    ctx.set_finite_elements((1, fo.unit_support_points(d, 1)), bg.dofs.to(torch.int32).contiguous(), bg.n_dofs,
                            (0, fo.unit_support_points(d - 1, 0)), im.dofs.to(torch.int32).contiguous(), im.n_dofs,
                            bg.c_dof.to(torch.int32).contiguous(), bg.c_ptr.contiguous(), bg.c_master.to(torch.int32).contiguous(), bg.c_weight.contiguous())
    ctx.assemble_fe()
    assert ctx.info("merge_path") == 2
    _same(_gpu_matrix(ctx, bg.n_dofs, im.n_dofs), A)
    ctx.close()


@pytest.mark.parametrize("p,q,order", [(1, 1, "lexicographic"), (2, 0, "lexicographic"), (2, 1, "shuffled"), (2, 2, "lexicographic"), (3, 1, "shuffled"), (3, 3, "lexicographic")])
def test_2d_degrees_against_the_oracle(c_oracle, p, q, order):
    """Own quadrature of the library (one sub-segment per pair) against the oracle on the reference's subdivision, with
    the drivers' rule n = 2p + 1 (capped at 8 points: exact up to degree 15 >= 2p + q), made-up constraint lines, and
    both dof orders."""
    lo, hi = uniform_background(2, 12)
    im = circle_grid(56)
    bg = _bg(lo, hi, 2)
    n1d = min(2 * p + 1, 8)
    r = c_oracle.assemble_with_quadrature(bg, im, n1d, 1e-15)
    pim, pbg = _pairs(r["accepted"])
    sb, si = fo.unit_support_points(2, p, order, 3), fo.unit_support_points(1, q, order, 4)
    bd, ns = continuous_dofs(lo, hi, sb)
    idf, ni = discontinuous_dofs(im.n_cells, si.shape[0], seed=5 if order == "shuffled" else None)
    cd, cp, cm, cw = random_constraint_lines(ns, max(3, ns // 9), seed=p * 10 + q)
    want = fo.coupling_matrix(lo, hi, bd, ns, cd, cp, cm, cw, im.verts.numpy(), idf, ni, pim, pbg, r["q_offset"], r["points"], r["weights"], sb, si)
    ctx = _lib.Context(0)
    ctx.set_background(2, lo=torch.from_numpy(lo), hi=torch.from_numpy(hi))
    ctx.set_immersed(1, 2, im.verts.contiguous(), None, 0)
    c = ctx.intersect(n1d, 1e-15)
    assert c["n_accepted"] == len(pim)
    ctx.set_finite_elements((p, sb), t(bd), ns, (q, si), t(idf), ni, t(cd), t(cp), t(cm), t(cw))
    nnz = ctx.assemble_fe()
    assert nnz == want.nnz
    got = _gpu_matrix(ctx, ns, ni)
    _same(got, want)
    # the products and the local numbering (cell, local dof)
    g = np.random.default_rng(0)
    x, u = g.random(ni), g.random(ns)
    close = lambda a, b: np.abs(a - b).max() <= 1e-12 * np.abs(b).max()     # higher-degree bases change sign: entries of A x cancel
    assert close(ctx.spmv(torch.from_numpy(x), transpose=False).numpy(), want @ x)
    assert close(ctx.spmv(torch.from_numpy(u), transpose=True).numpy(), want.T @ u)
    z_loc = ctx.spmv_local(torch.from_numpy(u[ctx.csr_compact()[0].long().numpy()]), transpose=True)
    assert z_loc.numel() == im.n_cells * si.shape[0]
    rows = ctx.csr_compact()[0].long().numpy()
    u0 = np.zeros(ns); u0[rows] = u[rows]
    assert close(z_loc.numpy(), (want.T @ u0)[idf.reshape(-1).astype(np.int64)])
    # the same from the caller's tuples (the reference's own signature)
    ctx.assemble_fe(t(pim.astype(np.uint32)), t(pbg.astype(np.uint32)), t(r["q_offset"].astype(np.uint64)), torch.from_numpy(r["points"]),
                    torch.from_numpy(r["weights"]))
    _same(_gpu_matrix(ctx, ns, ni), want)
    ctx.close()


def _planar_sheet(n):
    """n x n parallelogram cells of a tilted plane through the cube: eta(x) is affine, so Q1 x DGQ1 has degree 5 in the
    plane and the 7-point rule is exact."""
    o, a, b = np.array([-0.61, -0.43, -0.52]), np.array([1.3, 0.2, 0.55]), np.array([0.15, 1.1, 0.4])
    s = np.linspace(0, 1, n + 1)
    P = o[None, None] + s[:, None, None] * a[None, None] + s[None, :, None] * b[None, None]
    q = np.stack([P[:-1, :-1], P[1:, :-1], P[:-1, 1:], P[1:, 1:]], axis=2).reshape(-1, 4, 3)
    return ImmersedGrid(2, 3, torch.from_numpy(np.ascontiguousarray(q)), torch.arange(n * n, dtype=torch.int32)[:, None], n * n)


@pytest.mark.parametrize("p,q,sheet", [(1, 1, True), (2, 1, True), (1, 1, False), (2, 2, False)])
def test_3d_degrees_against_the_oracle(c_oracle, p, q, sheet):
    lo, hi = uniform_background(3, 6)
    im = _planar_sheet(5) if sheet else sphere_grid(3)
    bg = _bg(lo, hi, 3)
    r = c_oracle.assemble_with_quadrature(bg, im, 3, 1e-15)
    pim, pbg = _pairs(r["accepted"])
    sb, si = fo.unit_support_points(3, p, "shuffled", 1), fo.unit_support_points(2, q)
    bd, ns = continuous_dofs(lo, hi, sb)
    idf, ni = discontinuous_dofs(im.n_cells, si.shape[0])
    cd, cp, cm, cw = random_constraint_lines(ns, ns // 11, seed=7)
    want = fo.coupling_matrix(lo, hi, bd, ns, cd, cp, cm, cw, im.verts.numpy(), idf, ni, pim, pbg, r["q_offset"], r["points"], r["weights"], sb, si)
    ctx = _lib.Context(0)
    ctx.set_background(3, lo=torch.from_numpy(lo), hi=torch.from_numpy(hi))
    ctx.set_immersed(2, 3, im.verts.contiguous(), None, 0)
    ctx.intersect(3, 1e-15)
    ctx.set_finite_elements((p, sb), t(bd), ns, (q, si), t(idf), ni, t(cd), t(cp), t(cm), t(cw))
    # identical points: the reference's own signature, tuples in, matrix out
    ctx.assemble_fe(t(pim.astype(np.uint32)), t(pbg.astype(np.uint32)), t(r["q_offset"].astype(np.uint64)), torch.from_numpy(r["points"]),
                    torch.from_numpy(r["weights"]))
    _same(_gpu_matrix(ctx, ns, ni), want)
    # the library's own subdivision
    ctx.assemble_fe()
    got = _gpu_matrix(ctx, ns, ni)
    got.sort_indices()
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    err = np.linalg.norm(got.data - want.data) / np.linalg.norm(want.data)
    if sheet and p == 1:
        assert err <= TOL, err                       # degree 5 in the plane: exact, subdivision independent
    else:
        assert err <= 1e-2, err                      # quadrature error of the 7-point rule (integrand degree > 5 / not polynomial)
    ctx.close()


def test_error_behaviour_of_the_general_path():
    bg, im = make_config("disk", 0, band_only=False)
    ctx = coupling.make_context(bg, im, boxes=True)
    ctx.intersect(3, 1e-15)
    sp2 = fo.unit_support_points(2, 2)
    dofs9 = torch.zeros(bg.n_cells, 9, dtype=torch.int32)
    with pytest.raises(_lib.NmError) as e:       # 8 support points are not a tensor grid of degree 2
        ctx.set_finite_elements((2, sp2[:8]), dofs9[:, :8].contiguous(), 10, (0, np.zeros((1, 1))), im.dofs.to(torch.int32).contiguous(), im.n_dofs)
    assert e.value.code == 3
    ctx.set_finite_elements((2, sp2), dofs9, 10, (0, np.full((1, 1), 0.5)), im.dofs.to(torch.int32).contiguous(), im.n_dofs)
    with pytest.raises(_lib.NmError) as e:       # the general matrix is never built implicitly
        ctx.csr()
    assert e.value.code == 4
    bad = dofs9.clone(); bad[:, 3] = 99
    ctx.set_finite_elements((2, sp2), bad, 10, (0, np.full((1, 1), 0.5)), im.dofs.to(torch.int32).contiguous(), im.n_dofs)
    with pytest.raises(_lib.NmError) as e:       # dof table beyond the declared dofs
        ctx.assemble_fe()
    assert e.value.code == 1
    ctx.close()


def test_python_mirror_for_elements():
    """coupling.create_coupling_mass_matrix_for_elements with (p, q) = (1, 0) described by support points = the matrix of
    the optimised path; lagrange_element orders its points like the Q1 vertices of the grids."""
    bg, im = make_config("sphere", 1, band_only=False)
    info = coupling.collect_quadratures_on_overlapped_grids(bg, im, 3, 1e-15, with_quadrature=False)
    A = _gpu_matrix(info.ctx, bg.n_dofs, im.n_dofs)
    con = (bg.c_dof.to(torch.int32).contiguous(), bg.c_ptr.contiguous(), bg.c_master.to(torch.int32).contiguous(), bg.c_weight.contiguous())
    C = coupling.create_coupling_mass_matrix_for_elements(bg, im, info, coupling.lagrange_element(3, 1), bg.dofs, bg.n_dofs,
                                                          coupling.lagrange_element(2, 0), im.dofs, im.n_dofs, constraints=con)
    assert info.ctx.info("merge_path") == 2
    _same(_gpu_matrix(C.ctx, bg.n_dofs, im.n_dofs), A)
    info.ctx.close()
