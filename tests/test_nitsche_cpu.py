"""Oracle of the interface-penalisation terms (oracle/nm_oracle.c::nmo_nitsche) against a dense numpy restatement of
code/include/coupling_utilities.h:306-481 on hand-made data, including the constraint algebra of
AffineConstraints::distribute_local_to_global (square and vector versions, [UPSTREAM] deal.II 9.4)."""
import numpy as np
import pytest
import torch

from non_matching_test_suite_b200.grids import BackgroundGrid


def _two_cells(d, constraints=True):
    """Two unit boxes side by side along x sharing a face; dof 0 hangs on (1, 2) with weights 1/2, the last dof is a
    Dirichlet dof (constraint line without masters)."""
    nl = 1 << d
    lo = torch.zeros(2, d, dtype=torch.float64); hi = torch.ones(2, d, dtype=torch.float64)
    lo[1, 0] = 1.0; hi[1, 0] = 2.0
    hi[:, -1] = 0.5                                        # anisotropic: h = diameter, not an edge length
    n_x = 3
    dofs = torch.zeros(2, nl, dtype=torch.int32)
    for c in range(2):
        for i in range(nl):
            ix = (i & 1) + c
            rest = i >> 1
            dofs[c, i] = ix + n_x * rest
    n_dofs = n_x * (nl // 2)
    if constraints:
        c_dof = torch.tensor([0, n_dofs - 1], dtype=torch.int32)
        c_ptr = torch.tensor([0, 2, 2], dtype=torch.int64)
        c_master = torch.tensor([1, 2], dtype=torch.int32)
        c_weight = torch.tensor([0.5, 0.5], dtype=torch.float64)
    else:
        c_dof = torch.empty(0, dtype=torch.int32); c_ptr = torch.zeros(1, dtype=torch.int64)
        c_master = torch.empty(0, dtype=torch.int32); c_weight = torch.empty(0, dtype=torch.float64)
    return BackgroundGrid(d, lo, hi, dofs, n_dofs, torch.zeros(2, dtype=torch.int32), c_dof, c_ptr, c_master, c_weight)


def _dense_reference(bg, ids, qo, pts, wts, coeff, fval, penalty):
    d = bg.spacedim; nl = 1 << d; n = bg.n_dofs
    A = np.zeros((n, n)); b = np.zeros(n)
    lines = {int(bg.c_dof[k]): [(int(bg.c_master[a]), float(bg.c_weight[a])) for a in range(int(bg.c_ptr[k]), int(bg.c_ptr[k + 1]))]
             for k in range(bg.c_dof.shape[0])}
    for p, cell in enumerate(ids):
        lo, hi = bg.lo[cell].numpy(), bg.hi[cell].numpy()
        h = np.sqrt(((hi - lo) ** 2).sum())
        L = np.zeros((nl, nl)); F = np.zeros(nl)
        for q in range(qo[p], qo[p + 1]):
            xi = (pts[q] - lo) / (hi - lo)
            phi = np.array([np.prod([xi[k] if (i >> k) & 1 else 1 - xi[k] for k in range(d)]) for i in range(nl)])
            L += coeff[q] * penalty / h * np.outer(phi, phi) * wts[q]
            F += coeff[q] * penalty / h * phi * fval[q] * wts[q]
        dofs = bg.dofs[cell].numpy()
        res = [lines.get(int(g), [(int(g), 1.0)]) if int(g) in lines else [(int(g), 1.0)] for g in dofs]
        for i in range(nl):
            for r, wr in res[i]:
                b[r] += wr * F[i]
                for j in range(nl):
                    for c, wc in res[j]:
                        A[r, c] += wr * wc * L[i, j]
            if int(dofs[i]) in lines:
                A[dofs[i], dofs[i]] += abs(L[i, i]) if L[i, i] != 0 else np.abs(np.diag(L)).mean()
    return A, b


@pytest.mark.parametrize("d", [2, 3])
@pytest.mark.parametrize("constraints", [False, True])
def test_oracle_nitsche_matches_dense_restatement(c_oracle, d, constraints):
    bg = _two_cells(d, constraints)
    rng = np.random.default_rng(5)
    ids = np.array([0, 1, 1, 0], dtype=np.uint32)
    nq = np.array([3, 7, 0, 5])
    qo = np.concatenate([[0], np.cumsum(nq)]).astype(np.uint64)
    pts = np.zeros((qo[-1], d))
    for p, cell in enumerate(ids):
        lo, hi = bg.lo[cell].numpy(), bg.hi[cell].numpy()
        pts[qo[p]:qo[p + 1]] = lo + rng.random((nq[p], d)) * (hi - lo)
    pts[0, 0] = 0.0                       # a point on the face x = 0: phi of the x = 1 vertices vanishes there
    wts = rng.random(qo[-1]) * 0.1
    coeff = 2.0 + rng.random(qo[-1]); coeff[3] = -1.5          # sign change: |local(i,i)| matters
    fval = rng.standard_normal(qo[-1])
    A, rhs = c_oracle.nitsche(bg, ids, qo, pts, wts, coeff, fval, penalty=10.0)
    Ad, bd = _dense_reference(bg, ids, qo.astype(np.int64), pts, wts, coeff, fval, 10.0)
    assert np.abs(A.toarray() - Ad).max() <= 1e-14 * np.abs(Ad).max()
    assert np.abs(rhs - bd).max() <= 1e-14 * np.abs(bd).max()
    if constraints:
        # the Dirichlet dof receives nothing but its diagonal; the hanging dof's row holds only its diagonal
        last = bg.n_dofs - 1
        assert A[last].nnz == 1 and A[last, last] > 0 and rhs[last] == 0
        assert A[0].nnz == 1 and A[0, 0] > 0 and rhs[0] == 0
    else:
        # partition of unity: sum_ij A_ij = sum_q c penalty / h w
        h = np.sqrt(((bg.hi - bg.lo).numpy() ** 2).sum(axis=1))
        per_q = np.repeat(h[ids], nq)
        assert abs(A.sum() - (coeff * 10.0 / per_q * wts).sum()) <= 1e-13 * abs(A).sum()
        assert abs(rhs.sum() - (coeff * 10.0 / per_q * wts * fval).sum()) <= 1e-13 * np.abs(rhs).sum()


def test_oracle_nitsche_zero_diagonal_uses_block_average(c_oracle):
    """set_matrix_diagonals [UPSTREAM]: a constrained dof whose local diagonal is exactly zero gets the mean |diagonal|."""
    bg = _two_cells(2, True)
    ids = np.array([0], dtype=np.uint32); qo = np.array([0, 1], dtype=np.uint64)
    pts = np.array([[1.0, 0.25]])          # on the face x = 1 of cell 0: phi_0 = phi_2 = 0, dof 0 is constrained
    wts = np.array([0.3])
    A, _ = c_oracle.nitsche(bg, ids, qo, pts, wts, None, None, penalty=4.0)
    h = np.sqrt(1 + 0.25)
    phi = np.array([0.0, 0.5, 0.0, 0.5])
    diag = 4.0 / h * phi * phi * 0.3
    assert A[0, 0] == pytest.approx(np.abs(diag).mean(), rel=1e-15)


@pytest.mark.parametrize("name,cycle", [("disk", 1), ("sphere", 0)])
def test_oracle_quadrature_in_the_reference_subdivision(c_oracle, name, cycle):
    """nmo_pairs_quadrature emits the points the oracle integrates with: per pair the weights sum to sum_w and the
    Q1 local vector recomputed from the points is the oracle's."""
    from non_matching_test_suite_b200.grids import make_config
    bg, im = make_config(name, cycle, band_only=False)
    r = c_oracle.assemble_with_quadrature(bg, im)
    d = bg.spacedim
    qo = r["q_offset"].astype(np.int64)
    assert qo[-1] == r["nq"].sum() and r["points"].shape == (qo[-1], d)
    sw = np.add.reduceat(r["weights"], qo[:-1])
    assert np.abs(sw - r["sum_w"]).max() <= 1e-15 * r["sum_w"].max()
    cells = (r["accepted"] & np.uint64(0xffffffff)).astype(np.int64)
    lo, hi = bg.lo.numpy()[cells], bg.hi.numpy()[cells]
    per_q = np.repeat(np.arange(len(cells)), np.diff(qo))
    xi = (r["points"] - lo[per_q]) / (hi[per_q] - lo[per_q])
    assert xi.min() >= -1e-12 and xi.max() <= 1 + 1e-12          # every point lies in its background cell
    for i in range(1 << d):
        phi = np.prod([xi[:, k] if (i >> k) & 1 else 1 - xi[:, k] for k in range(d)], axis=0)
        loc = np.add.reduceat(phi * r["weights"], qo[:-1])
        assert np.abs(loc - r["local"][:, i]).max() <= 1e-14 * np.abs(r["local"]).max()


def _clip_segment(p0, p1, lo, hi):
    """Liang-Barsky: parameter range of the segment inside the closed box, or None."""
    t0, t1 = 0.0, 1.0
    for k in range(2):
        dk = p1[k] - p0[k]
        for bound, sgn in ((lo[k], 1.0), (hi[k], -1.0)):
            a, b = sgn * dk, sgn * (bound - p0[k])          # a t >= b
            if a == 0:
                if b > 0:
                    return None
            elif a > 0:
                t0 = max(t0, b / a)
            else:
                t1 = min(t1, b / a)
    return (t0, t1) if t0 < t1 else None


def test_nitsche_2d_does_not_depend_on_the_subdivision(c_oracle):
    """Rule R1 for the Nitsche terms in 2D: phi_i phi_j has degree 4 along a segment and the 3-point rule is exact to
    degree 5, so the matrix built on the reference's subdivision (segment cut at the CDT diagonal of the background
    quad) equals the one built on the undivided clipped segment; same for a rhs with a linear boundary function."""
    from non_matching_test_suite_b200.grids import make_config
    bg, im = make_config("disk", 1, band_only=False)
    r = c_oracle.assemble_with_quadrature(bg, im)
    cells = (r["accepted"] & np.uint64(0xffffffff)).astype(np.int64)
    segs = (r["accepted"] >> np.uint64(32)).astype(np.int64)
    g = 0.5 + 0.5 * np.sqrt(3.0 / 5.0) * np.array([-1.0, 0.0, 1.0]); gw = np.array([5.0, 8.0, 5.0]) / 18.0
    pts, wts, qo = [], [], [0]
    for c, m in zip(cells, segs):
        p0, p1 = im.verts[m, 0].numpy(), im.verts[m, 1].numpy()
        tr = _clip_segment(p0, p1, bg.lo[c].numpy(), bg.hi[c].numpy())
        assert tr is not None
        a, b = p0 + tr[0] * (p1 - p0), p0 + tr[1] * (p1 - p0)
        pts.append(a[None, :] + g[:, None] * (b - a)[None, :]); wts.append(np.hypot(*(b - a)) * gw); qo.append(qo[-1] + 3)
    pts, wts, qo = np.concatenate(pts), np.concatenate(wts), np.array(qo, dtype=np.uint64)
    f_ref, f_one = 1.0 + r["points"][:, 0] - 2.0 * r["points"][:, 1], 1.0 + pts[:, 0] - 2.0 * pts[:, 1]
    A1, b1 = c_oracle.nitsche(bg, cells, r["q_offset"], r["points"], r["weights"], None, f_ref, 10.0)
    A2, b2 = c_oracle.nitsche(bg, cells, qo, pts, wts, None, f_one, 10.0)
    assert r["points"].shape[0] > pts.shape[0]                 # the reference's subdivision really has more pieces
    assert np.array_equal(A1.indptr, A2.indptr) and np.array_equal(A1.indices, A2.indices)
    assert np.linalg.norm(A1.data - A2.data) <= 1e-13 * np.linalg.norm(A1.data)
    assert np.linalg.norm(b1 - b2) <= 1e-13 * np.linalg.norm(b1)


def test_nitsche_2d_against_exact_rational_blocks(c_oracle):
    """The C oracle's Nitsche matrix on its own (reference-subdivision) quadrature against exact rational integration
    of phi_i phi_j over (segment n cell) (oracle/exact_oracle.py::nitsche_block_2d), with constant coefficient, pushed
    through the same constraint algebra in numpy: pins quadrature + formula + scatter of the 2D case."""
    from oracle import exact_oracle as eo
    from non_matching_test_suite_b200.grids import make_config
    bg, im = make_config("disk", 0, band_only=False)
    r = c_oracle.assemble_with_quadrature(bg, im)
    cells = (r["accepted"] & np.uint64(0xffffffff)).astype(np.int64)
    segs = (r["accepted"] >> np.uint64(32)).astype(np.int64)
    pen, c0 = 10.0, 2.0
    A, _ = c_oracle.nitsche(bg, cells, r["q_offset"], r["points"], r["weights"], np.full(r["weights"].shape[0], c0), None, pen)
    n = bg.n_dofs
    D = np.zeros((n, n))
    lines = {int(bg.c_dof[k]): [(int(bg.c_master[a]), float(bg.c_weight[a])) for a in range(int(bg.c_ptr[k]), int(bg.c_ptr[k + 1]))]
             for k in range(bg.c_dof.shape[0])}
    for c, m in zip(cells, segs):
        lo, hi = bg.lo[c].numpy(), bg.hi[c].numpy()
        h = np.sqrt(((hi - lo) ** 2).sum())
        L = c0 * pen / h * np.array(eo.nitsche_block_2d(lo, hi, im.verts[m].numpy()))
        dofs = bg.dofs[c].numpy()
        res = [lines[int(g)] if int(g) in lines else [(int(g), 1.0)] for g in dofs]
        for i in range(4):
            for rr, wr in res[i]:
                for j in range(4):
                    for cc, wc in res[j]:
                        D[rr, cc] += wr * wc * L[i, j]
            if int(dofs[i]) in lines:
                D[dofs[i], dofs[i]] += abs(L[i, i]) if L[i, i] != 0 else np.abs(np.diag(L)).mean()
    assert len(lines) > 0                                   # hanging-node lines take part
    assert np.abs(A.toarray() - D).max() <= 1e-13 * np.abs(D).max()
