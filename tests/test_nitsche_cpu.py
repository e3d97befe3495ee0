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
