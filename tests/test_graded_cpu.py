"""Backgrounds whose refinement level changes ACROSS the interface (grids.graded_config: the drivers' adaptive space
refinement, left off in the shipped parameter files): hanging nodes on cut cells all along the interface, so the
constraint expansion of a9 / a11 (keep_constrained_entries, values sent to the masters) is exercised on many rows.
C oracle against the exact-rational oracle, and the identity that survives the expansion: column sums = cell measure."""
import numpy as np
import pytest
import scipy.sparse as sp

from non_matching_test_suite_b200.grids import graded_config
from oracle import exact_oracle as eo


@pytest.mark.parametrize("name,cycle,stripe", [("disk", 0, 2), ("disk", 1, 1), ("sphere", 0, 1)])
def test_c_oracle_equals_exact_oracle_on_graded_grids(c_oracle, name, cycle, stripe):
    bg, im = graded_config(name, cycle, stripe=stripe, band_only=False)
    r = c_oracle.assemble(bg, im)
    e = eo.assemble(bg, im)
    cut = np.unique((r["accepted"] & np.uint64(0xffffffff)).astype(np.int64))
    on_cut = np.intersect1d(np.unique(bg.dofs.numpy()[cut].ravel()), bg.c_dof.numpy())
    assert len(on_cut) >= 4                                    # the point of the grid: constrained dofs on cut cells
    A = sp.csr_matrix((r["val"], r["col"].astype(np.int64), r["rowptr"].astype(np.int64)), shape=(bg.n_dofs, im.n_dofs))
    C = A.tocoo()
    assert {(int(i), int(j)) for i, j in zip(C.row, C.col)} == e["pattern"]
    scale = max(abs(v) for v in e["values"].values())
    assert max(abs(e["values"].get((int(i), int(j)), 0.0) - v) for i, j, v in zip(C.row, C.col, C.data)) <= 1e-13 * scale
    # constrained rows are kept in the pattern and stay zero; the weights of a line add up to one, so every column still
    # sums to the measure of what the immersed cell cuts out
    assert np.abs(A[on_cut].toarray()).max() == 0.0 and A[on_cut].nnz > 0
    m = (r["accepted"] >> np.uint64(32)).astype(np.int64)
    per_cell = np.bincount(m, weights=r["sum_w"], minlength=im.n_cells)
    assert np.abs(np.asarray(A.sum(axis=0)).ravel()[im.dofs.numpy()[:, 0]] - per_cell).max() <= 1e-15
