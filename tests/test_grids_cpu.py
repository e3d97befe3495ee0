"""Synthetic grid generator: mirrors the reference drivers' recipe (sizes in SURVEY.md section 6)."""
import numpy as np
import pytest
import torch

from non_matching_test_suite_b200.grids import make_config, shard_problem, shard_range


@pytest.mark.parametrize("name,n_bg,n_im", [("disk", 472, 32), ("flower", 2434, 256), ("sphere", 5160, 96)])
def test_cycle0_sizes(name, n_bg, n_im):
    bg, im = make_config(name, 0, band_only=False)
    assert (bg.n_cells, im.n_cells) == (n_bg, n_im)
    # cells tile [-1,1]^d
    vol = (bg.hi - bg.lo).prod(dim=1).sum().item()
    assert abs(vol - 2.0 ** bg.spacedim) < 1e-12
    # constraint lines: hanging-node weights sum to one, masters are unconstrained
    cp = bg.c_ptr.numpy()
    w = bg.c_weight.numpy()
    sums = np.add.reduceat(np.append(w, 0.0), cp[:-1])[np.diff(cp) > 0]
    # a line keeps weight 1 unless some of its masters were Dirichlet rows (resolved away on close())
    assert set(np.unique(sums)) <= {0.25, 0.5, 0.75, 1.0}
    assert (sums == 1.0).mean() > 0.9
    assert not np.isin(bg.c_master.numpy(), bg.c_dof.numpy()).any()
    assert set(np.unique(w)) <= {0.25, 0.5}


def test_hanging_node_is_the_mean_of_its_masters():
    bg, _ = make_config("sphere", 0, band_only=False)
    X = bg.dof_coords.numpy()
    cp, cm, cw, cd = bg.c_ptr.numpy(), bg.c_master.numpy(), bg.c_weight.numpy(), bg.c_dof.numpy()
    for i in range(0, len(cd), 7):
        if cp[i + 1] > cp[i]:
            assert np.allclose((X[cm[cp[i]:cp[i + 1]]] * cw[cp[i]:cp[i + 1], None]).sum(0), X[cd[i]])


def test_band_only_is_a_subset_with_the_same_candidates(c_oracle):
    full, im = make_config("disk", 3, band_only=False)
    band, _ = make_config("disk", 3, band_only=True)
    assert band.n_cells < full.n_cells
    a = c_oracle.assemble(full, im)
    b = c_oracle.assemble(band, im)
    assert a["accepted"].shape == b["accepted"].shape
    assert np.allclose(np.sort(a["sum_w"]), np.sort(b["sum_w"]), rtol=1e-13)
    assert abs(a["val"].sum() - b["val"].sum()) < 1e-13


def test_shards_partition_the_pairs(c_oracle):
    bg, im, _ = shard_problem("sphere", 2, 0, 1)
    full = c_oracle.assemble(bg, im)
    tot = 0
    for r in range(3):
        b, i, _ = shard_problem("sphere", 2, r, 3)
        assert b.n_dofs == bg.n_dofs
        tot += c_oracle.assemble(b, i)["accepted"].shape[0]
    assert tot == full["accepted"].shape[0]
    assert [shard_range(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
