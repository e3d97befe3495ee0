"""world_size-2 gloo test of the sharded path's host logic: immersed-cell ranges own disjoint columns,
Ct.vmult needs exactly one all-reduce, C.Tvmult results are disjoint."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import scipy.sparse as sp
    from non_matching_test_suite_b200.distributed import ShardedCouplingOperator
    from non_matching_test_suite_b200.grids import shard_problem
    from oracle import c_oracle as co
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bg, im, _ = shard_problem("disk", 2, rank, world)
    r = co.assemble(bg, im, threads=1)
    A = sp.csr_matrix((r["val"], r["col"].astype(np.int64), r["rowptr"].astype(np.int64)), shape=(bg.n_dofs, im.n_dofs))
    op = ShardedCouplingOperator(lambda x: torch.from_numpy(A @ x.numpy()), lambda u: torch.from_numpy(A.T @ u.numpy()),
                                 bg.n_dofs, im.n_dofs)
    g = torch.Generator().manual_seed(0)
    lam = torch.rand(im.n_dofs, dtype=torch.float64, generator=g)
    u = torch.rand(bg.n_dofs, dtype=torch.float64, generator=g)
    y, z = op.vmult(lam), op.Tvmult(u)
    y_own = op.vmult_owned(lam)          # reduce-scatter semantics: the slice of the dofs this rank owns
    # columns owned by this rank only
    owned = np.unique(A.nonzero()[1])
    if rank == 0:
        torch.save({"y": y, "z": z, "owned": owned, "y_own": y_own, "range": op.owned_range()}, out)
    else:
        torch.save({"owned": owned, "y_own": y_own, "range": op.owned_range()}, out + ".1")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path, c_oracle):
    import scipy.sparse as sp
    from non_matching_test_suite_b200.grids import shard_problem
    out = str(tmp_path / "r0.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    other = torch.load(out + ".1", weights_only=False)
    assert len(np.intersect1d(got["owned"], other["owned"])) == 0      # disjoint C rows
    bg, im, _ = shard_problem("disk", 2, 0, 1)
    r = c_oracle.assemble(bg, im)
    A = sp.csr_matrix((r["val"], r["col"].astype(np.int64), r["rowptr"].astype(np.int64)), shape=(bg.n_dofs, im.n_dofs))
    g = torch.Generator().manual_seed(0)
    lam = torch.rand(im.n_dofs, dtype=torch.float64, generator=g)
    u = torch.rand(bg.n_dofs, dtype=torch.float64, generator=g)
    assert np.allclose(got["y"].numpy(), A @ lam.numpy(), rtol=1e-13, atol=1e-16)
    assert np.allclose(got["z"].numpy(), A.T @ u.numpy(), rtol=1e-13, atol=1e-16)
    # the owned slices tile the background vector and concatenate to the replicated result
    assert got["range"][0] == 0 and got["range"][1] == other["range"][0] and other["range"][1] == bg.n_dofs
    assert torch.equal(torch.cat([got["y_own"], other["y_own"]]), got["y"])


def test_owned_dof_ranges_tile_the_vector():
    from non_matching_test_suite_b200.distributed import owned_dof_range, rows_per_rank
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            rpr = rows_per_rank(n, world)
            prev = 0
            for r in range(world):
                lo, hi = owned_dof_range(n, r, world)
                assert lo == prev and lo <= hi <= n
                # the device kernel's owner rule: min(dof // rows_per_rank, world - 1)
                for dof in {lo, hi - 1} if hi > lo else set():
                    assert min(dof // rpr, world - 1) == r
                prev = hi
            assert prev == n
