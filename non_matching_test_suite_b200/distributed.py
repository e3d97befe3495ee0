"""Multi-GPU host logic: one process per GPU, immersed-cell ranges per rank.

The immersed space is discontinuous, so a rank owns the columns of the stored matrix that
belong to its immersed cells: candidate search, clipping, sparsity and values need no
collective.  Only `Ct.vmult` (stored matrix x lambda -> background vector) produces partial
sums that are all-reduced; `C.Tvmult` results are disjoint per rank.

The local products are injected as callables so that the same class drives the GPU context
(NCCL) and, in the CPU tests, a scipy matrix (gloo).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


class ShardedCouplingOperator:
    def __init__(self, local_vmult: Callable, local_Tvmult: Callable, n_space_dofs: int, n_immersed_dofs: int,
                 group: Optional[dist.ProcessGroup] = None):
        self.local_vmult, self.local_Tvmult = local_vmult, local_Tvmult
        self.m, self.n = n_space_dofs, n_immersed_dofs
        self.group = group

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def vmult(self, lam: torch.Tensor) -> torch.Tensor:
        """y = A lam  (reference: Ct.vmult).  Each rank multiplies its own columns; one all-reduce(sum)."""
        y = self.local_vmult(lam)
        if self.world > 1:
            dist.all_reduce(y, op=dist.ReduceOp.SUM, group=self.group)
        return y

    def Tvmult(self, u: torch.Tensor) -> torch.Tensor:
        """z = A^T u  (reference: C.Tvmult).  A rank's result is non-zero only in its own immersed DoFs,
        so the replicated vector is the element-wise sum over ranks."""
        z = self.local_Tvmult(u)
        if self.world > 1:
            dist.all_reduce(z, op=dist.ReduceOp.SUM, group=self.group)
        return z


def from_context(ctx, group=None) -> ShardedCouplingOperator:
    return ShardedCouplingOperator(lambda x: ctx.spmv(x, transpose=False), lambda u: ctx.spmv(u, transpose=True),
                                   ctx.n_space_dofs, ctx.n_im_dofs, group)
