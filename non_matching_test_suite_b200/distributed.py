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


class FusedCtVmult:
    """`Ct.vmult` + cross-GPU reduction in ONE kernel over NVLink peer memory (nm_spmv_push).

    The result vector lives in symmetric memory (torch.distributed._symmetric_memory): every rank maps
    the result vectors of all ranks, or -- when the fabric supports it -- one multicast address on which
    the NVSwitch applies the reduction in every rank's copy (`multimem.red`).  A rank only sends the rows
    it has entries in (its immersed cells touch ~1/N of the background), instead of all-reducing the
    whole background vector.  `available` is False when symmetric memory cannot be set up; callers then
    keep the NCCL all-reduce of ShardedCouplingOperator."""

    def __init__(self, ctx, group=None, prefer_multicast: bool = True):
        self.ctx, self.group = ctx, group
        self.available, self.reason, self.multicast = False, "", False
        try:
            import torch.distributed._symmetric_memory as symm_mem
            dev = torch.device("cuda", ctx.device)
            self.y = symm_mem.empty(ctx.n_space_dofs, dtype=torch.float64, device=dev)
            grp = group if group is not None else dist.group.WORLD
            self.hdl = symm_mem.rendezvous(self.y, grp)
            self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
            mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
            self.multicast = bool(prefer_multicast and mc)
            self.targets = [mc] if self.multicast else self.ptrs
            self.available = True
        except Exception as e:  # no symmetric memory on this system: keep NCCL
            self.reason = "%s: %s" % (type(e).__name__, str(e)[:200])

    def vmult(self, lam: torch.Tensor) -> torch.Tensor:
        """Replicated y = A lam on every rank.  Two rank barriers bracket the kernel: all vectors are zero
        before anyone adds, and everyone's adds have landed before anyone reads."""
        self.y.zero_()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self.ctx.spmv_push(lam, self.targets, multicast=self.multicast)
        dist.barrier(group=self.group)
        return self.y
