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

    def owned_range(self, rank: Optional[int] = None):
        """Background dofs owned by `rank` when the result of Ct.vmult is distributed: contiguous ranges of
        ceil(m / world) dofs, the layout of a linear Epetra map."""
        return owned_dof_range(self.m, dist.get_rank(self.group) if rank is None and dist.is_initialized() else (rank or 0), self.world)

    def vmult_owned(self, lam: torch.Tensor) -> torch.Tensor:
        """The owned slice of y = A lam (reference: Ct*lambda lands in a vector distributed by dof ownership):
        reduce-scatter instead of all-reduce.  NCCL reduce_scatter needs equal chunks, so the partial vector is padded
        to world * rows_per_rank; gloo (CPU tests) has no reduce_scatter and takes the slice of an all-reduce."""
        y = self.local_vmult(lam)
        if self.world == 1:
            return y
        lo, hi = self.owned_range()
        rpr = rows_per_rank(self.m, self.world)
        if y.is_cuda:
            pad = torch.zeros(rpr * self.world, dtype=y.dtype, device=y.device)
            pad[:self.m] = y
            out = torch.empty(rpr, dtype=y.dtype, device=y.device)
            dist.reduce_scatter_tensor(out, pad, op=dist.ReduceOp.SUM, group=self.group)
            return out[:hi - lo]
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=self.group)
        return y[lo:hi].clone()

    def Tvmult(self, u: torch.Tensor) -> torch.Tensor:
        """z = A^T u  (reference: C.Tvmult).  A rank's result is non-zero only in its own immersed DoFs,
        so the replicated vector is the element-wise sum over ranks."""
        z = self.local_Tvmult(u)
        if self.world > 1:
            dist.all_reduce(z, op=dist.ReduceOp.SUM, group=self.group)
        return z


def rows_per_rank(n_dofs: int, world: int) -> int:
    return max(1, -(-n_dofs // world))


def owned_dof_range(n_dofs: int, rank: int, world: int):
    rpr = rows_per_rank(n_dofs, world)
    lo = min(n_dofs, rank * rpr)
    hi = n_dofs if rank == world - 1 else min(n_dofs, lo + rpr)
    return lo, hi


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

    def __init__(self, ctx, group=None, prefer_multicast: bool = True, device_barrier: bool = True):
        self.ctx, self.group = ctx, group
        self.available, self.reason, self.multicast = False, "", False
        self.sync_kind = "host barriers (NCCL) around the kernel"
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
            # the library launches on its own stream: zeroing, the rank barriers and the kernel are ordered on that
            # stream, so no host synchronisation sits between them
            self.stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
            self._dev_barrier = bool(device_barrier and hasattr(self.hdl, "barrier"))
            if self._dev_barrier:
                self.sync_kind = "device-side rank barriers in symmetric memory, stream-ordered with the kernel"
            self.y.zero_()
            torch.cuda.synchronize()
            self.available = True
        except Exception as e:  # no symmetric memory on this system: keep NCCL
            self.reason = "%s: %s" % (type(e).__name__, str(e)[:200])

    def _barrier(self):
        """All ranks have reached this point of the library stream.  Device side (signal pads of the symmetric-memory
        handle) when the handle offers it; otherwise the host waits for the stream and the ranks meet in NCCL."""
        if self._dev_barrier:
            try:
                self.hdl.barrier(channel=0)
                return
            except Exception:     # API not usable on this build: the same on every rank, so all fall back together
                self._dev_barrier = False
                self.sync_kind = "host barriers (NCCL) around the kernel"
        torch.cuda.current_stream().synchronize()
        dist.barrier(group=self.group)

    def vmult_owned(self, lam: torch.Tensor) -> torch.Tensor:
        """The slice of y = A lam this rank owns (contiguous dof ranges, `owned_dof_range`): every row result crosses
        NVLink once, to its owner, instead of once per rank (nm_spmv_push_owned).  Returns a view of the symmetric
        buffer.  Only the owned slice is zeroed and valid."""
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        lo, hi = owned_dof_range(self.ctx.n_space_dofs, rank, world)
        torch.cuda.current_stream().synchronize()          # lam is ready
        with torch.cuda.stream(self.stream):
            self.y[lo:hi].zero_()
            self._barrier()                                # every owner has cleared its slice
            self.ctx.spmv_push_owned(lam, self.ptrs, rows_per_rank(self.ctx.n_space_dofs, world), sync_producer=False)
            self._barrier()                                # every rank's adds have been issued and have landed
        torch.cuda.current_stream().wait_stream(self.stream)
        return self.y[lo:hi]

    def _pull_buffers(self, n_rows: int):
        """Send buffers every peer can read (symmetric memory): row results, their dof ids, the owner offsets.  Sized once,
        collectively, for 1.25 x the largest row count of any rank."""
        if getattr(self, "_pull", None) is not None:
            if n_rows > self._pull["cap"]:
                raise RuntimeError("the matrix has more stored rows (%d) than the send buffers hold (%d): create a new FusedCtVmult" %
                                   (n_rows, self._pull["cap"]))
            return self._pull
        import torch.distributed._symmetric_memory as symm_mem
        dev = torch.device("cuda", self.ctx.device)
        grp = self.group if self.group is not None else dist.group.WORLD
        cap = torch.tensor([n_rows], dtype=torch.int64, device=dev)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=self.group)
        cap = int(cap.item()) * 5 // 4 + 1024
        bufs = {}
        for name, n, dt in (("vals", cap, torch.float64), ("ids", cap, torch.int32), ("off", 32, torch.int64)):
            t = symm_mem.empty(n, dtype=dt, device=dev)
            h = symm_mem.rendezvous(t, grp)
            bufs[name] = (t, h, [int(p) for p in h.buffer_ptrs])
        self._pull = dict(cap=cap, **bufs)
        return self._pull

    def vmult_owned_pull(self, lam: torch.Tensor) -> torch.Tensor:
        """The owned slice of y = A lam by pulling (nm_spmv_rows / nm_owner_offsets / nm_pull_rows): every rank leaves the
        results of its stored rows and their dof ids in a send buffer; every owner adds the segment of each peer's buffer
        that falls into its dof range, source rank after source rank -- bulk reads over NVLink and plain adds in a fixed
        order, so the result is bit-reproducible, unlike the remote atomics of vmult_owned."""
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        lo, hi = owned_dof_range(self.ctx.n_space_dofs, rank, world)
        import os, time
        trace = os.environ.get("NMGPU_PULL_TRACE") == "1" and rank == 0
        t0 = time.perf_counter()
        marks = []

        def mark(label):
            nonlocal t0
            if trace:
                torch.cuda.synchronize()
                marks.append("%s %.2f" % (label, (time.perf_counter() - t0) * 1e3))
                t0 = time.perf_counter()
        # rows with entries: in compact-row mode (every shard of a multi-GPU run) that is every stored row
        self.ctx.build_sparsity()                          # (a no-op when the matrix is there)
        n_rows = self.ctx.info("stored_rows") if self.ctx.info("compact_rows") else self.ctx.n_rows_local()
        P = self._pull_buffers(n_rows)
        torch.cuda.current_stream().synchronize()          # lam is ready
        mark("n_rows")
        with torch.cuda.stream(self.stream):
            self.ctx.spmv_rows(lam, P["vals"][0], P["ids"][0])
            mark("spmv_rows")
            self.ctx.owner_offsets(world, rows_per_rank(self.ctx.n_space_dofs, world), P["off"][0])
            self.y[lo:hi].zero_()
            mark("offsets+zero")
            self._barrier()                                # every send buffer is complete
            mark("barrier")
            for src in range(world):                       # fixed order
                self.ctx.pull_rows(P["off"][2][src], rank, P["ids"][2][src], P["vals"][2][src], lo, self.y[lo:hi])
                mark("pull%d" % src)
            self._barrier()                                # every owner has read my buffer: it may be overwritten
            mark("barrier")
        torch.cuda.current_stream().wait_stream(self.stream)
        if trace:
            import sys
            sys.stderr.write("pull trace: " + " | ".join(marks) + "\n")
        return self.y[lo:hi]

    def vmult(self, lam: torch.Tensor) -> torch.Tensor:
        """Replicated y = A lam on every rank.  Two rank barriers bracket the kernel: all vectors are zero
        before anyone adds, and everyone's adds have landed before anyone reads."""
        torch.cuda.current_stream().synchronize()
        with torch.cuda.stream(self.stream):
            self.y.zero_()
            self._barrier()
            self.ctx.spmv_push(lam, self.targets, multicast=self.multicast, sync_producer=False)
            self._barrier()
        torch.cuda.current_stream().wait_stream(self.stream)
        return self.y
