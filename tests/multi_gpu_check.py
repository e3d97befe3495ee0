"""Run under torchrun on >= 2 GPUs:  the fused Ct.vmult (nm_spmv_push over NVLink peer memory / multicast)
must equal SpMV + NCCL all-reduce.  Prints one line per variant; exits non-zero on mismatch.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tests/multi_gpu_check.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from non_matching_test_suite_b200 import coupling, distributed as nmd  # noqa: E402
from non_matching_test_suite_b200.grids import shard_problem  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cycle = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    bg, im, info = shard_problem("sphere", cycle, rank, world, device=dev, weak=True, want_coords=False)
    torch.cuda.synchronize()
    ctx = coupling.make_context(bg, im, device=local, boxes=True, on_device=True)
    ctx.intersect(3, 1e-15)
    ctx.build_sparsity()
    lam = torch.rand(im.n_dofs, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
    ref_op = nmd.from_context(ctx)
    y_ref = ref_op.vmult(lam).clone()
    ok = True
    for prefer_mc in (False, True):
        f = nmd.FusedCtVmult(ctx, prefer_multicast=prefer_mc)
        if not f.available:
            if rank == 0:
                print("fused path unavailable:", f.reason, flush=True)
            continue
        y = f.vmult(lam)
        err = float((y - y_ref).abs().max() / y_ref.abs().max())
        # timing: NCCL path vs fused path
        def timed(fn, n=5):
            fn(); torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            torch.cuda.synchronize(); dist.barrier()
            return (time.perf_counter() - t0) / n * 1e3
        t_nccl = timed(lambda: ref_op.vmult(lam))
        t_fused = timed(lambda: f.vmult(lam))
        if rank == 0:
            print("world %d  n_space %d  multicast=%s  rel err %.2e  spmv+allreduce %.3f ms  fused push %.3f ms" % (
                world, bg.n_dofs, f.multicast, err, t_nccl, t_fused), flush=True)
        ok &= err < 1e-12
        if not prefer_mc:
            # reduce-scatter semantics: the owned slice of the fused push and of SpMV + NCCL reduce_scatter
            lo, hi = nmd.owned_dof_range(bg.n_dofs, rank, world)
            y_own = f.vmult_owned(lam).clone()
            y_rs = ref_op.vmult_owned(lam)
            scale = float(y_ref.abs().max())
            e_own = float((y_own - y_ref[lo:hi]).abs().max()) / scale if hi > lo else 0.0
            e_rs = float((y_rs - y_ref[lo:hi]).abs().max()) / scale if hi > lo else 0.0
            # the same slice by pulling (bulk reads of the peers' send buffers, adds in rank order): bit-reproducible
            y_pull = f.vmult_owned_pull(lam).clone()
            e_pull = float((y_pull - y_ref[lo:hi]).abs().max()) / scale if hi > lo else 0.0
            e_rep = 0.0 if torch.equal(f.vmult_owned_pull(lam), y_pull) else 1.0
            t_own = timed(lambda: f.vmult_owned(lam))
            t_pull = timed(lambda: f.vmult_owned_pull(lam))
            t_rs = timed(lambda: ref_op.vmult_owned(lam))
            errs = torch.tensor([e_own, e_rs, e_pull, e_rep], dtype=torch.float64, device=dev)
            dist.all_reduce(errs, op=dist.ReduceOp.MAX)
            if rank == 0:
                print("world %d  owned slices: rel err fused push %.2e  pull %.2e (repeat differs: %d)  nccl reduce_scatter %.2e   spmv+reduce_scatter %.3f ms  "
                      "fused owned push %.3f ms  pull %.3f ms" % (world, float(errs[0]), float(errs[2]), int(errs[3]), float(errs[1]), t_rs, t_own, t_pull), flush=True)
            ok &= float(errs.max()) < 1e-12
        if not f.multicast and prefer_mc:
            break
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
