"""The assembly and the Nitsche terms on a GRADED background (grids.graded_config: the refinement level changes across the
interface, hanging nodes on cut cells all along it) next to the plain band of the same cycle: what the constraint
expansion costs at scale.  Usage: python profiles/graded_timing.py [cycle] [stripe]  -> one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from non_matching_test_suite_b200 import _lib, coupling
from non_matching_test_suite_b200.grids import graded_config, make_config

cycle = int(sys.argv[1]) if len(sys.argv) > 1 else 6
stripe = int(sys.argv[2]) if len(sys.argv) > 2 else 1
out = {"config": "sphere", "cycle": cycle, "stripe": stripe}
for label, (bg, im) in (("plain_band", make_config("sphere", cycle, band_only=True, device="cuda")),
                        ("graded_band", graded_config("sphere", cycle, stripe=stripe, band_only=True, device="cuda"))):
    torch.cuda.empty_cache()          # the grid generators leave blocks in torch's caching allocator
    ctx = _lib.Context(0)
    for it in range(4):
        coupling.set_grids(ctx, bg, im, boxes=True, on_device=True)
        c = ctx.intersect(3, 1e-15)
        ctx.build_sparsity()
        ctx.assemble()
    ph = ctx.timings()
    # constrained dofs that sit on cut cells
    _, pb, _ = ctx.pairs(device="cuda")
    cut_dofs = torch.unique(bg.dofs[torch.unique(pb.long())].reshape(-1).long())
    n_on_cut = int(torch.isin(cut_dofs, bg.c_dof.long()).sum())
    off, pts, w = ctx.quadrature(device="cuda")
    coef = 2.0 + pts[:, 0]
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        (rp, col, val), _ = ctx.assemble_nitsche(pb, off, pts, w, coefficient=coef, penalty=10.0)
        torch.cuda.synchronize()
        t_nit = (time.perf_counter() - t0) * 1e3
    out[label] = {"background_cells": bg.n_cells, "pairs": c["n_accepted"], "constraint_lines": int(bg.c_dof.numel()),
                  "constrained_dofs_on_cut_cells": n_on_cut, "merge_ms": round(ph["merge"], 3), "clip_ms": round(ph["clip"], 3),
                  "transpose_ms": round(ph["transpose"], 3), "merge_path": ctx.info("merge_path"), "nnz": ctx.nnz,
                  "nitsche_matrix_ms": round(t_nit, 2), "nitsche_nnz": int(col.numel())}
    ctx.close()
    del bg, im
    torch.cuda.empty_cache()
print(json.dumps(out))
