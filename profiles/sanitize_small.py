#!/usr/bin/env python
"""Small driver for compute-sanitizer (memcheck / racecheck): every kernel of the step on four small grids -- background
with constraint lines (the merge that looks lines up), band-only backgrounds (the merge without lines), 2D and 3D -- plus
the products in global and local numbering.  usage: compute-sanitizer --tool racecheck python profiles/sanitize_small.py
(this round's GPU pool refuses to run compute-sanitizer, so only the plain run was made there: all four finish, finite)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from non_matching_test_suite_b200 import coupling
from non_matching_test_suite_b200.grids import make_config

g = torch.Generator().manual_seed(0)
for name, cycle, band in [("sphere", 1, False), ("sphere", 2, True), ("disk", 3, False), ("flower", 4, True)]:
    bg, im = make_config(name, cycle, band_only=band)
    ctx = coupling.make_context(bg, im, boxes=True)
    counts = ctx.intersect(3, 1e-15)
    rp, col, val = ctx.csr()
    y = ctx.spmv(torch.rand(im.n_dofs, dtype=torch.float64, generator=g), transpose=False)
    z = ctx.spmv(torch.rand(bg.n_dofs, dtype=torch.float64, generator=g), transpose=True)
    rows = ctx.csr_compact()[0].long()
    yl = ctx.spmv_local(torch.rand(im.n_cells, dtype=torch.float64, generator=g), transpose=False)
    zl = ctx.spmv_local(torch.rand(rows.numel(), dtype=torch.float64, generator=g), transpose=True)
    print(name, cycle, "band" if band else "full", counts["n_accepted"], "pairs, merge path", ctx.info("merge_path"), "sum", float(val.sum()),
          bool(torch.isfinite(y).all() and torch.isfinite(z).all() and torch.isfinite(yl).all() and torch.isfinite(zl).all()), flush=True)
    ctx.close()
print("done")
