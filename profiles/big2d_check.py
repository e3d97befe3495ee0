import sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
from non_matching_test_suite_b200 import coupling
from non_matching_test_suite_b200.grids import make_config
for name, cyc in [("flower", 14), ("disk", 14), ("disk", 17)]:
    t = time.time()
    bg, im = make_config(name, cyc, band_only=True, device="cuda")
    torch.cuda.synchronize(); tg = time.time() - t
    ctx = coupling.make_context(bg, im, on_device=True)
    for rep in range(3):
        t = time.time()
        coupling.set_grids(ctx, bg, im, on_device=True)
        c = ctx.intersect(3, 1e-15); ctx.build_sparsity(); ctx.assemble()
        x = torch.ones(im.n_dofs, dtype=torch.float64, device="cuda"); u = torch.ones(bg.n_dofs, dtype=torch.float64, device="cuda")
        y = ctx.spmv(x); z = ctx.spmv(u, transpose=True)
        torch.cuda.synchronize(); dt = time.time() - t
    aim, abg, sw = ctx.pairs(device="cuda")
    per_cell = torch.zeros(im.n_cells, dtype=torch.float64, device="cuda").index_add_(0, aim.long(), sw)
    meas = im.measure()
    err = float(((per_cell - meas).abs() / meas).max())
    pu = float(((z - per_cell).abs() / meas).max())
    print(name, cyc, "bg", bg.n_cells, "im", im.n_cells, "h_bg %.2e" % float((bg.hi - bg.lo).min()), c, "gen %.1fs step %.1f ms -> %.3g pairs/s" % (tg, dt * 1e3, c["n_accepted"] / dt),
          "max rel |sum_w - length| %.2e  partition-of-unity %.2e" % (err, pu), {k: round(v, 2) for k, v in ctx.timings().items() if v > 0.01}, flush=True)
    ctx.close()
