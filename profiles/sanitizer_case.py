import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from non_matching_test_suite_b200 import coupling
from non_matching_test_suite_b200.grids import make_config
for name, cyc in [("disk", 1), ("flower", 0), ("sphere", 1)]:
    bg, im = make_config(name, cyc, band_only=False)
    ctx = coupling.make_context(bg, im)
    c = ctx.intersect(3, 1e-15)
    off, pts, w = ctx.quadrature()
    rp, col, val = ctx.csr()
    rpt, colt, valt = ctx.csr(transpose=True)
    y = ctx.spmv(torch.ones(im.n_dofs, dtype=torch.float64)); z = ctx.spmv(torch.ones(bg.n_dofs, dtype=torch.float64), transpose=True)
    aim, abg, sw = ctx.pairs()
    ctx.assemble_from_quadrature(aim, abg, off, pts, w)
    rp2, col2, val2 = ctx.csr()
    f = ctx.flag_overlapped_background(bg.n_cells)
    print(name, c, float(y.sum()), float(z.sum()), float((val - val2).abs().max()))
    ctx.close()
