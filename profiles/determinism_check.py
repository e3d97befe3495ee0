import sys, torch
sys.path.insert(0,'/root/repo')
from non_matching_test_suite_b200 import coupling
from non_matching_test_suite_b200.grids import make_config
bg, im = make_config("sphere", 5, band_only=True, device="cuda")
vals=[]
for rep in range(4):
    ctx = coupling.make_context(bg, im, on_device=True, boxes=True)
    ctx.intersect(3,1e-15); rp,col,val = ctx.csr(device="cuda"); rpt,colt,valt = ctx.csr(transpose=True, device="cuda")
    z = ctx.spmv(torch.ones(bg.n_dofs, dtype=torch.float64, device="cuda"), transpose=True)
    vals.append((val.clone(), valt.clone(), z.clone(), col.clone())); ctx.close()
for k in range(1,4):
    print(k, torch.equal(vals[0][0], vals[k][0]), torch.equal(vals[0][1], vals[k][1]), torch.equal(vals[0][2], vals[k][2]), torch.equal(vals[0][3], vals[k][3]), float((vals[0][0]-vals[k][0]).abs().max()))
