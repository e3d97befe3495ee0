import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from non_matching_test_suite_b200 import _lib
from non_matching_test_suite_b200.grids import circle_grid
from oracle import fe_oracle as fo
from fe_helpers import uniform_background, continuous_dofs, discontinuous_dofs, random_constraint_lines, t
import scipy.sparse as sp
for (p, q) in [(3, 1), (3, 3), (2, 2)]:
    lo, hi = uniform_background(2, 12); im = circle_grid(56)
    sb, si = fo.unit_support_points(2, p), fo.unit_support_points(1, q)
    bd, ns = continuous_dofs(lo, hi, sb); idf, ni = discontinuous_dofs(im.n_cells, si.shape[0])
    cd, cp, cm, cw = random_constraint_lines(ns, max(3, ns // 9), seed=p * 10 + q)
    ctx = _lib.Context(0)
    ctx.set_background(2, lo=torch.from_numpy(lo), hi=torch.from_numpy(hi)); ctx.set_immersed(1, 2, im.verts.contiguous(), None, 0)
    ctx.intersect(min(2 * p + 1, 8), 1e-15)
    ctx.set_finite_elements((p, sb), t(bd), ns, (q, si), t(idf), ni, t(cd), t(cp), t(cm), t(cw))
    nnz = ctx.assemble_fe()
    rp, col, val = ctx.csr()
    A = sp.csr_matrix((val.numpy(), col.numpy(), rp.numpy()), shape=(ns, ni))
    x = np.random.default_rng(0).random(ni)
    y = ctx.spmv(torch.from_numpy(x), transpose=False).numpy()
    yd = ctx.spmv(torch.from_numpy(x).cuda(), transpose=False).cpu().numpy()
    print(p, q, "nnz", nnz, "rows", ns, "max|y-Ax| host %.3e dev %.3e  |Ax| %.3e  |y| %.3e" % (np.abs(y - A @ x).max(), np.abs(yd - A @ x).max(), np.abs(A @ x).max(), np.abs(y).max()),
          "info stored_rows", ctx.info("stored_rows"), "compact", ctx.info("compact_rows"), "avg", nnz / ns)
    ctx.close()
