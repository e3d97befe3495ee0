"""Shared by the tests that replay dumps of the REAL reference (oracle/reference_dump/README.md): loading a
tests/golden/ref_*.npz, turning it into the grids this repository works on, and -- to keep the converter and the tests
exercised while no such file exists -- writing a dump in the same text format from the repository's own oracle."""
import glob
import os

import numpy as np
import torch

from non_matching_test_suite_b200.grids import BackgroundGrid, ImmersedGrid

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def available():
    return sorted(glob.glob(os.path.join(GOLD, "ref_*.npz")))


def grids_of(D):
    d = int(D["spacedim"])
    V = D["space_cells_verts"]
    lo, hi = V[:, 0, :], V[:, (1 << d) - 1, :]
    for v in range(1 << d):          # deal.II vertex order on an axis-aligned cell
        for k in range(d):
            assert np.array_equal(V[:, v, k], hi[:, k] if (v >> k) & 1 else lo[:, k]), "background cell is not an axis-aligned box"
    q1 = D["space_cells_dofs"].shape[1] == (1 << d)
    dofs = D["space_cells_dofs"] if q1 else np.zeros((V.shape[0], 1 << d), dtype=np.uint32)
    bg = BackgroundGrid(d, torch.from_numpy(np.ascontiguousarray(lo)), torch.from_numpy(np.ascontiguousarray(hi)),
                        torch.from_numpy(dofs.astype(np.int32)), int(D["space_cells_n_dofs"]) if q1 else 1, torch.zeros(V.shape[0], dtype=torch.int32),
                        torch.from_numpy(D["c_dof"].astype(np.int32)) if q1 else torch.empty(0, dtype=torch.int32),
                        torch.from_numpy(D["c_ptr"].astype(np.int64)) if q1 else torch.zeros(1, dtype=torch.int64),
                        torch.from_numpy(D["c_master"].astype(np.int32)) if q1 else torch.empty(0, dtype=torch.int32),
                        torch.from_numpy(D["c_weight"].astype(np.float64)) if q1 else torch.empty(0, dtype=torch.float64))
    p0 = D["immersed_cells_dofs"].shape[1] == 1
    im = ImmersedGrid(d - 1, d, torch.from_numpy(np.ascontiguousarray(D["immersed_cells_verts"])),
                      torch.from_numpy((D["immersed_cells_dofs"] if p0 else np.arange(V.shape[0] * 0 + D["immersed_cells_verts"].shape[0])[:, None]).astype(np.int32)),
                      int(D["immersed_cells_n_dofs"]) if p0 else int(D["immersed_cells_verts"].shape[0]))
    return bg, im


def matrix_of(D):
    import scipy.sparse as sp
    A = sp.coo_matrix((D["vals"], (D["rows"].astype(np.int64), D["cols"].astype(np.int64))),
                      shape=(int(D["space_cells_n_dofs"]), int(D["immersed_cells_n_dofs"]))).tocsr()
    A.sort_indices()
    return A


def write_self_dump(path, bg, im, res, degree, tol=1e-15):
    """A dump in dump_reference.cc's format, filled from the repository's own oracle result `res`
    (c_oracle.assemble_with_quadrature): exercises tests/golden/ref_to_npz.py and the replay tests."""
    d = bg.spacedim
    V, Dd = bg.vertices().numpy(), bg.dofs.numpy()
    with open(path, "w") as f:
        f.write("spacedim %d degree %d tol %r fe 1 0\n" % (d, degree, tol))
        sp1 = [[(v >> k) & 1 for k in range(d)] for v in range(1 << d)]
        f.write("space_fe %d\n" % (1 << d) + "".join(" ".join(repr(float(x)) for x in p) + "\n" for p in sp1))
        f.write("immersed_fe 1\n" + " ".join(["0.5"] * (d - 1)) + "\n")
        f.write("space_cells %d %d\n" % (bg.n_cells, bg.n_dofs))
        for c in range(bg.n_cells):
            f.write(" ".join(repr(float(x)) for x in V[c].ravel()) + " " + " ".join(str(int(x)) for x in Dd[c]) + "\n")
        f.write("immersed_cells %d %d\n" % (im.n_cells, im.n_dofs))
        IV, ID = im.verts.numpy(), im.dofs.numpy()
        for c in range(im.n_cells):
            f.write(" ".join(repr(float(x)) for x in IV[c].ravel()) + " %d\n" % ID[c, 0])
        cd, cp, cm, cw = (t.numpy() for t in (bg.c_dof, bg.c_ptr, bg.c_master, bg.c_weight))
        f.write("constraints %d\n" % len(cd))
        for i in range(len(cd)):
            f.write("%d %d " % (cd[i], cp[i + 1] - cp[i]) + " ".join("%d %r" % (cm[k], float(cw[k])) for k in range(cp[i], cp[i + 1])) + "\n")
        acc = res["accepted"]
        f.write("pairs %d\n" % len(acc))
        off, pts, wts = res["q_offset"], res["points"], res["weights"]
        for i, key in enumerate(acc):
            a, b = int(off[i]), int(off[i + 1])
            f.write("%d %d %r %d " % (int(key) & 0xffffffff, int(key) >> 32, float(res["sum_w"][i]), b - a) +
                    " ".join(" ".join(repr(float(x)) for x in pts[k]) + " " + repr(float(wts[k])) for k in range(a, b)) + "\n")
        rp, col, val = res["rowptr"], res["col"], res["val"]
        f.write("matrix %d\n" % len(col))
        for r in range(bg.n_dofs):
            for k in range(int(rp[r]), int(rp[r + 1])):
                f.write("%d %d %r\n" % (r, int(col[k]), float(val[k])))
