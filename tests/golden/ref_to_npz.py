"""Turns the text dumps of oracle/reference_dump/dump_reference.cc (run inside the reference's Docker image) into
tests/golden/ref_*.npz.  usage: python ref_to_npz.py <dump dir> <golden dir>"""
import glob
import os
import sys

import numpy as np


def convert(path):
    tok = open(path).read().split()
    pos = [0]

    def take(n=1, conv=float):
        v = [conv(x) for x in tok[pos[0]:pos[0] + n]]
        pos[0] += n
        return v if n > 1 else v[0]
    assert take(1, str) == "spacedim"; d = take(1, int)
    assert take(1, str) == "degree"; degree = take(1, int)
    assert take(1, str) == "tol"; tol = take(1, float)
    assert take(1, str) == "fe"; p, q = take(2, int)
    fes = {}
    for tag, dim in (("space_fe", d), ("immersed_fe", d - 1)):
        assert take(1, str) == tag
        dpc = take(1, int)
        fes[tag] = np.array(take(dpc * dim), dtype=np.float64).reshape(dpc, dim) if dim * dpc > 1 else (np.array([take(dpc * dim)]).reshape(dpc, dim) if dim * dpc == 1 else np.zeros((dpc, 0)))
    out = dict(spacedim=d, degree=degree, tol=tol, p=p, q=q, space_support_points=fes["space_fe"], immersed_support_points=fes["immersed_fe"])
    for tag, nv, dpc in (("space_cells", 1 << d, fes["space_fe"].shape[0]), ("immersed_cells", 1 << (d - 1), fes["immersed_fe"].shape[0])):
        assert take(1, str) == tag
        n, ndofs = take(2, int)
        rec = np.array(take(n * (nv * d + dpc)), dtype=np.float64).reshape(n, nv * d + dpc)
        out[tag + "_verts"] = rec[:, :nv * d].reshape(n, nv, d)
        out[tag + "_dofs"] = rec[:, nv * d:].astype(np.uint32)
        out[tag + "_n_dofs"] = ndofs
    assert take(1, str) == "constraints"
    nc = take(1, int)
    c_dof, c_ptr, c_master, c_weight = [], [0], [], []
    for _ in range(nc):
        dof, nm = take(2, int)
        c_dof.append(dof)
        for _ in range(nm):
            c_master.append(take(1, int)); c_weight.append(take(1, float))
        c_ptr.append(len(c_master))
    out.update(c_dof=np.array(c_dof, dtype=np.uint32), c_ptr=np.array(c_ptr, dtype=np.uint64), c_master=np.array(c_master, dtype=np.uint32),
               c_weight=np.array(c_weight, dtype=np.float64))
    assert take(1, str) == "pairs"
    npairs = take(1, int)
    pbg, pim, sw, off, pts, wts = [], [], [], [0], [], []
    for _ in range(npairs):
        b, m = take(2, int); s = take(1, float); nq = take(1, int)
        rec = np.array(take(nq * (d + 1)), dtype=np.float64).reshape(nq, d + 1) if nq * (d + 1) > 1 else np.zeros((0, d + 1))
        pbg.append(b); pim.append(m); sw.append(s); pts.append(rec[:, :d]); wts.append(rec[:, d]); off.append(off[-1] + nq)
    out.update(pair_space=np.array(pbg, dtype=np.uint32), pair_immersed=np.array(pim, dtype=np.uint32), pair_sum_w=np.array(sw),
               q_offset=np.array(off, dtype=np.uint64), points=np.concatenate(pts) if pts else np.zeros((0, d)),
               weights=np.concatenate(wts) if wts else np.zeros(0))
    assert take(1, str) == "matrix"
    nnz = take(1, int)
    rec = np.array(take(nnz * 3), dtype=np.float64).reshape(nnz, 3)
    out.update(rows=rec[:, 0].astype(np.uint32), cols=rec[:, 1].astype(np.uint32), vals=rec[:, 2])
    return out


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    for f in sorted(glob.glob(os.path.join(src, "ref_*.txt"))):
        np.savez_compressed(os.path.join(dst, os.path.basename(f)[:-4] + ".npz"), **convert(f))
        print("wrote", os.path.basename(f)[:-4] + ".npz")
