"""The C++ drop-in header include/coupling_utilities.h (reference signatures over the C ABI).

CPU: it compiles against the deal.II stub (deal.II is not installed here) and links libnmgpu.so.
GPU: the compiled driver runs the reference's three calls and its matrix is compared with the oracle.
"""
import os
import subprocess

import numpy as np
import pytest

from non_matching_test_suite_b200.grids import make_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "test_dropin")


def build_driver():
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    lib = os.path.join(ROOT, "non_matching_test_suite_b200")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-pthread", "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "tests", "cpp", "stub"), os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp"), "-o", BIN,
           "-L" + lib, "-lnmgpu", "-Wl,-rpath," + lib]
    subprocess.check_call(cmd)


def write_grid(path, bg, im):
    d = bg.spacedim
    V, D = bg.vertices().numpy(), bg.dofs.numpy()
    with open(path, "w") as f:
        f.write("%d\n%d %d\n" % (22 if im.dim == d else d, bg.n_cells, bg.n_dofs))       # 22: the codimension-0 instantiation <2,2,2>
        for c in range(bg.n_cells):
            f.write(" ".join(repr(float(x)) for x in V[c].ravel()) + " " + " ".join(str(int(x)) for x in D[c]) + "\n")
        cd, cp, cm, cw = (t.numpy() for t in (bg.c_dof, bg.c_ptr, bg.c_master, bg.c_weight))
        f.write("%d\n" % len(cd))
        for i in range(len(cd)):
            f.write("%d %d " % (cd[i], cp[i + 1] - cp[i]) + " ".join("%d %r" % (cm[k], float(cw[k])) for k in range(cp[i], cp[i + 1])) + "\n")
        f.write("%d %d\n" % (im.n_cells, im.n_dofs))
        IV, ID = im.verts.numpy(), im.dofs.numpy()
        for c in range(im.n_cells):
            f.write(" ".join(repr(float(x)) for x in IV[c].ravel()) + " %d\n" % ID[c, 0])


def test_header_compiles_against_stub():
    build_driver()
    assert os.path.exists(BIN)


@pytest.mark.gpu
@pytest.mark.parametrize("name,cycle", [("disk", 0), ("sphere", 0)])
def test_cpp_driver_matches_oracle(tmp_path, c_oracle, name, cycle):
    build_driver()
    bg, im = make_config(name, cycle, band_only=False)
    grid, out = str(tmp_path / "grid.txt"), str(tmp_path / "out.txt")
    write_grid(grid, bg, im)
    subprocess.check_call([BIN, grid, out])
    ref = c_oracle.assemble(bg, im, n1d=3, tol=1e-15)
    lines = open(out).read().split("\n")
    n_pairs, area = lines[0].split()
    assert int(n_pairs) == ref["accepted"].shape[0]
    assert abs(float(area) - ref["sum_w"].sum()) <= 1e-13 * ref["sum_w"].sum()
    nnz = int(lines[1])
    assert nnz == ref["col"].shape[0]
    rows = np.repeat(np.arange(bg.n_dofs), np.diff(ref["rowptr"].astype(np.int64)))
    got = np.array([[float(x) for x in l.split()] for l in lines[2:2 + nnz]])
    assert np.array_equal(got[:, 0].astype(np.int64), rows)
    assert np.array_equal(got[:, 1].astype(np.uint32), ref["col"])
    assert np.linalg.norm(got[:, 2] - ref["val"]) <= 1e-12 * np.linalg.norm(ref["val"])
    # Nitsche matrix and rhs of the same driver against the oracle on the same quadrature (the library's own, which
    # is bitwise reproducible, collected again here)
    import scipy.sparse as sp
    from non_matching_test_suite_b200 import coupling
    info = coupling.collect_quadratures_on_overlapped_grids(bg, im, 3, 1e-15)
    A, b = c_oracle.nitsche(bg, info.space_cell, info.q_offset, info.points, info.weights, 2.0 + info.points[:, 0],
                            1.0 + info.points[:, 1] ** 2, 10.0)
    # rhs call of the driver used the constant coefficient 2
    _, b = c_oracle.nitsche(bg, info.space_cell, info.q_offset, info.points, info.weights, np.full(info.weights.shape[0], 2.0),
                            1.0 + info.points[:, 1] ** 2, 10.0)
    n_nit = int(lines[2 + nnz])
    ent = np.array([[float(x) for x in l.split()] for l in lines[3 + nnz:3 + nnz + n_nit]])
    G = sp.csr_matrix((ent[:, 2], (ent[:, 0].astype(np.int64), ent[:, 1].astype(np.int64))), shape=(bg.n_dofs, bg.n_dofs))
    A.eliminate_zeros()                      # matrix.add elides exact zeros
    assert abs(G - A).max() <= 1e-13 * abs(A).max()
    assert (G != 0).nnz == (A != 0).nnz
    rhs = np.array([float(x) for x in lines[3 + nnz + n_nit:3 + nnz + n_nit + bg.n_dofs]])
    assert np.linalg.norm(rhs - b) <= 1e-13 * np.linalg.norm(b)
    info.ctx.close()


@pytest.mark.gpu
def test_cpp_driver_codimension_zero(tmp_path, c_oracle):
    """The reference's three calls instantiated as <2,2,2> (a 2D body in the 2D background, code/src/coupling_utilities.cc:72-102)
    through the drop-in header against the oracle."""
    from non_matching_test_suite_b200.grids import solid_grid_2d
    build_driver()
    bg, _ = make_config("disk", 1, band_only=False)
    im = solid_grid_2d(12)
    grid, out = str(tmp_path / "grid.txt"), str(tmp_path / "out.txt")
    write_grid(grid, bg, im)
    subprocess.check_call([BIN, grid, out])
    ref = c_oracle.assemble(bg, im, n1d=3, tol=1e-15)
    lines = open(out).read().split("\n")
    n_pairs, area = lines[0].split()
    assert int(n_pairs) == ref["accepted"].shape[0]
    assert abs(float(area) - float(im.measure().sum())) <= 1e-13
    nnz = int(lines[1])
    assert nnz == ref["col"].shape[0]
    got = np.array([[float(x) for x in l.split()] for l in lines[2:2 + nnz]])
    rows = np.repeat(np.arange(bg.n_dofs), np.diff(ref["rowptr"].astype(np.int64)))
    assert np.array_equal(got[:, 0].astype(np.int64), rows) and np.array_equal(got[:, 1].astype(np.uint32), ref["col"])
    assert np.linalg.norm(got[:, 2] - ref["val"]) <= 1e-12 * np.linalg.norm(ref["val"])


@pytest.mark.gpu
def test_cpp_driver_higher_degree_elements(tmp_path, c_oracle):
    """The reference's three calls with FE_Q(2) x FE_DGQ(1) handlers through the drop-in header (general path, elements
    described by their unit support points) against oracle/fe_oracle.py -- from the vector collect() returned (device-
    resident quadrature) and from a copy of it (tuples through the host)."""
    import torch
    from oracle import fe_oracle as fo
    from fe_helpers import uniform_background, continuous_dofs, discontinuous_dofs, random_constraint_lines
    from non_matching_test_suite_b200.grids import BackgroundGrid, circle_grid
    build_driver()
    lo, hi = uniform_background(2, 10)
    im = circle_grid(48)
    e = torch.empty(0, dtype=torch.int32)
    # the first part of the file: Q1 x P0 numbering of the same grids (any valid numbering will do)
    sb1 = fo.unit_support_points(2, 1)
    d1, n1 = continuous_dofs(lo, hi, sb1)
    bg = BackgroundGrid(2, torch.from_numpy(lo), torch.from_numpy(hi), torch.from_numpy(d1.astype(np.int32)), n1, torch.zeros(lo.shape[0], dtype=torch.int32),
                        e, torch.zeros(1, dtype=torch.int64), e, torch.empty(0, dtype=torch.float64))
    grid, out = str(tmp_path / "grid.txt"), str(tmp_path / "out.txt")
    write_grid(grid, bg, im)
    p, q = 2, 1
    sb, si = fo.unit_support_points(2, p, "shuffled", 1), fo.unit_support_points(1, q)
    bd, ns = continuous_dofs(lo, hi, sb)
    idf, ni = discontinuous_dofs(im.n_cells, si.shape[0])
    cd, cp, cm, cw = random_constraint_lines(ns, ns // 10, seed=3)
    with open(grid, "a") as f:
        f.write("%d %d\n" % (sb.shape[0], p) + " ".join(repr(float(x)) for x in sb.ravel()) + "\n%d\n" % ns)
        f.write(" ".join(str(int(x)) for x in bd.ravel()) + "\n%d\n" % len(cd))
        for i in range(len(cd)):
            f.write("%d %d " % (cd[i], cp[i + 1] - cp[i]) + " ".join("%d %r" % (cm[k], float(cw[k])) for k in range(int(cp[i]), int(cp[i + 1]))) + "\n")
        f.write("%d %d\n" % (si.shape[0], q) + " ".join(repr(float(x)) for x in si.ravel()) + "\n%d\n" % ni)
        f.write(" ".join(str(int(x)) for x in idf.ravel()) + "\n")
    subprocess.check_call([BIN, grid, out])
    r = c_oracle.assemble_with_quadrature(bg, im, 2 * p + 1, 1e-15)
    pim, pbg = (r["accepted"] >> np.uint64(32)).astype(np.int64), (r["accepted"] & np.uint64(0xffffffff)).astype(np.int64)
    want = fo.coupling_matrix(lo, hi, bd, ns, cd, cp, cm, cw, im.verts.numpy(), idf, ni, pim, pbg, r["q_offset"], r["points"], r["weights"], sb, si)
    want.sort_indices()
    lines = open(out + ".fe").read().split("\n")
    nnz = int(lines[0])
    got = np.array([[float(x) for x in l.split()] for l in lines[1:1 + nnz]])
    rows = np.repeat(np.arange(ns), np.diff(want.indptr))
    assert nnz == want.nnz and np.array_equal(got[:, 0].astype(np.int64), rows) and np.array_equal(got[:, 1].astype(np.int64), want.indices)
    for k in (2, 3):
        assert np.linalg.norm(got[:, k] - want.data) <= 1e-12 * np.linalg.norm(want.data)
