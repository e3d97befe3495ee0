"""The geometry core of clip_moments3_kernel (csrc/nm_clip_core.cuh) compiled for the HOST and compared, candidate by
candidate, with the C oracle: accepted set identical, measures and the eight local entries to rounding.  Both polygon
constructions are covered -- the section algorithm the kernel uses (plane of the triangle cut by the box, clipped by
the triangle's edges) and the box-plane clipping it replaced."""
import os
import subprocess

import numpy as np
import pytest
import torch

from non_matching_test_suite_b200.grids import make_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")


VARIANTS = {"section": ["-DNM_CLIP_SECTION=1"], "clip_v1": ["-DNM_CLIP_SECTION=0", "-DNM_CLIP_BITS=1"],
            "clip_v2": ["-DNM_CLIP_SECTION=0", "-DNM_CLIP_BITS=2"],
            # what clip_moments3_kernel is built with: plane-major outcodes, vertex pool of nine slots (bound-checked here)
            "clip_v2_pool9": ["-DNM_CLIP_SECTION=0", "-DNM_CLIP_BITS=2", "-DNM_CLIP_SLOTS=1", "-DHOST_POOL9=1"],
            "clip_v2_planeslots": ["-DNM_CLIP_SECTION=0", "-DNM_CLIP_BITS=2", "-DNM_CLIP_SLOTS=2"]}


def _exe(variant):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "clip_core_host_" + variant)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off"] + VARIANTS[variant] +
                          ["-o", exe, os.path.join(ROOT, "tests", "cpp", "clip_core_host.cpp")])
    return exe


def _run(exe, tmp_path, lo, hi, quads, codes, tol=1e-15):
    n = lo.shape[0]
    rec = np.concatenate([lo, hi, quads.reshape(n, 12), codes.reshape(n, 1).astype(np.float64)], axis=1)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        np.array([float(n)]).tofile(f)
        rec.astype(np.float64).tofile(f)
    subprocess.check_call([exe, fin, fout, repr(tol)])
    out = np.fromfile(fout, dtype=np.float64).reshape(n, 11)
    return out[:, 0], out[:, 1:9], out[:, 9].astype(int), out[:, 10].astype(int)


def _problem(bg, im, c_oracle):
    ilo, ihi = im.boxes()
    cand = c_oracle.candidates(bg.lo, bg.hi, ilo, ihi)
    codes = c_oracle.quad_split(im.verts)
    sumw, local, nq, dropped = c_oracle.pairs(3, cand, bg.lo, bg.hi, im.verts, codes, 3, 1e-15)
    b, m = (cand & np.uint64(0xffffffff)).astype(np.int64), (cand >> np.uint64(32)).astype(np.int64)
    return (bg.lo.numpy()[b], bg.hi.numpy()[b], im.verts.numpy()[m], codes[m]), sumw, local, dropped


@pytest.mark.parametrize("section", list(VARIANTS))
@pytest.mark.parametrize("cycle", [0, 2, 3])
def test_core_matches_oracle_on_sphere(tmp_path, c_oracle, section, cycle):
    bg, im = make_config("sphere", cycle, band_only=cycle > 2)
    args, sumw, local, dropped = _problem(bg, im, c_oracle)
    got_w, got_l, nfan, drop = _run(_exe(section), tmp_path, *args)
    acc, acc_ref = got_w > 1e-15, sumw > 1e-15
    assert np.array_equal(acc, acc_ref)
    scale = np.abs(sumw).max()
    # sub-triangles under the reference's absolute |J| < 1e-12 (area < 5e-13) are dropped per piece of ITS subdivision
    # (tetrahedra x triangles, oracle) resp. of the fan here: the only admissible difference, bounded by their count
    slack = 5e-13 * (dropped + int(drop.sum()))
    assert np.abs(got_w - sumw).max() <= 1e-12 * scale + slack
    assert np.abs(got_l[acc] - local[acc]).max() <= 1e-12 * scale + slack
    assert abs(got_w.sum() - sumw.sum()) <= 1e-13 * sumw.sum() + slack
    assert (nfan[acc] >= 1).all()


@pytest.mark.parametrize("section", list(VARIANTS))
def test_core_on_degenerate_and_random_pairs(tmp_path, c_oracle, section):
    """Faces shared by two cells (counted for both), quads through grid vertices, flat quads on grid planes, plus random
    boxes against random, slightly warped quads of all relative sizes."""
    from degenerate_cases import case_3d
    bg, im = case_3d()
    args, sumw, local, _ = _problem(bg, im, c_oracle)
    got_w, got_l, _, _ = _run(_exe(section), tmp_path, *args)
    assert np.array_equal(got_w > 1e-15, sumw > 1e-15)
    assert np.abs(got_w - sumw).max() <= 1e-13 and np.abs(got_l - local * (sumw > 1e-15)[:, None]).max() <= 1e-13
    from test_gpu_parity import _random_problem
    for seed, (nb, ni) in enumerate([(400, 200), (2000, 60)]):
        bgr, imr = _random_problem(3, nb, ni, seed=11 + seed)
        args, sumw, local, dropped = _problem(bgr, imr, c_oracle)
        got_w, got_l, _, drop = _run(_exe(section), tmp_path, *args)
        tiny = np.abs(got_w - sumw) <= 1e-12 * max(1, dropped)          # slivers under |J| < 1e-12 depend on the subdivision
        assert tiny.all()
        acc = (got_w > 1e-15) & (sumw > 1e-15)
        assert ((got_w > 1e-15) == (sumw > 1e-15))[np.abs(sumw) > 1e-11].all()
        scale = np.abs(sumw).max()
        assert np.abs(got_l[acc] - local[acc]).max() <= 1e-12 * scale + 1e-12 * max(1, dropped)


def test_both_outcode_layouts_give_identical_bits(tmp_path, c_oracle):
    """item_clip_polygon_v2 (plane-major 32-bit outcode words, slots of cut-off vertices reused: nine suffice) is the same algorithm as v1 (position-major 64-bit list):
    same planes, same runs, same vertices -- every output double identical."""
    from degenerate_cases import case_3d
    from test_gpu_parity import _random_problem
    problems = [make_config("sphere", 3, band_only=True), case_3d(), _random_problem(3, 2000, 60, seed=5), _random_problem(3, 300, 300, seed=6)]
    for bg, im in problems:
        args, _, _, _ = _problem(bg, im, c_oracle)
        r1 = _run(_exe("clip_v1"), tmp_path, *args)
        for variant in ("clip_v2", "clip_v2_pool9", "clip_v2_planeslots"):
            r2 = _run(_exe(variant), tmp_path, *args)
            for a, b in zip(r1, r2):
                assert np.array_equal(a, b)
