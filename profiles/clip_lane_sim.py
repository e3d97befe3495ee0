#!/usr/bin/env python
"""Lane utilisation of clip_moments3_kernel's two loops, replayed on the CPU from the real item stream.

    g++ -O2 -std=c++17 -o /tmp/clip_item_stats profiles/clip_item_stats.cpp
    python profiles/clip_lane_sim.py [cycle]          # sphere, band-only background (default cycle 5)

clip_item_stats.cpp runs the host build of the kernel's geometry core (csrc/nm_clip_core.cuh) over every candidate and
records, per (candidate, triangle) item: gate passed, gate without the nine cross-product axes, number of cutting planes
(= trip count of the plane loop) and vertices of the clipped polygon (fan loop trips = n - 2).  A warp runs as many trips
as its slowest lane; this script counts warp trips for the ring order the kernel uses and for alternatives.
Sphere cycle 5 (1.66e6 candidates): 45 % of the items pass the gate (62 % without the cross axes); planes 2/3/4/5/6 =
4/44/40/11/1 %; plane loop utilisation 0.72 in ring order, 0.86 with two rings split at three planes, 1.0 when sorted;
fan loop 0.46."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import c_oracle                                    # noqa: E402  (measurement script, not the product)
from non_matching_test_suite_b200.grids import make_config     # noqa: E402


def main():
    cycle = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    exe = "/tmp/clip_item_stats"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "profiles", "clip_item_stats.cpp")])
    bg, im = make_config("sphere", cycle, band_only=True)
    ilo, ihi = im.boxes()
    cand = c_oracle.candidates(bg.lo, bg.hi, ilo, ihi)
    codes = c_oracle.quad_split(im.verts)
    b, m = (cand & np.uint64(0xffffffff)).astype(np.int64), (cand >> np.uint64(32)).astype(np.int64)
    n = len(cand)
    rec = np.concatenate([bg.lo.numpy()[b], bg.hi.numpy()[b], im.verts.numpy()[m].reshape(n, 12), codes[m].reshape(n, 1).astype(np.float64)], axis=1)
    with open("/tmp/clip_items_in.bin", "wb") as f:
        np.array([float(n)]).tofile(f)
        rec.tofile(f)
    subprocess.check_call([exe, "/tmp/clip_items_in.bin", "/tmp/clip_items_out.bin"])
    o = np.fromfile("/tmp/clip_items_out.bin", dtype=np.int8).reshape(-1, 2, 4)
    gate, planes, pn, gate_nc = o[:, :, 0].astype(bool), o[:, :, 1], o[:, :, 2], o[:, :, 3].astype(bool)
    print("candidates", n, "items", 2 * n, "pass gate", int(gate.sum()), "pass gate without cross axes", int(gate_nc.sum()))
    g = gate.reshape(-1)
    pl, nn = planes.reshape(-1)[g], pn.reshape(-1)[g]
    print("cutting planes histogram", np.bincount(pl, minlength=7))
    print("polygon vertices histogram", np.bincount(nn, minlength=10))

    def rounds(x, w=32):
        k = len(x) // w * w
        r = x[:k].reshape(-1, w)
        return r.max(1).sum(), r.sum() / w

    for name, x in [("plane loop", pl), ("fan loop", np.maximum(nn - 2, 0))]:
        tot, ideal = rounds(x)
        print("%s, ring order: warp trips %d, ideal %d, utilisation %.3f" % (name, tot, ideal, ideal / tot))
    for thr in (2, 3, 4):
        tot = ideal = 0
        for c in (False, True):
            t, i = rounds(pl[(pl > thr) == c])
            tot, ideal = tot + t, ideal + i
        print("plane loop, two rings split at %d planes: utilisation %.3f" % (thr, ideal / tot))


if __name__ == "__main__":
    main()
