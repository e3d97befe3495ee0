"""Throughput of the codimension-0 instantiation <2,2,2> (a warped n x n body of convex quads in the disk background):
device-resident inputs -> candidates, clip + integrate, merge, transpose; phase times from the library's CUDA events.
Usage: python profiles/codim0_timing.py [cycle] [n]   -> one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from non_matching_test_suite_b200 import _lib, coupling
from non_matching_test_suite_b200.grids import CONFIGS, background_grid, build_level_map, circle_grid, solid_grid_2d

cycle = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1600
dev = "cuda"
im = solid_grid_2d(n, device=dev)
blo0, bhi0 = circle_grid(32, device=dev).boxes()
lmap = build_level_map(2, 4, CONFIGS["disk"]["band_rounds"], blo0, bhi0)
bg = background_grid(lmap, cycle, near=im.boxes())                     # the leaves that meet the body
ctx = _lib.Context(0)
res = []
for it in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    coupling.set_grids(ctx, bg, im, boxes=True, on_device=True)
    c = ctx.intersect(3, 1e-15)
    ctx.build_sparsity()
    ctx.assemble()
    torch.cuda.synchronize()
    res.append((time.perf_counter() - t0) * 1e3)
ph = ctx.timings() if hasattr(ctx, "timings") else {}
area = float(im.measure().sum())
_, _, sw = ctx.pairs(device="cuda")
print(json.dumps({"instantiation": "<2,2,2>", "background_cycle": cycle, "background_cells": bg.n_cells, "immersed_quads": im.n_cells,
                  "candidates": c["n_candidates"], "pairs": c["n_accepted"], "qpoints": c["n_qpoints"], "nnz": ctx.nnz,
                  "ms_per_assembly_wall": round(min(res), 3), "pairs_per_s": round(c["n_accepted"] / (min(res) * 1e-3), 1),
                  "phases_ms": ph, "area_error": abs(float(sw.sum()) - area), "n_dropped": c["n_dropped"], "device_bytes": ctx.device_bytes()}))
