"""Throughput of nm_assemble_nitsche on a device-resident quadrature stream (sphere, band-only background).
Usage: python profiles/nitsche_timing.py [cycle]   -> one JSON line."""
import json
import sys
import time

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from non_matching_test_suite_b200 import coupling
from non_matching_test_suite_b200.grids import make_config

cycle = int(sys.argv[1]) if len(sys.argv) > 1 else 6
bg, im = make_config("sphere", cycle, band_only=True)
ctx = coupling.make_context(bg, im, boxes=True)
c = ctx.intersect(3, 1e-15)
_, abg, _ = ctx.pairs(device="cuda")
off, pts, w = ctx.quadrature(device="cuda")
coef = 2.0 + pts[:, 0]
f = 1.0 + pts[:, 1] ** 2
res = {}
for label, kw in (("matrix", dict(coefficient=coef)), ("matrix+rhs", dict(coefficient=coef, rhs_values=f)),
                  ("rhs", dict(coefficient=coef, rhs_values=f, matrix=False))):
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        (rp, col, val), rhs = ctx.assemble_nitsche(abg, off, pts, w, penalty=10.0, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    res[label] = {"ms": round(dt * 1e3, 2), "pairs_per_s": round(c["n_accepted"] / dt, 1), "nnz": 0 if col is None else int(col.numel())}
print(json.dumps({"cycle": cycle, "pairs": c["n_accepted"], "qpoints": c["n_qpoints"], "space_dofs": bg.n_dofs, "constraint_lines": int(bg.c_dof.numel()),
                  "device_bytes": ctx.device_bytes(), **res}))
