#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the exact-rational oracle (oracle/exact_oracle.py).

The reference (deal.II + CGAL) cannot run in this environment and ships no golden vectors, so
these fixtures are NOT reference outputs: they are the exact integrals of the reference's
formulation (rules R1-R6) on the synthetic grids of non_matching_test_suite_b200.grids,
rounded to double.  Parity therefore stays "unpinned" with respect to a real deal.II/CGAL run.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from non_matching_test_suite_b200.grids import make_config  # noqa: E402
from oracle import exact_oracle as eo  # noqa: E402

CASES = [("disk", 0), ("disk", 1), ("flower", 0), ("sphere", 0)]


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    for name, cycle in CASES:
        bg, im = make_config(name, cycle, band_only=False)
        r = eo.assemble(bg, im, tol=1e-15)
        cand = np.array([(m << 32) | b for m, b in r["candidates"]], dtype=np.uint64)
        acc = np.array([(m << 32) | b for m, b in r["accepted"]], dtype=np.uint64)
        pat = sorted(r["pattern"])
        rows = np.array([p[0] for p in pat], dtype=np.int64)
        cols = np.array([p[1] for p in pat], dtype=np.uint32)
        vals = np.array([r["values"].get(p, 0.0) for p in pat])
        rowptr = np.zeros(bg.n_dofs + 1, dtype=np.uint64)
        np.add.at(rowptr, rows + 1, 1)
        rowptr = np.cumsum(rowptr).astype(np.uint64)
        codes = np.array(r["codes"], dtype=np.uint8) if r["codes"] is not None else np.zeros(0, dtype=np.uint8)
        np.savez_compressed(os.path.join(out, "%s_c%d.npz" % (name, cycle)), candidates=cand, accepted=acc,
                            sum_w=np.array(r["sum_w"]), rowptr=rowptr, col=cols, val=vals, codes=codes,
                            n_bg=bg.n_cells, n_im=im.n_cells, n_space_dofs=bg.n_dofs)
        print(name, cycle, "cand", cand.size, "acc", acc.size, "nnz", cols.size)


if __name__ == "__main__":
    main()
