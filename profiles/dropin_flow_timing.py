"""The reference's three-call flow through the Python mirror (host quadrature round trip), timed:
collect_quadratures_on_overlapped_grids -> create_coupling_sparsity_pattern... -> create_coupling_mass_matrix...(given quadrature)
vs. the fused device path.  Sphere cycle 6 (4.0e6 pairs, 7.8e7 quadrature points = 2.5 GB of host arrays)."""
import sys, time, torch
sys.path.insert(0, '/root/repo')
from non_matching_test_suite_b200 import coupling
from non_matching_test_suite_b200.grids import make_config

bg, im = make_config("sphere", 6, band_only=True)
for rep in range(2):
    t0 = time.perf_counter()
    info = coupling.collect_quadratures_on_overlapped_grids(bg, im, 3, 1e-15)
    t1 = time.perf_counter()
    rp, col = coupling.create_coupling_sparsity_pattern_with_exact_intersections(info, bg, im)
    t2 = time.perf_counter()
    M = coupling.create_coupling_mass_matrix_with_exact_intersections(bg, im, info, use_given_quadrature=True)
    rp1, col1, val1 = M.csr()
    t3 = time.perf_counter()
    info.ctx.close()
    ctx = coupling.make_context(bg, im, boxes=True)
    t4 = time.perf_counter()
    ctx.intersect(3, 1e-15); ctx.build_sparsity(); ctx.assemble(); rp2, col2, val2 = ctx.csr()
    t5 = time.perf_counter()
    ctx.close()
    print("rep %d pairs %d points %d | collect (incl. D2H of quadrature) %.0f ms | sparsity %.0f ms | mass matrix from given quadrature (H2D) %.0f ms | "
          "three calls %.0f ms | fused path incl. CSR read-back %.0f ms | rel diff %.1e" % (
              rep, len(info), info.weights.shape[0], (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3, (t5 - t4) * 1e3,
              float((val1 - val2).norm() / val2.norm())), flush=True)
