"""Downstream L2 / H1 errors of the Lagrange-multiplier Poisson problem with the coupling matrix assembled AND the Schur
complement solved on the GPU (nm_intersect / nm_assemble / nm_schur_solve), against the all-CPU restatement
(tests/test_lm_poisson_cpu.py): same outer iteration count, error norms equal to six significant digits -- the
north_star's "downstream L2/H1 to 6 digits", with the caveat of every parity statement here: the other side is the
repository's own restatement, not a deal.II run."""
import numpy as np
import pytest
import torch

from non_matching_test_suite_b200 import coupling
from oracle import lm_poisson as lp
from test_lm_poisson_cpu import solve_cpu

pytestmark = pytest.mark.gpu


def _csr(B):
    return (torch.from_numpy(B.indptr.astype(np.int64)), torch.from_numpy(B.indices.astype(np.int32)), torch.from_numpy(B.data.astype(np.float64)))


@pytest.mark.parametrize("name,cycle", [("disk", 0), ("disk", 1), ("disk", 2), ("disk", 3), ("sphere", 0), ("sphere", 1)])
def test_error_tables_match_the_cpu_pipeline(c_oracle, name, cycle):
    ref = solve_cpu(c_oracle, name, cycle)
    bg, im = ref["bg"], ref["im"]
    d = bg.spacedim
    u_ex, g_ex, _ = lp.smooth_solution(d)
    ctx = coupling.make_context(bg, im, boxes=True)
    info = coupling.collect_quadratures_on_overlapped_grids(bg, im, 3, 1e-15, ctx=ctx, with_quadrature=False)
    C = coupling.create_coupling_mass_matrix_with_exact_intersections(bg, im, info, use_given_quadrature=False)
    lam, u, st = coupling.solve(C, _csr(ref["K"]), _csr(ref["M"]), torch.from_numpy(ref["f"]), torch.from_numpy(ref["g"]),
                                inner=(20000, 1e-16, 1e-13), outer=(bg.n_dofs, 1e-11, 1e-2))     # PoissonLM::solve()
    assert st["outer_iterations"] == ref["it"]
    l2, h1 = lp.errors(bg, u.numpy(), u_ex, g_ex)
    lam_norm = float(np.sqrt(np.sum(lam.numpy() ** 2 * ref["J"])))
    assert abs(l2 - ref["l2"]) <= 1e-6 * ref["l2"]
    assert abs(h1 - ref["h1"]) <= 1e-6 * ref["h1"]
    assert abs(lam_norm - ref["lam_norm"]) <= 1e-6 * ref["lam_norm"]
    # the solution itself, hanging nodes and Dirichlet rows included (constraints.distribute on the device)
    assert np.linalg.norm(u.numpy() - ref["u"]) <= 1e-7 * np.linalg.norm(ref["u"])
    ctx.close()
