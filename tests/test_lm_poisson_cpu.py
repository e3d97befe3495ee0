"""The downstream quantities of the Lagrange-multiplier drivers (L2 / H1 error of the solution, size of the multiplier:
what the reference writes into its tables, 2d3d/lagrange_multiplier.cc:783-802) from the CPU restatements alone:
coupling matrix from the C oracle, Poisson system and error norms from oracle/lm_poisson.py, Schur solve from
oracle/schur_oracle.py.  The smooth configuration must converge at the optimal rates of FE_Q(1) (2 in L2, 1 in H1) with a
multiplier that tends to its exact value, zero -- a wrong coupling matrix, constraint treatment or solve shows up here."""
import numpy as np
import scipy.sparse as sp

from non_matching_test_suite_b200.grids import make_config
from oracle import lm_poisson as lp
from oracle import schur_oracle as so


def solve_cpu(c_oracle, name, cycle):
    d = 2 if name == "disk" else 3
    u_ex, g_ex, rhs = lp.smooth_solution(d)
    bg, im = make_config(name, cycle, band_only=False)
    K, f, T = lp.poisson_system(bg, rhs)
    M, g, J = lp.interface_data(im, u_ex)
    r = c_oracle.assemble(bg, im)
    A = sp.csr_matrix((r["val"], r["col"].astype(np.int64), r["rowptr"].astype(np.int64)), shape=(bg.n_dofs, im.n_dofs))
    lam, ut, it, _ = so.solve(K, M, A, f, g, outer=(bg.n_dofs, 1e-11, 1e-2))      # the drivers' ReductionControl
    l2, h1 = lp.errors(bg, T @ ut, u_ex, g_ex)
    return dict(bg=bg, im=im, K=K, M=M, f=f, g=g, J=J, T=T, lam=lam, u=T @ ut, it=it, l2=l2, h1=h1, lam_norm=float(np.sqrt(np.sum(lam ** 2 * J))))


def test_smooth_disk_converges_at_the_optimal_rates(c_oracle):
    res = [solve_cpu(c_oracle, "disk", c) for c in range(4)]
    for a, b in zip(res, res[1:]):
        assert 1.9 < np.log2(a["l2"] / b["l2"]) < 2.1
        assert 0.95 < np.log2(a["h1"] / b["h1"]) < 1.05
        assert b["lam_norm"] < 0.5 * a["lam_norm"]
    assert all(r["it"] <= 12 for r in res)


def test_smooth_sphere_first_cycle(c_oracle):
    r = solve_cpu(c_oracle, "sphere", 0)
    assert 5e-2 < r["l2"] < 8e-2 and 2.0 < r["h1"] < 3.0 and r["it"] <= 8


def test_hanging_node_reduction_reproduces_linear_functions():
    """T u~ interpolates: a linear function given on the free dofs is reproduced on the hanging nodes."""
    bg, _ = make_config("disk", 1, band_only=False, dirichlet=False)
    T, free = lp.reduction_matrix(bg)
    x = bg.dof_coords.numpy()
    lin = 0.3 + 1.7 * x[:, 0] - 0.9 * x[:, 1]
    assert np.abs(T @ np.where(free, lin, 0.0) - lin).max() < 1e-14
