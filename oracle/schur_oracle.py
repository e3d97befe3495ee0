"""TEST INFRASTRUCTURE ONLY -- numpy/scipy restatement of PoissonLM::solve()
(code/examples/lagrange_multiplier/2d3d/lagrange_multiplier.cc:613-659, 1d2d/lagrange_multiplier.cc:675-730):
S = C K^-1 Ct, `preconditioner` = C K Ct + M applied by multiplication, left-preconditioned restarted GMRES with deal.II's
ReductionControl (|r| <= tol or |r| <= reduce |r_0|), solution = K^-1 (f - Ct lambda), constraints.distribute.
K^-1 is a sparse direct solve here (the reference: CG + AMG to its own tolerance).  PARITY UNPINNED ([UPSTREAM] GMRES
details: left preconditioning, modified Gram-Schmidt, 28 Krylov vectors per restart)."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def solve(K, M, A, f, g, outer=(None, 1e-11, 1e-2), basis=28):
    """A = stored coupling matrix (n_space x n_immersed) so that Ct = A, C = A^T.  Returns lambda, solution (before
    distribute), outer iterations, last preconditioned residual."""
    lu = spla.splu(sp.csc_matrix(K))
    Ct = lambda x: A @ x
    C = lambda y: A.T @ y
    S = lambda x: C(lu.solve(Ct(x)))
    P = lambda x: C(K @ Ct(x)) + M @ x
    b = C(lu.solve(f)) - g
    n = b.shape[0]
    max_it, tol, reduce = outer
    max_it = max_it or K.shape[0]          # the 3D driver: ReductionControl(solution.size(), ...)
    x = np.zeros(n)
    it, rho0, rho, first, done = 0, 0.0, 0.0, True, False
    while not done and it < max_it:
        r = P(b - (S(x) if not first else 0.0))
        rho = np.linalg.norm(r)
        if first:
            rho0, first = rho, False
            if rho0 <= tol:
                break
        if rho <= tol or rho <= reduce * rho0:
            break
        V = [r / rho]
        H = np.zeros((basis + 1, basis))
        cs, sn, gam = np.zeros(basis), np.zeros(basis), np.zeros(basis + 1)
        gam[0] = rho
        k = 0
        while k < basis and it < max_it:
            w = P(S(V[k]))
            for i in range(k + 1):
                H[i, k] = w @ V[i]
                w = w - H[i, k] * V[i]
            H[k + 1, k] = np.linalg.norm(w)
            V.append(w / H[k + 1, k] if H[k + 1, k] > 0 else w)
            for i in range(k):
                a, c = H[i, k], H[i + 1, k]
                H[i, k], H[i + 1, k] = cs[i] * a + sn[i] * c, -sn[i] * a + cs[i] * c
            a, c = H[k, k], H[k + 1, k]
            d = np.hypot(a, c)
            cs[k], sn[k] = (a / d, c / d) if d > 0 else (1.0, 0.0)
            H[k, k], H[k + 1, k] = d, 0.0
            gam[k + 1], gam[k] = -sn[k] * gam[k], cs[k] * gam[k]
            it += 1
            rho = abs(gam[k + 1])
            k += 1
            if rho <= tol or rho <= reduce * rho0:
                done = True
                break
        y = np.linalg.solve(np.triu(H[:k, :k]), gam[:k]) if k else np.zeros(0)
        for i in range(k):
            x = x + y[i] * V[i]
    u = lu.solve(f - Ct(x))
    return x, u, it, rho


def q1_laplace_uniform(d, n, lo=-1.0, hi=1.0):
    """Q1 stiffness matrix of the uniform n^d grid on [lo, hi]^d, dofs numbered lexicographically (x fastest), no
    boundary treatment."""
    h = (hi - lo) / n
    K1 = sp.diags([-np.ones(n), np.r_[1.0, 2.0 * np.ones(n - 1), 1.0], -np.ones(n)], [-1, 0, 1]) / h
    M1 = sp.diags([np.ones(n) / 6, np.r_[1.0 / 3, 2.0 / 3 * np.ones(n - 1), 1.0 / 3], np.ones(n) / 6], [-1, 0, 1]) * h
    if d == 2:
        return (sp.kron(M1, K1) + sp.kron(K1, M1)).tocsr()       # kron(A_y, B_x): x fastest
    return (sp.kron(sp.kron(M1, M1), K1) + sp.kron(sp.kron(M1, K1), M1) + sp.kron(sp.kron(K1, M1), M1)).tocsr()


def constrain(K, f, c_dof, diag=None):
    """deal.II style elimination of homogeneous constraint rows without masters (Dirichlet): row and column zeroed,
    a positive value on the diagonal, right-hand side zero."""
    K = K.tolil(copy=True)
    f = f.copy()
    dval = diag if diag is not None else float(np.mean(K.diagonal()))
    keep = np.ones(K.shape[0], dtype=bool); keep[c_dof] = False
    D = sp.diags(keep.astype(float))
    K2 = (D @ K.tocsr() @ D).tolil()
    for c in c_dof:
        K2[int(c), int(c)] = dval
    f[c_dof] = 0.0
    return K2.tocsr(), f
