"""TEST INFRASTRUCTURE ONLY -- the Poisson problem of the Lagrange-multiplier drivers on the synthetic grids, restated in
numpy/scipy so that the downstream quantities the reference tabulates (L2 / H1 error of the solution, norm of the
multiplier) can be compared between the GPU solve (nm_schur_solve) and the CPU restatement (oracle/schur_oracle.py).

Follows PoissonLM::assemble_system and ::output_results of
  code/examples/lagrange_multiplier/1d2d/lagrange_multiplier.cc:565-672, 733-800   (2d3d: :520-610, 662-730)
for the smooth configurations (data/lm_1d2d_disk.smooth.prm, lm_2d3d.smooth_sphere.prm):
  -Laplace u = f in [-1,1]^d, u = u_exact on the boundary, u = u_exact on the immersed interface (multiplier),
  stiffness and load with QGauss(2 p + 1) on FE_Q(1), distributed through the space constraints (hanging nodes +
  Dirichlet lines), embedded right-hand side and mass matrix with QGauss(2 q + 1) = the midpoint rule on FE_DGQ(0).
The exact solutions of the smooth configurations vanish on the boundary of [-1,1]^d, so every constraint line is
homogeneous.  PARITY UNPINNED: deal.II is not available here; the tables of the paper are not in the repository."""
import numpy as np
import scipy.sparse as sp


def gauss01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def _q1(d, xi):
    """values [n_q, 2^d] and reference gradients [n_q, 2^d, d] of the Q1 shape functions (deal.II vertex order) at xi [n_q, d]"""
    nq = xi.shape[0]
    val = np.ones((nq, 1 << d))
    grad = np.ones((nq, 1 << d, d))
    for v in range(1 << d):
        for k in range(d):
            b = (v >> k) & 1
            f = xi[:, k] if b else 1.0 - xi[:, k]
            df = np.ones(nq) if b else -np.ones(nq)
            val[:, v] *= f
            for j in range(d):
                grad[:, v, j] *= df if j == k else f
    return val, grad


def tensor_rule(d, n):
    x, w = gauss01(n)
    g = np.meshgrid(*([x] * d), indexing="ij")
    gw = np.meshgrid(*([w] * d), indexing="ij")
    return np.stack([a.ravel() for a in g], 1), np.prod(np.stack([a.ravel() for a in gw], 1), 1)


def reduction_matrix(bg):
    """T with u = T u~: identity on free dofs, the weights of its line on a hanging node, a zero row on a Dirichlet line."""
    n = bg.n_dofs
    c_dof, c_ptr = bg.c_dof.numpy().astype(np.int64), bg.c_ptr.numpy().astype(np.int64)
    free = np.ones(n, dtype=bool)
    free[c_dof] = False
    rows = [np.nonzero(free)[0]]
    cols = [np.nonzero(free)[0]]
    vals = [np.ones(int(free.sum()))]
    rows.append(np.repeat(c_dof, np.diff(c_ptr)))
    cols.append(bg.c_master.numpy().astype(np.int64))
    vals.append(bg.c_weight.numpy())
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)), free


def poisson_system(bg, rhs):
    """(K~, f~) as AffineConstraints::distribute_local_to_global leaves them: T^t K T with a positive diagonal on the
    constrained rows, T^t f with zeros there."""
    d = bg.spacedim
    lo, hi = bg.lo.numpy(), bg.hi.numpy()
    h = hi - lo
    dofs = bg.dofs.numpy().astype(np.int64)
    n, nv = lo.shape[0], 1 << d
    xi, w = tensor_rule(d, 3)                                    # QGauss(2 * 1 + 1)
    val, grad = _q1(d, xi)
    vol = np.prod(h, axis=1)
    # element stiffness: sum_q w_q vol sum_k grad_a,k grad_b,k / h_k^2
    Ke = np.zeros((n, nv, nv))
    for k in range(d):
        G = np.einsum("q,qa,qb->ab", w, grad[:, :, k], grad[:, :, k])
        Ke += (vol / h[:, k] ** 2)[:, None, None] * G[None]
    xq = lo[:, None, :] + xi[None] * h[:, None, :]               # [n, n_q, d]
    fe = np.einsum("q,nq,qa->na", w, rhs(xq.reshape(-1, d)).reshape(n, -1), val) * vol[:, None]
    K = sp.csr_matrix((Ke.ravel(), (np.repeat(dofs, nv, axis=1).ravel(), np.tile(dofs, (1, nv)).ravel())), shape=(bg.n_dofs, bg.n_dofs))
    f = np.bincount(dofs.ravel(), weights=fe.ravel(), minlength=bg.n_dofs)
    T, free = reduction_matrix(bg)
    Kt = (T.T @ K @ T).tocsr()
    ft = T.T @ f
    dval = float(np.mean(np.abs(Kt.diagonal()[free])))
    Kt = Kt + sp.diags(np.where(free, 0.0, dval))
    ft[~free] = 0.0
    return Kt.tocsr(), ft, T


def interface_data(im, u_exact):
    """FE_DGQ(0) on the immersed grid with QGauss(1): mass matrix diag(J(centre)), right-hand side u(centre) J(centre)."""
    v = im.verts.numpy()
    if im.dim == 1:
        c = 0.5 * (v[:, 0] + v[:, 1])
        J = np.linalg.norm(v[:, 1] - v[:, 0], axis=1)
    else:                                                         # bilinear quad, deal.II vertex order
        c = 0.25 * v.sum(axis=1)
        du = 0.5 * ((v[:, 1] - v[:, 0]) + (v[:, 3] - v[:, 2]))
        dv = 0.5 * ((v[:, 2] - v[:, 0]) + (v[:, 3] - v[:, 1]))
        J = np.linalg.norm(np.cross(du, dv), axis=1)
    return sp.diags(J).tocsr(), u_exact(c) * J, J


def errors(bg, u, u_exact, grad_exact):
    """L2 norm and H1 seminorm of u_h - u_exact with QGauss(3) on every cell (VectorTools::integrate_difference)."""
    d = bg.spacedim
    lo, hi = bg.lo.numpy(), bg.hi.numpy()
    h = hi - lo
    dofs = bg.dofs.numpy().astype(np.int64)
    n = lo.shape[0]
    xi, w = tensor_rule(d, 3)
    val, grad = _q1(d, xi)
    vol = np.prod(h, axis=1)
    ue = u[dofs]                                                  # [n, 2^d]
    uh = ue @ val.T                                               # [n, n_q]
    gh = np.einsum("na,qak->nqk", ue, grad) / h[:, None, :]
    xq = (lo[:, None, :] + xi[None] * h[:, None, :]).reshape(-1, d)
    e0 = uh - u_exact(xq).reshape(n, -1)
    e1 = gh - grad_exact(xq).reshape(n, -1, d)
    l2 = np.sqrt(np.sum(vol[:, None] * w[None] * e0 ** 2))
    h1 = np.sqrt(np.sum(vol[:, None] * w[None] * np.sum(e1 ** 2, axis=2)))
    return float(l2), float(h1)


def smooth_solution(d):
    """u = prod_k sin(2 pi x_k): data/lm_1d2d_disk.smooth.prm:45,131 (3D: lm_2d3d.smooth_sphere.prm), f = 4 d pi^2 u"""
    tp = 2.0 * np.pi

    def u(x):
        return np.prod(np.sin(tp * x), axis=1)

    def grad(x):
        g = np.empty_like(x)
        for k in range(d):
            t = np.cos(tp * x[:, k]) * tp
            for j in range(d):
                if j != k:
                    t = t * np.sin(tp * x[:, j])
            g[:, k] = t
        return g

    def rhs(x):
        return d * tp * tp * u(x)

    return u, grad, rhs
