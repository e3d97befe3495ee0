"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the coupling matrix for general tensor-product Lagrange elements
(FE_Q(p) background, FE_DGQ(q) immersed): reference code/include/coupling_utilities.h:238-289 (values) and :100-139
(pattern, keep_constrained_entries = true), on a GIVEN list of (space cell, immersed cell, points, weights) tuples.

PARITY UNPINNED: deal.II is absent here.  [UPSTREAM] facts restated: FE_Q(p) / FE_DGQ(q) are the Lagrange bases of
their unit support points (Gauss-Lobatto nodes per direction); Mapping::transform_real_to_unit_cell on a codimension-one
cell returns the closest point of the (bi)linear cell.  Pinned by: reduction to the Q1 x P0 oracle for (p, q) = (1, 0),
partition of unity, independence of the dof order, and exactness of the 2D rules (tests/test_fe_cpu.py).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def gauss_lobatto_nodes(degree: int) -> np.ndarray:
    """The p + 1 Gauss-Lobatto nodes on [0, 1] (support points of FE_Q(p) / FE_DGQ(p) per direction; p = 0: the midpoint)."""
    if degree == 0:
        return np.array([0.5])
    if degree == 1:
        return np.array([0.0, 1.0])
    # interior nodes = roots of P'_p
    c = np.zeros(degree + 1); c[degree] = 1.0
    r = np.sort(np.real(np.polynomial.legendre.legroots(np.polynomial.legendre.legder(c))))
    return np.concatenate([[0.0], 0.5 + 0.5 * r, [1.0]])


def unit_support_points(dim: int, degree: int, order: str = "lexicographic", seed: int = 0) -> np.ndarray:
    """[ (p+1)^dim, dim ] support points; `order` = "lexicographic" (x fastest) or "shuffled" (any fixed permutation:
    the library must not depend on the order, it is told the points)."""
    x = gauss_lobatto_nodes(degree)
    if dim == 0:
        return np.zeros((1, 0))
    grids = np.meshgrid(*([x] * dim), indexing="ij")             # grids[a] varies along array axis a
    pts = np.stack([grids[dim - 1 - k].reshape(-1) for k in range(dim)], axis=1)   # coordinate 0 runs fastest
    if order == "shuffled":
        pts = pts[np.random.default_rng(seed).permutation(pts.shape[0])]
    return np.ascontiguousarray(pts)


def lagrange_values(nodes: np.ndarray, x: np.ndarray) -> np.ndarray:
    """[len(x), len(nodes)] values of the Lagrange polynomials of `nodes` at x."""
    out = np.ones((x.shape[0], nodes.shape[0]))
    for a in range(nodes.shape[0]):
        for b in range(nodes.shape[0]):
            if a != b:
                out[:, a] *= (x - nodes[b]) / (nodes[a] - nodes[b])
    return out


def shape_values(support_points: np.ndarray, xi: np.ndarray) -> np.ndarray:
    """[n_points, dofs_per_cell]: tensor-product Lagrange basis of the support points at unit coordinates xi [n, dim]."""
    dpc, dim = support_points.shape
    if dim == 0 or dpc == 1:
        return np.ones((xi.shape[0], 1))
    nodes = np.unique(np.round(support_points[:, 0], 12))
    vals = np.ones((xi.shape[0], dpc))
    for k in range(dim):
        L = lagrange_values(nodes, xi[:, k])
        idx = np.array([int(np.argmin(np.abs(nodes - s))) for s in support_points[:, k]])
        vals *= L[:, idx]
    return vals


def immersed_unit_coords(verts: np.ndarray, x: np.ndarray) -> np.ndarray:
    """Closest point of the immersed cell (segment [2, 2] or bilinear quad [4, 3], deal.II vertex order) to every x, in
    unit coordinates (coupling_utilities.h:251-254)."""
    if verts.shape[0] == 2:
        e = verts[1] - verts[0]
        return (((x - verts[0]) @ e) / (e @ e))[:, None]
    v0, v1, v2, v3 = verts
    a, b, c = v1 - v0, v2 - v0, v0 - v1 - v2 + v3
    eta = np.full((x.shape[0], 2), 0.5)
    for _ in range(50):
        s, t = eta[:, 0:1], eta[:, 1:2]
        r = v0 + a * s + b * t + c * s * t - x
        Js, Jt = a + c * t, b + c * s
        gs, gt = (Js * r).sum(1), (Jt * r).sum(1)
        hss, htt, hst = (Js * Js).sum(1), (Jt * Jt).sum(1), (Js * Jt).sum(1) + (r * c).sum(1)
        det = hss * htt - hst * hst
        ds, dt = (htt * gs - hst * gt) / det, (hss * gt - hst * gs) / det
        eta = eta - np.stack([ds, dt], 1)
        if np.abs(ds).max() + np.abs(dt).max() < 1e-16:
            break
    return eta


def coupling_matrix(bg_lo, bg_hi, bg_dofs, n_space, c_dof, c_ptr, c_master, c_weight, im_verts, im_dofs, n_immersed, pair_im, pair_bg,
                    q_off, points, weights, sp_bg, sp_im):
    """CSR (with explicit zeros where the reference keeps constrained entries) of the stored n_space x n_immersed matrix."""
    bg_lo, bg_hi, im_verts = (np.asarray(a, dtype=np.float64) for a in (bg_lo, bg_hi, im_verts))
    line = {int(d): (np.asarray(c_master[c_ptr[i]:c_ptr[i + 1]], dtype=np.int64), np.asarray(c_weight[c_ptr[i]:c_ptr[i + 1]], dtype=np.float64))
            for i, d in enumerate(np.asarray(c_dof, dtype=np.int64))}
    rows, cols, vals = [], [], []
    for p in range(len(pair_im)):
        m, b = int(pair_im[p]), int(pair_bg[p])
        x = np.asarray(points[int(q_off[p]):int(q_off[p + 1])], dtype=np.float64)
        w = np.asarray(weights[int(q_off[p]):int(q_off[p + 1])], dtype=np.float64)
        xi = (x - bg_lo[b]) / (bg_hi[b] - bg_lo[b])
        phi = shape_values(sp_bg, xi)
        psi = shape_values(sp_im, immersed_unit_coords(im_verts[m], x))
        local = np.einsum("qi,qj,q->ij", phi, psi, w)
        for i, r in enumerate(np.asarray(bg_dofs[b], dtype=np.int64)):
            for j, c in enumerate(np.asarray(im_dofs[m], dtype=np.int64)):
                if int(r) in line:
                    rows.append(int(r)); cols.append(int(c)); vals.append(0.0)                  # kept, structurally zero
                    mm, ww = line[int(r)]
                    for k in range(len(mm)):
                        rows.append(int(mm[k])); cols.append(int(c)); vals.append(ww[k] * local[i, j])
                else:
                    rows.append(int(r)); cols.append(int(c)); vals.append(local[i, j])
    A = sp.coo_matrix((np.array(vals), (np.array(rows, dtype=np.int64), np.array(cols, dtype=np.int64))), shape=(n_space, n_immersed)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A
