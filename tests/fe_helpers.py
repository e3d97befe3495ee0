"""Grids and DoF tables for the higher-degree tests: a uniform (single-level, conforming) background on [-1, 1]^d whose
FE_Q(p) dofs are numbered by distinct physical support point, FE_DGQ(q) dofs cell by cell on the immersed grid, and
made-up constraint lines (the algebra of distribute_local_to_global does not care where a line comes from)."""
import numpy as np
import torch

from oracle import fe_oracle as fo


def uniform_background(d, n):
    h = 2.0 / n
    idx = np.stack(np.meshgrid(*([np.arange(n)] * d), indexing="ij"), axis=-1).reshape(-1, d)[:, ::-1]   # x fastest
    lo = -1.0 + idx * h
    return np.ascontiguousarray(lo), np.ascontiguousarray(lo + h)


def continuous_dofs(lo, hi, support_points):
    """[n_cells, dofs_per_cell] global numbering: one dof per distinct physical support point."""
    phys = lo[:, None, :] + support_points[None, :, :] * (hi - lo)[:, None, :]
    key = np.round(phys.reshape(-1, lo.shape[1]) * 2 ** 30).astype(np.int64)
    _, inv = np.unique(key, axis=0, return_inverse=True)
    dofs = inv.reshape(lo.shape[0], support_points.shape[0])
    return dofs.astype(np.uint32), int(inv.max()) + 1


def discontinuous_dofs(n_cells, dofs_per_cell, seed=None):
    d = np.arange(n_cells * dofs_per_cell, dtype=np.uint32).reshape(n_cells, dofs_per_cell)
    if seed is not None:   # any numbering is legal for a DG space: permute the global ids
        d = np.random.default_rng(seed).permutation(n_cells * dofs_per_cell).astype(np.uint32)[d]
    return d, n_cells * dofs_per_cell


def random_constraint_lines(n_dofs, n_lines, seed):
    """Closed lines: constrained dofs are never masters; weights like hanging nodes (1/2, 1/4) plus one Dirichlet-like
    line without masters."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(n_dofs)
    con, free = np.sort(perm[:n_lines]), perm[n_lines:]
    c_ptr, c_master, c_weight = [0], [], []
    for i, _ in enumerate(con):
        k = 0 if i == 0 else (2 if i % 2 else 4)
        m = rng.choice(free, size=k, replace=False)
        c_master += list(np.sort(m)); c_weight += [1.0 / k] * k if k else []
        c_ptr.append(len(c_master))
    return con.astype(np.uint32), np.array(c_ptr, dtype=np.uint64), np.array(c_master, dtype=np.uint32), np.array(c_weight, dtype=np.float64)


def t(a, dtype=None):
    """numpy -> contiguous torch tensor (uint32 / uint64 travel as int32 / int64 bit patterns)."""
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    if a.dtype == np.uint64:
        a = a.view(np.int64)
    return torch.from_numpy(a.copy())
