"""Hand-made degenerate configurations (SURVEY.md 8(c) rules R3 and R5): touching at a point or along a
lower-dimensional set contributes nothing; an immersed simplex lying exactly in a face shared by two
background cells is counted for both of them (that is what the reference does)."""
import torch

from non_matching_test_suite_b200.grids import BackgroundGrid, ImmersedGrid


def _grid(d, n):
    """n^d unit cells on [0,n]^d, Q1 dofs numbered lexicographically, no constraints."""
    rng = torch.arange(n, dtype=torch.float64)
    mesh = torch.stack(torch.meshgrid(*([rng] * d), indexing="ij"), dim=-1).reshape(-1, d).flip(-1)  # x fastest
    lo, hi = mesh.clone(), mesh + 1.0
    nv = n + 1
    dofs = torch.zeros(lo.shape[0], 1 << d, dtype=torch.int32)
    for v in range(1 << d):
        idx = torch.zeros(lo.shape[0], dtype=torch.int64)
        for k in range(d):
            idx += (lo[:, k].long() + ((v >> k) & 1)) * nv ** k
        dofs[:, v] = idx.to(torch.int32)
    e = torch.empty(0, dtype=torch.int32)
    return BackgroundGrid(d, lo, hi, dofs, nv ** d, torch.zeros(lo.shape[0], dtype=torch.int32), e, torch.zeros(1, dtype=torch.int64), e,
                          torch.empty(0, dtype=torch.float64))


def case_2d():
    bg = _grid(2, 3)
    segs = torch.tensor([
        [[1.0, 0.2], [1.0, 0.8]],     # lies exactly in the edge shared by cells (0,0) and (1,0): counted twice
        [[1.0, 1.0], [1.6, 1.7]],     # starts at a grid vertex: cells that only touch the point get nothing
        [[0.3, 0.4], [2.0, 0.4]],     # ends exactly on the face x = 2
        [[0.5, 1.0], [2.5, 1.0]],     # runs along a grid line across three cells (each counted from both sides)
        [[2.25, 2.25], [2.75, 2.5]],  # strictly inside one cell
        [[0.0, 3.0], [0.0, 2.5]],     # on the domain boundary, ending at a corner of the domain
    ], dtype=torch.float64)
    im = ImmersedGrid(1, 2, segs, torch.arange(segs.shape[0], dtype=torch.int32)[:, None], segs.shape[0])
    return bg, im


def case_3d():
    bg = _grid(3, 3)
    quads = torch.tensor([
        [[0.2, 0.2, 1.0], [0.8, 0.2, 1.0], [0.2, 0.8, 1.0], [0.8, 0.8, 1.0]],      # in the face z = 1 shared by two cells
        [[1.0, 1.0, 1.0], [1.5, 1.0, 1.2], [1.0, 1.5, 1.2], [1.6, 1.6, 1.5]],      # one vertex on a grid vertex
        [[0.5, 0.5, 0.5], [2.5, 0.5, 0.5], [0.5, 2.5, 0.5], [2.5, 2.5, 0.5]],      # flat, crosses 9 cells, edges off the grid
        [[1.0, 0.0, 0.0], [1.0, 3.0, 0.0], [1.0, 0.0, 3.0], [1.0, 3.0, 3.0]],      # the whole plane x = 1 (vertical: projects to a line)
        [[0.1, 0.2, 2.3], [0.9, 0.1, 2.6], [0.2, 0.8, 2.4], [0.8, 0.9, 2.9]],      # generic, inside one cell
        [[2.0, 2.0, 2.0], [3.0, 2.0, 2.0], [2.0, 3.0, 2.0], [3.0, 3.0, 2.5]],      # bottom edge... two vertices in the face z = 2
    ], dtype=torch.float64)
    im = ImmersedGrid(2, 3, quads, torch.arange(quads.shape[0], dtype=torch.int32)[:, None], quads.shape[0])
    return bg, im
