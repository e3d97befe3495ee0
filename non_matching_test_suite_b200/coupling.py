"""Host-side mirror of the reference's three entry points, on top of the C ABI.

Same names, argument meaning and error behaviour as
  dealii::NonMatching::collect_quadratures_on_overlapped_grids            (code/include/coupling_utilities.h:508-515)
  create_coupling_sparsity_pattern_with_exact_intersections               (code/include/coupling_utilities.h:28-38)
  create_coupling_mass_matrix_with_exact_intersections                    (code/include/coupling_utilities.h:154-168)
and, as the first widening row, of the interface-penalisation pair
  assemble_nitsche_with_exact_intersections                               (code/include/coupling_utilities.h:306-398)
  create_nitsche_rhs_with_exact_intersections                             (code/include/coupling_utilities.h:400-481)
with the deal.II objects replaced by the flat arrays a wrapper extracts from them
(`grids.BackgroundGrid` plays GridTools::Cache + DoFHandler + AffineConstraints of the space
grid, `grids.ImmersedGrid` those of the immersed grid).  The C++ drop-in header with the
reference's exact signatures is include/coupling_utilities.h; this module is what the tests
and the benchmark drive.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from . import _lib
from .grids import BackgroundGrid, ImmersedGrid


def _u32(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.int32).contiguous()


def make_context(space: BackgroundGrid, immersed: ImmersedGrid, device: int = 0, boxes: bool = False,
                 on_device: bool = False) -> _lib.Context:
    """Marshal both grids into a fresh context.  boxes=True sends lo/hi instead of the 2^d vertices;
    on_device=True hands over CUDA tensors (inputs already resident in HBM)."""
    ctx = _lib.Context(device)
    set_grids(ctx, space, immersed, boxes=boxes, on_device=on_device)
    return ctx


def set_grids(ctx: _lib.Context, space: BackgroundGrid, immersed: ImmersedGrid, boxes: bool = False, on_device: bool = False):
    dev = "cuda:%d" % ctx.device if on_device else "cpu"
    mv = lambda t: t.to(dev).contiguous()
    kw = dict(dofs=mv(_u32(space.dofs)), n_dofs=space.n_dofs, c_dof=mv(_u32(space.c_dof)), c_ptr=mv(space.c_ptr.to(torch.int64)),
              c_master=mv(_u32(space.c_master)), c_weight=mv(space.c_weight))
    if boxes:
        ctx.set_background(space.spacedim, lo=mv(space.lo), hi=mv(space.hi), **kw)
    else:
        ctx.set_background(space.spacedim, verts=mv(space.vertices()), **kw)
    ctx.set_immersed(immersed.dim, immersed.spacedim, mv(immersed.verts), mv(_u32(immersed.dofs)), immersed.n_dofs,
                     dofs_per_cell=int(immersed.dofs.shape[1]))


@dataclass
class IntersectionInfo:
    """The reference's std::vector<std::tuple<cell, cell, Quadrature<spacedim>>>, flattened:
    pair i = (space_cell[i], immersed_cell[i]) with points[q_offset[i]:q_offset[i+1]]."""

    ctx: _lib.Context
    space_cell: torch.Tensor
    immersed_cell: torch.Tensor
    sum_w: torch.Tensor
    q_offset: Optional[torch.Tensor] = None
    points: Optional[torch.Tensor] = None
    weights: Optional[torch.Tensor] = None
    counts: Optional[dict] = None

    def __len__(self):
        return int(self.space_cell.shape[0])

    def quadrature(self, i: int) -> Tuple[torch.Tensor, torch.Tensor]:
        a, b = int(self.q_offset[i]), int(self.q_offset[i + 1])
        return self.points[a:b], self.weights[a:b]


def collect_quadratures_on_overlapped_grids(space: BackgroundGrid, immersed: ImmersedGrid, degree: int, tol: float = 1e-14,
                                            ctx: Optional[_lib.Context] = None, with_quadrature: bool = True,
                                            device: int = 0) -> IntersectionInfo:
    """Reference: code/src/coupling_utilities.cc:22-69.  `degree` must be > 0 (AssertThrow there, ValueError here)."""
    if immersed.dim > space.spacedim:
        raise ValueError("Intrinsic dimension of the immersed object must be smaller than dim0.")
    if degree <= 0:
        raise ValueError("Invalid quadrature degree.")
    if ctx is None:
        ctx = make_context(space, immersed, device=device)
    counts = ctx.intersect(degree, tol)
    im, bg, sw = ctx.pairs()
    info = IntersectionInfo(ctx, bg, im, sw, counts=counts)
    if with_quadrature:
        info.q_offset, info.points, info.weights = ctx.quadrature()
    return info


def create_coupling_sparsity_pattern_with_exact_intersections(intersections_info: IntersectionInfo, space: BackgroundGrid,
                                                              immersed: ImmersedGrid):
    """Reference: code/include/coupling_utilities.h:28-152.  Returns the pattern of the stored
    n_space_dofs x n_immersed_dofs matrix as CSR (rowptr, col), columns ascending, constrained
    rows kept (keep_constrained_entries = true)."""
    ctx = intersections_info.ctx
    if ctx.n_space_dofs != space.n_dofs or ctx.n_im_dofs != immersed.n_dofs:
        raise ValueError("sparsity dimensions do not match the DoF handlers")      # AssertDimension :40-41
    rp, col, _ = ctx.csr(transpose=False, values=False)
    return rp, col


class CouplingMatrix:
    """The stored coupling matrix; vmult / Tvmult are the reference's `Ct` / `C` operators
    (code/examples/lagrange_multiplier/2d3d/lagrange_multiplier.cc:627-628)."""

    def __init__(self, ctx: _lib.Context):
        self.ctx = ctx
        self.m, self.n = ctx.n_space_dofs, ctx.n_im_dofs

    def vmult(self, src, dst=None):
        return self.ctx.spmv(src, transpose=False, out=dst)

    def Tvmult(self, src, dst=None):
        return self.ctx.spmv(src, transpose=True, out=dst)

    def csr(self, transpose=False, device="cpu"):
        return self.ctx.csr(transpose=transpose, values=True, device=device)


def create_coupling_mass_matrix_with_exact_intersections(space: BackgroundGrid, immersed: ImmersedGrid,
                                                         cells_and_quads: IntersectionInfo, use_given_quadrature: bool = True
                                                         ) -> CouplingMatrix:
    """Reference: code/include/coupling_utilities.h:154-304 (the function compresses the matrix itself).
    use_given_quadrature=True feeds the points/weights of `cells_and_quads` back through the
    inverse-mapping + shape-function kernel, exactly like the reference signature;
    False uses the fused device path (the points never left the GPU)."""
    ctx = cells_and_quads.ctx
    if ctx.n_space_dofs != space.n_dofs or ctx.n_im_dofs != immersed.n_dofs:
        raise ValueError("matrix dimensions do not match the DoF handlers")         # AssertDimension :170-171
    if use_given_quadrature and cells_and_quads.points is not None:
        ctx.assemble_from_quadrature(cells_and_quads.immersed_cell.contiguous(), cells_and_quads.space_cell.contiguous(),
                                     cells_and_quads.q_offset.contiguous(), cells_and_quads.points.contiguous(),
                                     cells_and_quads.weights.contiguous())
    else:
        ctx.assemble()
    return CouplingMatrix(ctx)


def _values_at(fn, points: torch.Tensor) -> Optional[torch.Tensor]:
    """Function<spacedim>::value_list of the reference: a callable on the [n_q, spacedim] points, a number, or None."""
    if fn is None:
        return None
    if callable(fn):
        v = fn(points)
    else:
        v = torch.full((points.shape[0],), float(fn), dtype=torch.float64, device=points.device)
    if tuple(v.shape) != (points.shape[0],):
        raise ValueError("function values must have one entry per quadrature point")
    return v.to(torch.float64).contiguous()


def assemble_nitsche_with_exact_intersections(space: BackgroundGrid, cells_and_quads: IntersectionInfo, nitsche_coefficient=1.0,
                                              penalty: float = 1.0):
    """Reference: code/include/coupling_utilities.h:306-398.  Returns the contribution
    sum_pairs sum_q c(x_q) penalty/h phi_i phi_j w_q, distributed through the space constraints, as CSR
    (rowptr, col, val) of an n_space x n_space matrix; the caller adds it to its system matrix, as the reference's
    distribute_local_to_global does in place (:387-388)."""
    ctx = cells_and_quads.ctx
    if ctx.n_space_dofs != space.n_dofs:
        raise ValueError("matrix dimensions do not match the DoF handler")                        # AssertDimension :319-320
    if len(cells_and_quads) == 0:
        raise ValueError("The background and immersed mesh must overlap in order to use this function.")   # Assert :323-325
    if cells_and_quads.points is None:
        raise ValueError("cells_and_quads carries no quadrature (collect with with_quadrature=True)")
    pts = cells_and_quads.points.contiguous()
    csr, _ = ctx.assemble_nitsche(cells_and_quads.space_cell.contiguous(), cells_and_quads.q_offset.contiguous(), pts,
                                  cells_and_quads.weights.contiguous(), _values_at(nitsche_coefficient, pts), None, penalty)
    return csr


def create_nitsche_rhs_with_exact_intersections(space: BackgroundGrid, cells_and_quads: IntersectionInfo, rhs_function,
                                                coefficient=1.0, penalty: float = 1.0) -> torch.Tensor:
    """Reference: code/include/coupling_utilities.h:400-481.  Returns the contribution to the right-hand side
    sum_pairs sum_q c(x_q) penalty/h phi_i f(x_q) w_q, distributed through the space constraints (:458-459)."""
    ctx = cells_and_quads.ctx
    if ctx.n_space_dofs != space.n_dofs:
        raise ValueError("vector size does not match the DoF handler")                            # AssertDimension :413
    if cells_and_quads.points is None:
        raise ValueError("cells_and_quads carries no quadrature (collect with with_quadrature=True)")
    pts = cells_and_quads.points.contiguous()
    if len(cells_and_quads) == 0:
        return torch.zeros(space.n_dofs, dtype=torch.float64, device=pts.device)
    _, rhs = ctx.assemble_nitsche(cells_and_quads.space_cell.contiguous(), cells_and_quads.q_offset.contiguous(), pts,
                                  cells_and_quads.weights.contiguous(), _values_at(coefficient, pts), _values_at(rhs_function, pts),
                                  penalty, matrix=False)
    return rhs


# --------------------------------------------------------------------------- #
# general tensor-product Lagrange elements and the drivers' solve()           #
# --------------------------------------------------------------------------- #
@dataclass
class FiniteElement:
    """What the coupling path needs to know about a deal.II FiniteElement: its degree and fe.get_unit_support_points(), in
    the element's own dof order (no numbering convention is assumed: the shape functions are the Lagrange polynomials of
    these points)."""

    degree: int
    unit_support_points: torch.Tensor       # [dofs_per_cell, dim] in [0, 1]^dim

    @property
    def dofs_per_cell(self) -> int:
        return int(self.unit_support_points.shape[0])


def lagrange_element(dim: int, degree: int, nodes_1d: Optional[torch.Tensor] = None) -> FiniteElement:
    """Tensor-product Lagrange element with lexicographic dof order (x fastest) on the given 1D nodes; default: equidistant
    nodes, the single midpoint for degree 0 (FE_DGQ(0))."""
    if nodes_1d is None:
        nodes_1d = torch.tensor([0.5], dtype=torch.float64) if degree == 0 else torch.linspace(0.0, 1.0, degree + 1, dtype=torch.float64)
    g = torch.meshgrid(*([nodes_1d] * dim), indexing="ij")
    pts = torch.stack([a.reshape(-1) for a in reversed(g)], dim=1)          # first coordinate runs fastest
    return FiniteElement(degree, pts.contiguous())


def create_coupling_mass_matrix_for_elements(space: BackgroundGrid, immersed: ImmersedGrid, cells_and_quads: IntersectionInfo,
                                             space_fe: FiniteElement, space_dofs: torch.Tensor, n_space_dofs: int,
                                             immersed_fe: FiniteElement, immersed_dofs: torch.Tensor, n_immersed_dofs: int,
                                             constraints=None, use_given_quadrature: bool = False) -> CouplingMatrix:
    """create_coupling_mass_matrix_with_exact_intersections (code/include/coupling_utilities.h:154-304) for FE_Q(p) on the
    background and FE_DGQ(q) on the immersed grid: `space_dofs` [n_cells, dofs_per_cell] / `immersed_dofs` are what
    cell->get_dof_indices returns, `constraints` = (c_dof, c_ptr, c_master, c_weight) the lines of the closed
    AffineConstraints (default: those of `space`, which fit its Q1 numbering only)."""
    ctx = cells_and_quads.ctx
    if space_dofs.shape[0] != space.n_cells or immersed_dofs.shape[0] != immersed.n_cells:
        raise ValueError("DoF tables do not match the grids")
    if constraints is None:
        constraints = (None, None, None, None)
    ctx.set_finite_elements((space_fe.degree, space_fe.unit_support_points.numpy()), _u32(space_dofs), n_space_dofs,
                            (immersed_fe.degree, immersed_fe.unit_support_points.numpy()), _u32(immersed_dofs), n_immersed_dofs, *constraints)
    if use_given_quadrature and cells_and_quads.points is not None:
        ctx.assemble_fe(cells_and_quads.immersed_cell.contiguous(), cells_and_quads.space_cell.contiguous(), cells_and_quads.q_offset.contiguous(),
                        cells_and_quads.points.contiguous(), cells_and_quads.weights.contiguous())
    else:
        ctx.assemble_fe()
    return CouplingMatrix(ctx)


def solve(coupling_matrix: CouplingMatrix, stiffness_matrix, mass_matrix, space_rhs: torch.Tensor, embedded_rhs: torch.Tensor,
          inner=(2000, 1e-14, 1e-12), outer=(0, 1e-11, 1e-2), inhomogeneity: Optional[torch.Tensor] = None, gmres_basis: int = 28):
    """PoissonLM::solve() (code/examples/lagrange_multiplier/2d3d/lagrange_multiplier.cc:613-659) around the GPU products of
    `coupling_matrix`: lambda = S^-1 (C K^-1 f - g), S = C K^-1 Ct by GMRES preconditioned with C K Ct + M, K^-1 by CG, then
    solution = K^-1 (f - Ct lambda) with the space constraints distributed.  `stiffness_matrix`, `mass_matrix` = CSR triples
    (rowptr int64, col int32, val float64).  Returns (lambda, solution, stats)."""
    return coupling_matrix.ctx.schur_solve(stiffness_matrix, mass_matrix, space_rhs, embedded_rhs, inner=inner, outer=outer,
                                           inhomogeneity=inhomogeneity, gmres_basis=gmres_basis)
