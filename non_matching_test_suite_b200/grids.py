"""Synthetic grids that mirror the geometry the reference's LM drivers build.

Nothing here is on the hot path: it produces the *inputs* (cell vertices, DoF
indices, hanging-node constraint lines) that a deal.II wrapper would extract
from `Triangulation`/`DoFHandler`/`AffineConstraints` objects.  It is written
on torch ops only so the same code runs on the CPU (tests) and on the GPU
(bench.py, where 10^7..10^8 cells must be produced in seconds).

Reference recipe followed (file:line relative to /root/reference/code):
  * background  `GridGenerator::hyper_cube(-1,1)` + `refine_global(4)`
      examples/lagrange_multiplier/2d3d/lagrange_multiplier.cc:258-263
  * band refinement rounds ("Space pre refinements cycles"): every active cell
    whose closed bounding box intersects an immersed cell's bounding box is
    refined                          .../2d3d/lagrange_multiplier.cc:376-394
  * immersed post refinement happens after the band rounds   :397-401
  * one `refine_global(1)` of both grids per cycle           :775-778
  * immersed grids: hyper_sphere<1,2> R=.3 c=(.5,.5)  1d2d/lagrange_multiplier.cc:297-304
                    flower polygon grids/flower_interface.vtk (256 segments)
                    hyper_sphere<2,3> R=.3 c=(.5,.5,.5)  2d3d/lagrange_multiplier.cc:261-265
  * Q1 background DoFs, hanging-node + Dirichlet constraint lines
      .../2d3d/lagrange_multiplier.cc:413-427
deal.II enforces at most one level of difference across faces (and edges in
3D); the generator applies the same 2:1 balance.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import torch

I64 = torch.int64
F64 = torch.float64


# --------------------------------------------------------------------------- #
# containers
# --------------------------------------------------------------------------- #
@dataclass
class BackgroundGrid:
    """Axis-aligned quad/hex leaves of a 2:1 balanced tree on [-1,1]^d with Q1 DoFs."""

    spacedim: int
    lo: torch.Tensor        # [n, d] f64  lower corner
    hi: torch.Tensor        # [n, d] f64  upper corner
    dofs: torch.Tensor      # [n, 2^d] int32, deal.II lexicographic vertex order
    n_dofs: int
    level: torch.Tensor     # [n] int32
    # constraint lines, CSR over the sorted list of constrained DoFs
    c_dof: torch.Tensor     # [n_c] int32, ascending
    c_ptr: torch.Tensor     # [n_c+1] int64
    c_master: torch.Tensor  # [nnz_c] int32
    c_weight: torch.Tensor  # [nnz_c] f64
    dof_coords: Optional[torch.Tensor] = None  # [n_dofs, d] f64 (for tests)

    @property
    def n_cells(self) -> int:
        return int(self.lo.shape[0])

    def vertices(self) -> torch.Tensor:
        """[n, 2^d, d] cell-major vertex coordinates in deal.II order
        (vertex v has bit k of v selecting hi along axis k)."""
        d = self.spacedim
        out = torch.empty((self.n_cells, 1 << d, d), dtype=F64, device=self.lo.device)
        for v in range(1 << d):
            for k in range(d):
                out[:, v, k] = self.hi[:, k] if (v >> k) & 1 else self.lo[:, k]
        return out

    def to(self, device) -> "BackgroundGrid":
        kw = {}
        for name, val in self.__dict__.items():
            kw[name] = val.to(device) if isinstance(val, torch.Tensor) else val
        return BackgroundGrid(**kw)


@dataclass
class ImmersedGrid:
    """Codimension-1 cells: segments in 2D (dim=1) or quads in 3D (dim=2); FE_DGQ(0) DoFs.  Codimension 0 in 2D
    (dim = spacedim = 2, the reference's <2,2,2> instantiation): convex quads in the plane of the background."""

    dim: int
    spacedim: int
    verts: torch.Tensor     # [n, 2^dim, spacedim] f64, deal.II lexicographic order
    dofs: torch.Tensor      # [n, dofs_per_cell] int32
    n_dofs: int

    @property
    def n_cells(self) -> int:
        return int(self.verts.shape[0])

    def boxes(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.verts.min(dim=1).values, self.verts.max(dim=1).values

    def measure(self) -> torch.Tensor:
        """Per-cell measure of the geometry the reference integrates over: segment
        length, or the area of the two flat triangles of the quad split along
        diagonal (v0,v3) -- use `measure_split` when the split is known."""
        v = self.verts
        if self.dim == 1:
            return (v[:, 1] - v[:, 0]).norm(dim=1)
        if self.spacedim == 2:          # planar quad, deal.II vertex order: shoelace over v0, v1, v3, v2
            r = v[:, [0, 1, 3, 2]]
            n = torch.roll(r, -1, dims=1)
            return 0.5 * (r[..., 0] * n[..., 1] - n[..., 0] * r[..., 1]).sum(dim=1).abs()
        a = torch.linalg.cross(v[:, 1] - v[:, 0], v[:, 3] - v[:, 0]).norm(dim=1)
        b = torch.linalg.cross(v[:, 3] - v[:, 0], v[:, 2] - v[:, 0]).norm(dim=1)
        return 0.5 * (a + b)

    def slice(self, first: int, last: int) -> "ImmersedGrid":
        return ImmersedGrid(self.dim, self.spacedim, self.verts[first:last],
                            self.dofs[first:last], self.n_dofs)

    def to(self, device) -> "ImmersedGrid":
        return ImmersedGrid(self.dim, self.spacedim, self.verts.to(device),
                            self.dofs.to(device), self.n_dofs)


# --------------------------------------------------------------------------- #
# immersed grids
# --------------------------------------------------------------------------- #
def circle_grid(n_segments: int, R: float = 0.3, center=(0.5, 0.5), device="cpu") -> ImmersedGrid:
    """hyper_sphere<1,2>: vertices on the circle, first vertex at angle pi/4."""
    k = torch.arange(n_segments + 1, dtype=F64, device=device)
    th = math.pi / 4 + 2.0 * math.pi * k / n_segments
    p = torch.stack([center[0] + R * torch.cos(th), center[1] + R * torch.sin(th)], dim=1)
    p[-1] = p[0]  # close the loop bit-exactly
    verts = torch.stack([p[:-1], p[1:]], dim=1)
    dofs = torch.arange(n_segments, dtype=torch.int32, device=device)[:, None]
    return ImmersedGrid(1, 2, verts.contiguous(), dofs, n_segments)


def solid_grid_2d(n: int, side: float = 0.6, angle: float = 0.3, center=(0.5, 0.5), warp: float = 0.04, device="cpu") -> ImmersedGrid:
    """A 2D body for the codimension-0 instantiation <2,2,2>: n x n convex quads covering a square of edge `side`, rotated
    by `angle` about `center` and smoothly warped (non-affine cells), one FE_DGQ(0) dof per cell."""
    t = torch.linspace(-0.5, 0.5, n + 1, dtype=F64, device=device)
    X, Y = torch.meshgrid(t, t, indexing="xy")                                # X varies along the second index (x fastest)
    Xw = X + warp * torch.sin(math.pi * (Y + 0.5)) * torch.sin(2 * math.pi * (X + 0.5))
    Yw = Y + warp * torch.sin(math.pi * (X + 0.5)) * torch.sin(2 * math.pi * (Y + 0.5))
    c, s_ = math.cos(angle), math.sin(angle)
    P = torch.stack([center[0] + side * (c * Xw - s_ * Yw), center[1] + side * (s_ * Xw + c * Yw)], dim=-1)   # [n+1 (y), n+1 (x), 2]
    v = torch.stack([P[:-1, :-1], P[:-1, 1:], P[1:, :-1], P[1:, 1:]], dim=2).reshape(n * n, 4, 2)            # deal.II order
    dofs = torch.arange(n * n, dtype=torch.int32, device=device)[:, None]
    return ImmersedGrid(2, 2, v.contiguous(), dofs, n * n)


def flower_points(device="cpu") -> torch.Tensor:
    """The 256 polygon vertices of grids/flower_interface.vtk: r = .3 + .1 cos(6 th) about
    (.5,.5), th_k = 2 pi (k+1)/256, coordinates as stored in the file (6 significant digits)."""
    pts = []
    for k in range(256):
        th = 2.0 * math.pi * (k + 1) / 256
        r = 0.3 + 0.1 * math.cos(6 * th)
        pts.append((float("%.6g" % (0.5 + r * math.cos(th))), float("%.6g" % (0.5 + r * math.sin(th)))))
    return torch.tensor(pts, dtype=F64, device=device)


def flower_grid(n_bisections: int = 0, device="cpu") -> ImmersedGrid:
    """Flat-manifold refinement: every cycle splits each segment at its midpoint."""
    p = flower_points(device)
    for _ in range(n_bisections):
        nxt = torch.roll(p, -1, dims=0)
        mid = 0.5 * (p + nxt)
        p = torch.stack([p, mid], dim=1).reshape(-1, 2)
    verts = torch.stack([p, torch.roll(p, -1, dims=0)], dim=1)
    n = verts.shape[0]
    dofs = torch.arange(n, dtype=torch.int32, device=device)[:, None]
    return ImmersedGrid(1, 2, verts.contiguous(), dofs, n)


def sphere_grid(n_per_edge: int, R: float = 0.3, center=(0.5, 0.5, 0.5), device="cpu",
                patches=range(6), rows: Optional[int] = None) -> ImmersedGrid:
    """hyper_sphere<2,3>: six patches of the projected cube, n x n bilinear quads each with
    vertices on the sphere (equi-angular spacing along patch edges).  `rows` keeps only the first
    rows of every patch (a sample of the grid, bit-identical to the same cells of the whole one)."""
    n = n_per_edge
    t = torch.tan(math.pi / 4 * (2.0 * torch.arange(n + 1, dtype=F64, device=device) / n - 1.0))
    t[0], t[-1] = -1.0, 1.0
    if n % 2 == 0:
        t[n // 2] = 0.0
    a, b = torch.meshgrid(t if rows is None else t[:min(n, rows) + 1], t, indexing="ij")
    one = torch.ones_like(a)
    faces = [
        (one, a, b), (-one, b, a),      # +x, -x
        (b, one, a), (a, -one, b),      # +y, -y
        (a, b, one), (b, a, -one),      # +z, -z
    ]
    c = torch.tensor(center, dtype=F64, device=device)
    all_verts = []
    for f in patches:
        x, y, z = faces[f]
        d = torch.stack([x, y, z], dim=-1)
        d = d / d.norm(dim=-1, keepdim=True)
        p = c + R * d                                       # [n+1, n+1, 3]
        q = torch.stack([p[:-1, :-1], p[1:, :-1], p[:-1, 1:], p[1:, 1:]], dim=2)  # [n,n,4,3]
        all_verts.append(q.reshape(-1, 4, 3))
    verts = torch.cat(all_verts, dim=0).contiguous()
    m = verts.shape[0]
    dofs = torch.arange(m, dtype=torch.int32, device=device)[:, None]
    return ImmersedGrid(2, 3, verts, dofs, m)


# --------------------------------------------------------------------------- #
# tree helpers
# --------------------------------------------------------------------------- #
def _cell_coord(idx: torch.Tensor, level: int) -> torch.Tensor:
    """Exact coordinate of grid line `idx` at `level` on [-1,1]: -1 + idx * 2^(1-level)."""
    return idx.to(F64) * (2.0 ** (1 - level)) - 1.0


def _floor_index(x: torch.Tensor, level: int) -> torch.Tensor:
    """Largest i with coord(i) <= x (exact on the raw doubles; unclamped)."""
    s = 2.0 ** (1 - level)
    i = torch.floor((x + 1.0) / s).to(I64)
    for _ in range(2):
        i = torch.where(_cell_coord(i, level) > x, i - 1, i)
        i = torch.where(_cell_coord(i + 1, level) <= x, i + 1, i)
    return i


def overlapping_cells(blo: torch.Tensor, bhi: torch.Tensor, level: int, chunk: int = 1 << 20) -> torch.Tensor:
    """Unique integer positions [m, d] of the level-`level` cells whose CLOSED box
    intersects at least one of the boxes [blo, bhi] (comparison on the raw doubles)."""
    d = blo.shape[1]
    n = 1 << level
    outs = []
    for s0 in range(0, blo.shape[0], chunk):
        lo_c, hi_c = blo[s0:s0 + chunk], bhi[s0:s0 + chunk]
        i0 = _floor_index(lo_c, level)
        i0 = torch.where(_cell_coord(i0, level) == lo_c, i0 - 1, i0)  # touching the previous cell
        j0 = _floor_index(hi_c, level)
        i0 = i0.clamp(0, n - 1)
        j0 = j0.clamp(0, n - 1)
        K = int((j0 - i0).max().item()) + 1
        offs = torch.cartesian_prod(*[torch.arange(K, device=blo.device)] * d).reshape(-1, d)
        idx = i0[:, None, :] + offs[None, :, :]
        ok = (idx <= j0[:, None, :]).all(dim=2)
        idx = idx[ok]
        outs.append(torch.unique(_pack(idx), sorted=True))
    keys = torch.unique(torch.cat(outs)) if len(outs) > 1 else outs[0]
    return _unpack(keys, d)


def _bits(d: int) -> int:
    return 31 if d == 2 else 21


def _pack(idx: torch.Tensor) -> torch.Tensor:
    """[m, d] non-negative ints (< 2^31 in 2D, < 2^21 in 3D) -> one int64 key, last axis most significant."""
    b = _bits(idx.shape[1])
    key = idx[:, 0].clone()
    for k in range(1, idx.shape[1]):
        key = key | (idx[:, k] << (b * k))
    return key


def _unpack(key: torch.Tensor, d: int) -> torch.Tensor:
    b = _bits(d)
    mask = (1 << b) - 1
    return torch.stack([(key >> (b * k)) & mask for k in range(d)], dim=1)


def _morton(idx: torch.Tensor) -> torch.Tensor:
    d = idx.shape[1]
    code = torch.zeros(idx.shape[0], dtype=I64, device=idx.device)
    for b in range(_bits(d) - 1, -1, -1):
        for k in range(d - 1, -1, -1):
            code = (code << 1) | ((idx[:, k] >> b) & 1)
    return code


@dataclass
class LevelMap:
    """Dense map, at the finest cycle-0 level `lf`, of the leaf level covering each position."""

    d: int
    lf: int
    lvl: torch.Tensor  # [2^lf]*d int16

    def lookup(self, pos: torch.Tensor, level: int) -> torch.Tensor:
        """Cycle-0 leaf level at integer positions given at `level` >= lf."""
        p = pos >> (level - self.lf) if level >= self.lf else pos << (self.lf - level)
        p = p.clamp(0, (1 << self.lf) - 1)
        if self.d == 2:
            return self.lvl[p[:, 0], p[:, 1]].to(I64)
        return self.lvl[p[:, 0], p[:, 1], p[:, 2]].to(I64)


def build_level_map(d: int, initial_level: int, band_rounds: int, blo: torch.Tensor, bhi: torch.Tensor,
                    balance: bool = True) -> LevelMap:
    """Cycle-0 background tree: uniform `initial_level`, then `band_rounds` rounds of
    "refine every leaf whose closed box meets an immersed box", with 2:1 balance."""
    lf = initial_level + band_rounds
    n = 1 << lf
    dev = blo.device
    lvl = torch.full((n,) * d, initial_level, dtype=torch.int16, device=dev)
    for r in range(band_rounds):
        cur = initial_level + r
        pos = overlapping_cells(blo, bhi, cur)
        # by induction every leaf meeting an immersed box sits at level `cur`
        w = 1 << (lf - cur)
        mark = torch.zeros((1 << cur,) * d, dtype=torch.bool, device=dev)
        mark[tuple(pos[:, k] for k in range(d))] = True
        for k in range(d):
            mark = mark.repeat_interleave(w, dim=k)
        lvl = torch.where(mark, torch.tensor(cur + 1, dtype=torch.int16, device=dev), lvl)
        if balance:
            lvl = _balance(lvl, d, initial_level, lf)
    return LevelMap(d, lf, lvl)


def graded_config(name: str, cycle: int, stripe: int = 8, band_only: bool = True, device="cpu", dirichlet: bool = True):
    """A variant of config `name` whose background changes level ACROSS the interface: one more refinement round is applied
    only around every other block of `stripe` immersed cells (of the coarse grid the band rounds see), so that hanging nodes
    sit on cut cells all along the interface -- the adaptive refinement the drivers offer (`Use space refinement`), which
    the shipped parameter files leave off.  Exercises the constraint expansion of the assembly at scale."""
    cfg = CONFIGS[name]
    d = cfg["d"]
    if name == "disk":
        im0, im = circle_grid(32, device=device), circle_grid(32 * (1 << cycle), device=device)
    elif name == "flower":
        im0, im = flower_grid(0, device=device), flower_grid(cycle, device=device)
    else:
        im0, im = sphere_grid(2, device=device), sphere_grid(4 * (1 << cycle), device=device)
    blo0, bhi0 = im0.boxes()
    rounds = cfg["band_rounds"]
    lf = 4 + rounds + 1
    n = 1 << lf
    lvl = torch.full((n,) * d, 4, dtype=torch.int16, device=blo0.device)
    pick = (torch.arange(im0.n_cells, device=blo0.device) // stripe) % 2 == 0
    for r in range(rounds + 1):
        cur = 4 + r
        lo_r, hi_r = (blo0, bhi0) if r < rounds else (blo0[pick], bhi0[pick])
        pos = overlapping_cells(lo_r, hi_r, cur)
        w = 1 << (lf - cur)
        mark = torch.zeros((1 << cur,) * d, dtype=torch.bool, device=blo0.device)
        mark[tuple(pos[:, k] for k in range(d))] = True
        for k in range(d):
            mark = mark.repeat_interleave(w, dim=k)
        # the extra round refines only leaves that did reach the band level
        lvl = torch.where(mark & (lvl == cur), torch.tensor(cur + 1, dtype=torch.int16, device=blo0.device), lvl)
        lvl = _balance(lvl, d, 4, lf)
    lmap = LevelMap(d, lf, lvl)
    bg = background_grid(lmap, cycle, near=im.boxes() if band_only else None, dirichlet=dirichlet)
    return bg, im


def _balance(lvl: torch.Tensor, d: int, l0: int, lf: int) -> torch.Tensor:
    """Enforce |level difference| <= 1 across faces (2D, 3D) and edges (3D)."""
    import torch.nn.functional as F
    changed = True
    while changed:
        changed = False
        x = lvl.to(torch.float32)[None, None]
        # neighbourhood max over face (+edge) neighbours at finest resolution
        shifts = []
        rng = (-1, 0, 1)
        if d == 2:
            shifts = [(a, b) for a in rng for b in rng if abs(a) + abs(b) == 1]
        else:
            shifts = [(a, b, c) for a in rng for b in rng for c in rng if 1 <= abs(a) + abs(b) + abs(c) <= 2]
        nb = lvl.clone()
        for sh in shifts:
            rolled = lvl
            for k, s in enumerate(sh):
                if s == 0:
                    continue
                rolled = torch.roll(rolled, shifts=s, dims=k)
                # undo wrap-around: the rolled-in slab keeps its own value
                sl = [slice(None)] * d
                sl[k] = slice(0, 1) if s == 1 else slice(-1, None)
                rolled = rolled.clone()
                rolled[tuple(sl)] = lvl[tuple(sl)]
            nb = torch.maximum(nb, rolled)
        for cur in range(l0, lf - 1):
            # leaves at level `cur` with a neighbour at level >= cur+2 get refined once
            w = 1 << (lf - cur)
            blk = nb
            is_cur = lvl == cur
            viol = is_cur & (nb >= cur + 2)
            if not bool(viol.any()):
                continue
            # block-wise OR at level cur
            v = viol
            for k in range(d):
                shp = list(v.shape)
                shp[k] = shp[k] // w
                shp.insert(k + 1, w)
                v = v.reshape(shp).any(dim=k + 1, keepdim=True).expand(shp).reshape(viol.shape)
            lvl = torch.where(v & is_cur, torch.tensor(cur + 1, dtype=lvl.dtype, device=lvl.device), lvl)
            changed = True
            break
    return lvl


# --------------------------------------------------------------------------- #
# background grid
# --------------------------------------------------------------------------- #
def background_grid(lmap: LevelMap, cycles: int, near: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                    dirichlet: bool = True, key_exchange=None, want_coords: bool = True) -> BackgroundGrid:
    """Leaves of the cycle-0 tree refined uniformly `cycles` times.

    near=None       -> every leaf (what the reference holds; small cycles only)
    near=(blo,bhi)  -> band-only: just the leaves whose closed box meets one of the boxes
                       (the only cells the coupling path ever touches), plus the master
                       vertices of their hanging nodes.
    """
    d, lf = lmap.d, lmap.lf
    dev = lmap.lvl.device
    Lfine = lf + cycles
    assert Lfine <= _bits(d) - 1, "vertex keys are packed in %d bits per axis" % _bits(d)
    levels0 = [int(v) for v in torch.unique(lmap.lvl).tolist()]
    pos_l, lev_l = [], []
    for l0 in levels0:
        L = l0 + cycles
        if near is None:
            w = 1 << (lf - l0)
            sub = lmap.lvl[tuple([slice(None, None, w)] * d)]
            p0 = torch.nonzero(sub == l0).to(I64)                      # positions at level l0
            if cycles:
                offs = torch.cartesian_prod(*[torch.arange(1 << cycles, device=dev)] * d).reshape(-1, d)
                p = ((p0 << cycles)[:, None, :] + offs[None, :, :]).reshape(-1, d)
            else:
                p = p0
        else:
            p = overlapping_cells(near[0], near[1], L)
            keep = lmap.lookup(p, L) == l0
            p = p[keep]
        if p.shape[0] == 0:
            continue
        order = torch.argsort(_morton(p))
        pos_l.append(p[order] << (Lfine - L))                          # lower corner in fine units
        lev_l.append(torch.full((p.shape[0],), L, dtype=I64, device=dev))
    pos = torch.cat(pos_l)
    lev = torch.cat(lev_l)
    size = (1 << (Lfine - lev))                                         # edge length in fine units
    # vertices, deal.II lexicographic: bit k of v -> +size along axis k
    nv = 1 << d
    corner = torch.tensor([[(v >> k) & 1 for k in range(d)] for v in range(nv)], dtype=I64, device=dev)
    vpos = pos[:, None, :] + corner[None, :, :] * size[:, None, None]   # [n, nv, d]
    vkey = _pack(vpos.reshape(-1, d))

    # hanging nodes of the vertices we hold; add their masters until closed
    all_keys = torch.unique(vkey)
    lines_dof, lines_master, lines_w = [], [], []
    frontier = all_keys
    while frontier.numel():
        hk, mk, mw = _hanging(frontier, lmap, cycles, Lfine)
        lines_dof.append(hk); lines_master.append(mk); lines_w.append(mw)
        new = torch.unique(mk)
        isin = torch.isin(new, all_keys)
        frontier = new[~isin]
        all_keys = torch.unique(torch.cat([all_keys, frontier]))
    if key_exchange is not None:
        # several ranks hold different parts of the band: number the union of their vertices
        all_keys = key_exchange(all_keys)
    n_dofs = int(all_keys.numel())
    dofs = torch.searchsorted(all_keys, vkey).reshape(-1, nv).to(torch.int32)
    con_dof = torch.searchsorted(all_keys, torch.cat(lines_dof))
    con_master = torch.searchsorted(all_keys, torch.cat(lines_master))
    con_w = torch.cat(lines_w)
    con_dof, con_master, con_w = _resolve_chains(con_dof, con_master, con_w)

    coords = _cell_coord(_unpack(all_keys, d), Lfine) if want_coords else None
    # Dirichlet lines (no masters) on the boundary of [-1,1]^d
    dir_dofs = torch.empty(0, dtype=I64, device=dev)
    if dirichlet:
        ik = _unpack(all_keys, d)
        on_b = ((ik == 0) | (ik == (1 << Lfine))).any(dim=1)
        del ik
        dir_dofs = torch.nonzero(on_b).flatten()
        # a hanging node on the boundary keeps its hanging-node line (deal.II: boundary
        # values do not overwrite existing lines)
        dir_dofs = dir_dofs[~torch.isin(dir_dofs, con_dof)]
        # closing the constraints resolves masters that are Dirichlet rows: they only feed the
        # inhomogeneity, so they leave the line (AffineConstraints::close())
        keep = ~torch.isin(con_master, dir_dofs)
        hang_dofs = torch.unique(con_dof)
        con_dof, con_master, con_w = con_dof[keep], con_master[keep], con_w[keep]
        dir_dofs = torch.unique(torch.cat([dir_dofs, hang_dofs]))   # lines that lost all masters stay as lines
    c_dof_all = torch.unique(torch.cat([con_dof, dir_dofs]))
    row = torch.searchsorted(c_dof_all, con_dof)
    order = torch.argsort(row * n_dofs + con_master)
    row, con_master, con_w = row[order], con_master[order], con_w[order]
    counts = torch.bincount(row, minlength=c_dof_all.numel())
    c_ptr = torch.zeros(c_dof_all.numel() + 1, dtype=I64, device=dev)
    c_ptr[1:] = torch.cumsum(counts, 0)

    lo = _cell_coord(pos, Lfine)
    hi = _cell_coord(pos + size[:, None], Lfine)
    return BackgroundGrid(d, lo, hi, dofs, n_dofs, lev.to(torch.int32), c_dof_all.to(torch.int32), c_ptr,
                          con_master.to(torch.int32), con_w, coords)


def _hanging(vkeys: torch.Tensor, lmap: LevelMap, cycles: int, Lfine: int):
    """For packed fine-level vertex keys return the hanging-node lines
    (dof key, master key, weight) of those that sit inside an edge/face of a coarser leaf."""
    d = lmap.d
    dev = vkeys.device
    V = _unpack(vkeys, d)
    nfine = 1 << Lfine
    best_size = torch.zeros(V.shape[0], dtype=I64, device=dev)
    for c in range(1 << d):
        delta = torch.tensor([(c >> k) & 1 for k in range(d)], dtype=I64, device=dev)
        p = V - delta                                         # fine cell touching the vertex
        inside = ((p >= 0) & (p < nfine)).all(dim=1)
        l0 = lmap.lookup(p.clamp(0, nfine - 1) >> cycles, lmap.lf)
        size = (1 << (Lfine - (l0 + cycles)))
        misaligned = ((V % size[:, None]) != 0).any(dim=1) & inside
        best_size = torch.where(misaligned & (size > best_size), size, best_size)
    hang = best_size > 0
    if not bool(hang.any()):
        e = torch.empty(0, dtype=I64, device=dev)
        return e, e, torch.empty(0, dtype=F64, device=dev)
    Vh, S = V[hang], best_size[hang]
    kh = vkeys[hang]
    off = (Vh % S[:, None]) != 0                              # axes along which the vertex is interior
    na = off.sum(dim=1)
    base = Vh - (Vh % S[:, None])
    out_d, out_m, out_w = [], [], []
    for c in range(1 << d):
        bits = torch.tensor([(c >> k) & 1 for k in range(d)], dtype=I64, device=dev)
        # a master for every choice of end point along the interior axes (bits on aligned axes must be 0)
        valid = ((bits[None, :] == 0) | off).all(dim=1)
        m = torch.where(off, base + bits[None, :] * S[:, None], Vh)
        out_d.append(kh[valid]); out_m.append(_pack(m[valid])); out_w.append((0.5 ** na[valid].to(F64)))
    return torch.cat(out_d), torch.cat(out_m), torch.cat(out_w)


def _resolve_chains(dof, master, w):
    """Substitute constrained masters by their own lines until every master is free."""
    for _ in range(8):
        cd = torch.unique(dof)
        bad = torch.isin(master, cd)
        if not bool(bad.any()):
            break
        order = torch.argsort(dof)
        sd, sm, sw = dof[order], master[order], w[order]
        start = torch.searchsorted(sd, master[bad])
        end = torch.searchsorted(sd, master[bad], right=True)
        cnt = end - start
        rep = torch.repeat_interleave(torch.arange(cnt.numel(), device=dof.device), cnt)
        within = torch.arange(rep.numel(), device=dof.device) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt)
        src = start[rep] + within
        nd, nm, nw = dof[bad][rep], sm[src], w[bad][rep] * sw[src]
        dof = torch.cat([dof[~bad], nd]); master = torch.cat([master[~bad], nm]); w = torch.cat([w[~bad], nw])
        # merge duplicates
        key = dof * (int(master.max().item()) + 1) + master
        uk, inv = torch.unique(key, return_inverse=True)
        ww = torch.zeros(uk.numel(), dtype=F64, device=dof.device).index_add_(0, inv, w)
        first = torch.zeros(uk.numel(), dtype=I64, device=dof.device).scatter_(0, inv, torch.arange(inv.numel(), device=dof.device))
        dof, master, w = dof[first], master[first], ww
    return dof, master, w


# --------------------------------------------------------------------------- #
# named configurations (BASELINE.json `configs`)
# --------------------------------------------------------------------------- #
CONFIGS = {
    # name: (spacedim, band rounds p, generator)
    "disk": dict(d=2, band_rounds=2),
    "flower": dict(d=2, band_rounds=4),
    "sphere": dict(d=3, band_rounds=1),
}


def make_config(name: str, cycle: int, band_only: Optional[bool] = None, device="cpu",
                n_override: Optional[int] = None, balance: bool = True,
                dirichlet: bool = True) -> Tuple[BackgroundGrid, ImmersedGrid]:
    """Grids of reference config `name` at refinement cycle `cycle` (0 = first table row).

    disk   : data/lm_1d2d_disk.{smooth,non_smooth}.prm   32 * 2^cycle segments
    flower : data/lm_1d2d_flower.smooth.prm              256 * 2^cycle segments
    sphere : data/lm_2d3d.{smooth,non_smooth}_sphere.prm 6 * (4 * 2^cycle)^2 quads
    `n_override` replaces the immersed resolution (segments, or quads per patch edge) while the
    background keeps the level of `cycle` -- used for weak-scaling shards.
    """
    cfg = CONFIGS[name]
    d = cfg["d"]
    if band_only is None:
        band_only = cycle > (4 if d == 2 else 2)
    if name == "disk":
        im0 = circle_grid(32, device=device)
        im = circle_grid(n_override or 32 * (1 << cycle), device=device)
    elif name == "flower":
        im0 = flower_grid(0, device=device)
        im = flower_grid(cycle, device=device) if n_override is None else flower_grid(n_override, device=device)
    elif name == "sphere":
        im0 = sphere_grid(2, device=device)          # band rounds see the grid before post refinement
        im = sphere_grid(n_override or 4 * (1 << cycle), device=device)
    else:
        raise KeyError(name)
    blo0, bhi0 = im0.boxes()
    lmap = build_level_map(d, 4, cfg["band_rounds"], blo0, bhi0, balance=balance)
    near = im.boxes() if band_only else None
    bg = background_grid(lmap, cycle, near=near, dirichlet=dirichlet)
    return bg, im


# --------------------------------------------------------------------------- #
# sharding by immersed-cell ranges
# --------------------------------------------------------------------------- #
def shard_range(n_cells: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced immersed-cell range of `rank`."""
    base, rem = divmod(n_cells, world)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def restrict_background(bg: BackgroundGrid, blo: torch.Tensor, bhi: torch.Tensor) -> BackgroundGrid:
    """Keep the cells whose closed box meets one of the boxes; DoF numbering and constraint
    lines stay global (every rank addresses the same background vector)."""
    keep = torch.zeros(bg.n_cells, dtype=torch.bool, device=bg.lo.device)
    for lvl in torch.unique(bg.level).tolist():
        sel = torch.nonzero(bg.level == lvl).flatten()
        pos = _floor_index(bg.lo[sel] + 0.5 * (bg.hi[sel] - bg.lo[sel]), int(lvl))
        near = _pack(overlapping_cells(blo, bhi, int(lvl)))
        keep[sel] = torch.isin(_pack(pos), near)
    idx = torch.nonzero(keep).flatten()
    return BackgroundGrid(bg.spacedim, bg.lo[idx], bg.hi[idx], bg.dofs[idx], bg.n_dofs, bg.level[idx], bg.c_dof, bg.c_ptr,
                          bg.c_master, bg.c_weight, bg.dof_coords)


def _exchange_keys(local: torch.Tensor) -> torch.Tensor:
    """Sorted union of every rank's sorted key list (torch.distributed must be initialised)."""
    import torch.distributed as dist
    world = dist.get_world_size()
    n = torch.tensor([local.numel()], dtype=I64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    cap = int(max(int(x.item()) for x in sizes))
    pad = torch.full((cap,), -1, dtype=I64, device=local.device)
    pad[:local.numel()] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    parts = [b[:int(k.item())] for b, k in zip(bufs, sizes)]
    return torch.unique(torch.cat(parts))


def weak_scaling_sizes(name: str, cycle: int, world: int) -> Tuple[int, Optional[int]]:
    """(refinement cycle, immersed resolution override) of the global problem on `world` GPUs.  Every size is a TRUE
    refinement cycle of the named configuration (BASELINE.json: "synthetic refinements of the named configs"): the
    sphere gains one cycle (4x the pairs) per factor 4 of GPUs, rounded up, so that 1 / 2 / 4 / 8 GPUs run cycles
    c / c+1 / c+1 / c+2 -- with c = 8 that is 6.4e7 pairs on one GPU (configs[3]) and 1.0e9 pairs on eight
    (configs[4]); a rank holds either the single-GPU number of immersed cells or twice that.  The 2D curves gain one
    cycle (2x the pairs) per factor 2."""
    if world == 1:
        return cycle, None
    if name == "sphere":
        k = int(math.ceil(math.log2(world) / 2.0 - 1e-9))
        return cycle + k, 4 * (1 << (cycle + k))
    k = int(math.ceil(math.log2(world) - 1e-9))
    if name == "disk":
        return cycle + k, 32 * (1 << (cycle + k))
    return cycle + k, cycle + k                        # flower: bisections only


def sample_problem(name: str, cycle: int, n_cells: int, device="cpu"):
    """A bounded SAMPLE of config `name` at refinement `cycle`: the first `n_cells` immersed cells of the true grid of
    that cycle and the band of background cells near them, at the true mesh sizes of that cycle (used by the CPU
    baseline / reference arm of bench.py, which cannot afford the whole workload)."""
    cfg = CONFIGS[name]
    d = cfg["d"]
    if name == "disk":
        im0, im = circle_grid(32, device=device), circle_grid(32 * (1 << cycle), device=device)
    elif name == "flower":
        im0, im = flower_grid(0, device=device), flower_grid(cycle, device=device)
    else:
        n = 4 * (1 << cycle)
        im0 = sphere_grid(2, device=device)
        if n_cells >= n * n:
            im = sphere_grid(n, device=device, patches=range(min(6, -(-n_cells // (n * n)))))
        else:
            im = sphere_grid(n, device=device, patches=[0], rows=-(-n_cells // n))   # the first rows of the +x patch
    n_global = {"disk": 32 * (1 << cycle), "flower": 256 * (1 << cycle), "sphere": 6 * (4 * (1 << cycle)) ** 2}[name]
    k = min(n_cells, im.n_cells)
    im = ImmersedGrid(im.dim, im.spacedim, im.verts[:k].clone(), im.dofs[:k].clone(), n_global)
    lmap = build_level_map(d, 4, cfg["band_rounds"], *im0.boxes())
    bg = background_grid(lmap, cycle, near=im.boxes(), want_coords=False)
    return bg, im


def shard_problem(name: str, cycle: int, rank: int = 0, world: int = 1, device="cpu", weak: bool = False,
                  want_coords: bool = True, standalone: bool = False):
    """This rank's share of config `name`: a contiguous range of immersed cells and the background
    cells near them, with global DoF numbering.

    weak=False: the global problem is config `name` at `cycle` (strong scaling; every rank generates the
                whole band and keeps its part).
    weak=True : the global problem grows with `world` (weak_scaling_sizes); a rank only generates the band
                near its own immersed cells and the ranks agree on the DoF numbering by exchanging their
                vertex keys (needs an initialised process group when world > 1)."""
    if not weak or world == 1:
        bg, im = make_config(name, cycle, band_only=True, device=device)
        info = {"global_background_cells": bg.n_cells, "global_immersed_cells": im.n_cells, "background_dofs": bg.n_dofs}
        if world > 1:
            first, last = shard_range(im.n_cells, rank, world)
            im = im.slice(first, last)
            blo, bhi = im.boxes()
            bg = restrict_background(bg, blo, bhi)
        info.update({"rank0_background_cells": bg.n_cells, "rank0_immersed_cells": im.n_cells})
        return bg, im, info
    cfg = CONFIGS[name]
    d = cfg["d"]
    bg_cycle, n_im = weak_scaling_sizes(name, cycle, world)
    if name == "disk":
        im0, im = circle_grid(32, device=device), circle_grid(n_im, device=device)
    elif name == "flower":
        im0, im = flower_grid(0, device=device), flower_grid(n_im, device=device)
    else:
        im0, im = sphere_grid(2, device=device), sphere_grid(n_im, device=device)
    n_global = im.n_cells
    first, last = shard_range(n_global, rank, world)
    im = ImmersedGrid(im.dim, im.spacedim, im.verts[first:last].clone(), im.dofs[first:last].clone(), im.n_dofs)
    lmap = build_level_map(d, 4, cfg["band_rounds"], *im0.boxes())
    # standalone: this rank's share generated without a process group (its own dof numbering): one rank of a
    # `world`-GPU run reproduced on a single GPU, e.g. to size its memory
    bg = background_grid(lmap, bg_cycle, near=im.boxes(), key_exchange=None if standalone else _exchange_keys, want_coords=want_coords)
    info = {"global_immersed_cells": n_global, "background_dofs": bg.n_dofs, "background_cycle": bg_cycle,
            "immersed_resolution": n_im, "rank0_background_cells": bg.n_cells, "rank0_immersed_cells": im.n_cells}
    return bg, im, info
