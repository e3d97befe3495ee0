"""GPU tests of what round 2 added: the brick candidate index against the chain index, the compact read-back and the
products in local numbering against the global forms, the asynchronous read-back, the merge-path flag the advisor
found uninitialised, and the multi-GPU check as a test (skipped below 2 GPUs)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from non_matching_test_suite_b200 import _lib, coupling
from non_matching_test_suite_b200.grids import make_config, circle_grid

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _keys(im, bg):
    return (im.numpy().astype(np.uint64) << np.uint64(32)) | bg.numpy().astype(np.uint32).astype(np.uint64)


@pytest.mark.parametrize("name,cycle,band", [("disk", 3, False), ("flower", 2, False), ("sphere", 1, False), ("sphere", 4, True), ("flower", 8, True)])
def test_brick_index_equals_chain_index(monkeypatch, c_oracle, name, cycle, band):
    """Grid-aligned trees take the brick index; its candidate list is the chain index's, entry by entry, and both are
    the oracle's closed-box set."""
    bg, im = make_config(name, cycle, band_only=band)
    ctx = coupling.make_context(bg, im, boxes=True)
    ctx.intersect(3, 1e-15)
    assert ctx.info("index_kind") == 2
    k_brick = _keys(*ctx.candidates())
    ctx.close()
    monkeypatch.setenv("NMGPU_INDEX", "chains")
    ctx = coupling.make_context(bg, im, boxes=True)
    ctx.intersect(3, 1e-15)
    assert ctx.info("index_kind") == 1
    k_chain = _keys(*ctx.candidates())
    ctx.close()
    assert np.array_equal(k_brick, k_chain)
    ilo, ihi = im.boxes()
    assert np.array_equal(k_brick, c_oracle.candidates(bg.lo, bg.hi, ilo, ihi))


def test_brick_index_refuses_what_it_cannot_verify():
    """Boxes that are not cells of one grid (random sizes, overlapping) fail the per-cell verification and the chain
    index takes over without the caller noticing."""
    from test_gpu_parity import _random_problem
    bg, im = _random_problem(3, 2000, 300, seed=5)
    ctx = coupling.make_context(bg, im, boxes=True)
    c = ctx.intersect(3, 1e-15)
    assert ctx.info("index_kind") == 1 and c["n_candidates"] > 1000
    ctx.close()
    # a tree whose coordinates are NOT dyadic: [0, 0.3]^2 refined -- cells are grid cells only up to rounding
    n = 37
    g = 0.3 / n
    ii, jj = torch.meshgrid(torch.arange(n, dtype=torch.float64), torch.arange(n, dtype=torch.float64), indexing="ij")
    lo = torch.stack([ii.reshape(-1) * g, jj.reshape(-1) * g], 1)
    hi = torch.stack([(ii.reshape(-1) + 1) * g, (jj.reshape(-1) + 1) * g], 1)
    from non_matching_test_suite_b200.grids import BackgroundGrid
    e = torch.empty(0, dtype=torch.int32)
    dofs = torch.arange(n * n * 4, dtype=torch.int32).reshape(-1, 4)
    bg2 = BackgroundGrid(2, lo, hi, dofs, n * n * 4, torch.zeros(n * n, dtype=torch.int32), e, torch.zeros(1, dtype=torch.int64), e,
                         torch.empty(0, dtype=torch.float64))
    im2 = circle_grid(64, R=0.1, center=(0.15, 0.15))
    ctx = coupling.make_context(bg2, im2, boxes=True)
    ctx.intersect(3, 1e-15)
    got = _keys(*ctx.candidates())
    ilo, ihi = im2.boxes()
    ov = ((bg2.lo[None] <= ihi[:, None]) & (ilo[:, None] <= bg2.hi[None])).all(-1)        # [im, bg] brute force on the raw doubles
    want = np.array([(m << 32) | b for m, b in torch.nonzero(ov).tolist()], dtype=np.uint64)
    assert np.array_equal(got, want)          # whichever index ran, the set is the closed-box set
    ctx.close()


@pytest.mark.parametrize("name,cycle", [("disk", 2), ("sphere", 2)])
@pytest.mark.parametrize("compact", ["0", "1"])
def test_compact_readback_and_local_products(monkeypatch, name, cycle, compact):
    monkeypatch.setenv("NMGPU_COMPACT_ROWS", compact)
    bg, im = make_config(name, cycle, band_only=True)
    ctx = coupling.make_context(bg, im, boxes=True)
    ctx.intersect(3, 1e-15)
    rp, col, val = ctx.csr()
    ids, ln, col_c, val_c = ctx.csr_compact()
    lens = (rp[1:] - rp[:-1])
    nz = torch.nonzero(lens).flatten()
    assert torch.equal(ids.long(), nz) and torch.equal(ln.long(), lens[nz])
    assert torch.equal(col_c, col) and torch.equal(val_c, val)
    assert ctx.n_rows_local() == nz.numel() and ctx.info("compact_rows") == int(compact)
    # products in local numbering = the global products restricted to the local entries
    g = torch.Generator().manual_seed(3)
    x_loc = torch.rand(im.n_cells, dtype=torch.float64, generator=g)
    u_loc = torch.rand(nz.numel(), dtype=torch.float64, generator=g)
    x_glob = torch.zeros(im.n_dofs, dtype=torch.float64); x_glob[im.dofs.long().flatten()] = x_loc
    u_glob = torch.zeros(bg.n_dofs, dtype=torch.float64); u_glob[nz] = u_loc
    y = ctx.spmv(x_glob, transpose=False)
    z = ctx.spmv(u_glob, transpose=True)
    y_loc = ctx.spmv_local(x_loc, transpose=False)
    z_loc = ctx.spmv_local(u_loc, transpose=True)
    assert torch.equal(y_loc, y[nz]) and torch.equal(z_loc, z[im.dofs.long().flatten()])
    # device vectors, and the scatter / gather helpers
    yd = ctx.spmv_local(x_loc.cuda(), transpose=False)
    assert torch.equal(yd.cpu(), y_loc)
    gl = torch.full((bg.n_dofs,), -7.0, dtype=torch.float64, device="cuda")
    ctx.local_to_global(_lib.NM_LOCAL_ROWS, u_loc, gl)
    assert torch.equal(gl.cpu()[nz], u_loc) and float((gl.cpu() == -7.0).sum()) == bg.n_dofs - nz.numel()
    back = torch.empty(nz.numel(), dtype=torch.float64)
    ctx.global_to_local(_lib.NM_LOCAL_ROWS, gl, back)
    assert torch.equal(back, u_loc)
    ctx.close()


def test_asynchronous_readback_overlaps_the_next_assembly():
    """nm_get_csr_compact(NM_HOST_ASYNC): the matrix of step k is complete after nm_sync_downloads even though the next
    assembly on the same context was started in between."""
    bg, im = make_config("sphere", 3, band_only=True)
    bg2, im2 = make_config("sphere", 2, band_only=True)
    ctx = coupling.make_context(bg, im, boxes=True)
    ctx.intersect(3, 1e-15)
    want = [t.clone() for t in ctx.csr_compact()]
    nnz = ctx.nnz
    bufs = (torch.empty(want[0].numel() + 8, dtype=torch.int32).pin_memory(), torch.empty(want[0].numel() + 8, dtype=torch.int32).pin_memory(),
            torch.empty(nnz + 8, dtype=torch.int32).pin_memory(), torch.empty(nnz + 8, dtype=torch.float64).pin_memory())
    got = ctx.csr_compact(out=bufs, asynchronous=True)
    coupling.set_grids(ctx, bg2, im2, boxes=True)         # the next problem on the same context, before the copies are waited for
    ctx.intersect(3, 1e-15)
    ctx.build_sparsity()
    ctx.sync_downloads()
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    with pytest.raises(ValueError):
        ctx.csr_compact(out=tuple(t.clone() for t in want), asynchronous=True)      # pageable buffers are refused
    ctx.close()


def test_merge_path_flag_is_reset_after_a_fallback():
    """the `problem` flag of the single-pass merge was never cleared, so one overflow sent every later
    build of the context through both merges.  Force one fallback, then expect the fast path again."""
    bg, _ = make_config("disk", 4, band_only=False)
    ctx = coupling.make_context(bg, circle_grid(8), boxes=True)      # 8 long segments: > 48 distinct dofs per immersed cell
    ctx.intersect(3, 1e-15)
    ctx.build_sparsity()
    assert ctx.info("merge_path") == 1
    bg4, im4 = make_config("disk", 4, band_only=False)
    coupling.set_grids(ctx, bg4, im4, boxes=True)
    ctx.intersect(3, 1e-15)
    ctx.build_sparsity()
    assert ctx.info("merge_path") == 0
    fresh = coupling.make_context(bg4, im4, boxes=True)               # a brand-new context starts from cleared flags
    fresh.intersect(3, 1e-15)
    fresh.build_sparsity()
    assert fresh.info("merge_path") == 0
    assert torch.equal(fresh.csr()[2], ctx.csr()[2])
    ctx.close(); fresh.close()


def test_from_quadrature_accumulates_repeated_pairs_and_checks_ranges():
    """the same (space cell, immersed cell) pair given twice must add up like the reference's
    distribute_local_to_global does per tuple (coupling_utilities.h:285-287); out-of-range dof tables are refused."""
    bg, im = make_config("disk", 1, band_only=False)
    info = coupling.collect_quadratures_on_overlapped_grids(bg, im, 3, 1e-15)
    ctx = info.ctx
    _, _, val_once = ctx.csr()
    # every tuple split in two halves of its points, given as two tuples of the same pair
    off = info.q_offset
    n = len(info)
    im2 = torch.cat([info.immersed_cell, info.immersed_cell]); bg2 = torch.cat([info.space_cell, info.space_cell])
    first = (off[1:] - off[:-1]) // 2
    cnt = torch.cat([first, (off[1:] - off[:-1]) - first])
    off2 = torch.zeros(2 * n + 1, dtype=torch.int64); off2[1:] = torch.cumsum(cnt, 0)
    idx = torch.cat([torch.cat([torch.arange(int(off[i]), int(off[i] + first[i])) for i in range(n)]),
                     torch.cat([torch.arange(int(off[i] + first[i]), int(off[i + 1])) for i in range(n)])])
    ctx.assemble_from_quadrature(im2.contiguous(), bg2.contiguous(), off2, info.points[idx].contiguous(), info.weights[idx].contiguous())
    _, _, val_twice = ctx.csr()
    assert (val_twice - val_once).norm() <= 1e-13 * val_once.norm()
    bad = bg.dofs.clone(); bad[int(info.space_cell[0]), 0] = bg.n_dofs + 5     # a cell the merge reads (it holds an accepted pair)
    ctx.set_background_dofs(bad.to(torch.int32).contiguous(), bg.n_dofs)
    with pytest.raises(_lib.NmError) as e:      # checked before the merge dereferences the table
        ctx.build_sparsity()
    assert e.value.code == 1 and "DoF table" in str(e.value)
    ctx.close()


def test_fp64_peak_is_measured():
    ctx = _lib.Context(0)
    p = ctx.measure_fp64_peak(5.0)
    assert 5e12 < p < 1e14, p          # B200: ~37-40 TFLOP/s FP64
    ctx.close()


def test_fused_ct_vmult_on_two_gpus():
    """tests/multi_gpu_check.py under torchrun: fused Ct.vmult (replicated and owned) against SpMV + NCCL.  Needs 2 GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    port = 29600 + os.getpid() % 300
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py"), "4"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "rel err" in r.stdout


@pytest.mark.parametrize("name,cycle", [("sphere", 2), ("sphere", 4), ("flower", 7), ("disk", 3)])
def test_assembly_is_bitwise_reproducible(name, cycle):
    """The merge adds the contributions of one matrix entry in a fixed order (match.any groups, ascending lanes): both
    orientations and both products are bit-identical from run to run, hanging-node rows included."""
    bg, im = make_config(name, cycle, band_only=cycle > 3, device="cuda")
    assert cycle > 3 or bg.c_dof.numel() > 0      # the full backgrounds carry hanging-node lines
    runs = []
    for rep in range(3):
        ctx = coupling.make_context(bg, im, on_device=True, boxes=True)
        ctx.intersect(3, 1e-15)
        _, col, val = ctx.csr(device="cuda")
        _, colt, valt = ctx.csr(transpose=True, device="cuda")
        z = ctx.spmv(torch.ones(bg.n_dofs, dtype=torch.float64, device="cuda"), transpose=True)
        y = ctx.spmv(torch.ones(im.n_dofs, dtype=torch.float64, device="cuda"), transpose=False)
        assert ctx.info("merge_path") == 0
        runs.append([t.clone() for t in (col, val, colt, valt, y, z)])
        ctx.close()
    for r in runs[1:]:
        for a, b in zip(runs[0], r):
            assert torch.equal(a, b)


def test_small_merge_table_overflow_falls_back_to_the_large_one(c_oracle):
    """The single-pass merge tries 128 hash slots per immersed cell first (96 distinct dofs) when the cells average few
    accepted pairs.  One large quad among many small ones (7 x 7 background cells: 128 distinct dofs) overflows it: the
    256-slot table must redo the merge -- not the general two-pass path -- and the matrix must be the oracle's; the
    context then remembers to start with the large table."""
    from fe_helpers import uniform_background
    from non_matching_test_suite_b200.grids import BackgroundGrid, ImmersedGrid
    n = 16
    lo, hi = uniform_background(3, n)
    h = 2.0 / n
    nv = n + 1
    ij = np.round((lo + 1.0) / h).astype(np.int64)
    dofs = np.zeros((lo.shape[0], 8), dtype=np.int64)
    for v in range(8):
        dofs[:, v] = sum((ij[:, k] + ((v >> k) & 1)) * nv ** k for k in range(3))
    e32, e64 = torch.empty(0, dtype=torch.int32), torch.empty(0, dtype=torch.float64)
    bg = BackgroundGrid(3, torch.from_numpy(lo), torch.from_numpy(hi), torch.from_numpy(dofs.astype(np.int32)), nv ** 3,
                        torch.zeros(lo.shape[0], dtype=torch.int32), e32, torch.zeros(1, dtype=torch.int64), e32, e64)
    quads = []
    z = -1.0 + 5.37 * h                                                          # inside one layer of cells
    a, b = -1.0 + 2.21 * h, -1.0 + 8.83 * h                                      # spans 7 x 7 cells
    quads.append([[a, a, z], [b, a, z], [a, b, z], [b, b, z + 0.01 * h]])
    rng = np.random.default_rng(3)
    for _ in range(300):                                                        # small quads: one or two pairs each
        c = -0.9 + 1.8 * rng.random(3)
        s = 0.3 * h
        quads.append([c, c + [s, 0, 0.1 * s], c + [0, s, 0], c + [s, s, 0.05 * s]])
    verts = torch.tensor(np.array(quads, dtype=np.float64))
    im = ImmersedGrid(2, 3, verts.contiguous(), torch.arange(len(quads), dtype=torch.int32)[:, None], len(quads))
    ref = c_oracle.assemble(bg, im)
    ctx = coupling.make_context(bg, im, boxes=True)
    c = ctx.intersect(3, 1e-15)
    assert c["n_accepted"] <= 12 * im.n_cells
    rp, col, val = ctx.csr()
    assert ctx.info("merge_path") == 0 and ctx.info("merge_slots") == 256
    assert np.array_equal(col.numpy().astype(np.uint32), ref["col"]) and np.array_equal(rp.numpy().astype(np.uint64), ref["rowptr"])
    assert np.linalg.norm(val.numpy() - ref["val"]) <= 1e-12 * np.linalg.norm(ref["val"])
    # the usual grids stay on the small table
    bg2, im2 = make_config("sphere", 3)
    ctx2 = coupling.make_context(bg2, im2, boxes=True)
    ctx2.intersect(3, 1e-15)
    ctx2.build_sparsity()
    assert ctx2.info("merge_path") == 0 and ctx2.info("merge_slots") == 128
    ctx.close(); ctx2.close()


@pytest.mark.parametrize("name,cycle,compact", [("sphere", 2, False), ("sphere", 3, True), ("disk", 3, False)])
def test_both_products_in_one_call(monkeypatch, name, cycle, compact):
    """nm_spmv_local_pair = the two nm_spmv_local calls, bit for bit, from pinned host vectors (overlapped transfers) and
    from device vectors; also right after an asynchronous matrix read-back, which shares its download stream."""
    if compact:
        monkeypatch.setenv("NMGPU_COMPACT_ROWS", "1")
    bg, im = make_config(name, cycle)
    ctx = coupling.make_context(bg, im, boxes=True)
    ctx.intersect(3, 1e-15)
    ctx.build_sparsity()
    nr = ctx.n_rows_local()
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(im.n_cells, dtype=torch.float64, generator=g) - 0.5).pin_memory()
    u = (torch.rand(nr, dtype=torch.float64, generator=g) - 0.5).pin_memory()
    y_ref, z_ref = ctx.spmv_local(x, transpose=False), ctx.spmv_local(u, transpose=True)
    bufs = (torch.empty(nr, dtype=torch.int32).pin_memory(), torch.empty(nr, dtype=torch.int32).pin_memory(),
            torch.empty(ctx.nnz, dtype=torch.int32).pin_memory(), torch.empty(ctx.nnz, dtype=torch.float64).pin_memory())
    ctx.csr_compact(out=bufs, asynchronous=True)
    y, z = ctx.spmv_local_pair(x, u, out_ct=torch.empty(nr, dtype=torch.float64).pin_memory(), out_c=torch.empty(im.n_cells, dtype=torch.float64).pin_memory())
    assert torch.equal(y, y_ref) and torch.equal(z, z_ref)
    yd, zd = ctx.spmv_local_pair(x.cuda(), u.cuda())
    assert torch.equal(yd.cpu(), y_ref) and torch.equal(zd.cpu(), z_ref)
    ctx.sync_downloads()
    ctx.close()


@pytest.mark.parametrize("name,cycle", [("sphere", 3), ("flower", 2)])
def test_inputs_read_in_place_give_the_same_matrix_and_stay_untouched(name, cycle):
    """NM_DEVICE_INPLACE: DoF tables and immersed vertices are used where the caller keeps them.  Same matrix bit for bit, the
    arrays are never written, and the context goes back to its own copies when the inputs are set again the usual way."""
    bg, im = make_config(name, cycle)
    ref = coupling.make_context(bg, im, boxes=True)
    ref.intersect(3, 1e-15)
    want = ref.csr()
    dev = "cuda:0"
    lo, hi, dofs = bg.lo.to(dev), bg.hi.to(dev), bg.dofs.to(torch.int32).to(dev).contiguous()
    cd, cp, cm, cw = bg.c_dof.to(torch.int32).to(dev), bg.c_ptr.to(dev), bg.c_master.to(torch.int32).to(dev), bg.c_weight.to(dev)
    verts, idofs = im.verts.to(dev).contiguous(), im.dofs.to(torch.int32).to(dev).contiguous()
    snap = [t.clone() for t in (dofs, verts, idofs)]
    ctx = _lib.Context(0)
    for _ in range(2):                                     # twice: the second call replaces a borrowed array by a borrowed one
        ctx.set_background(bg.spacedim, lo=lo, hi=hi, dofs=dofs, n_dofs=bg.n_dofs, c_dof=cd, c_ptr=cp, c_master=cm, c_weight=cw, inplace=True)
        ctx.set_immersed(im.dim, im.spacedim, verts, idofs, im.n_dofs, inplace=True)
        ctx.intersect(3, 1e-15)
        got = ctx.csr()
        assert all(torch.equal(a, b) for a, b in zip(got, want))
    held_inplace = ctx.device_bytes()
    assert all(torch.equal(a, b) for a, b in zip((dofs, verts, idofs), snap))
    coupling.set_grids(ctx, bg, im, boxes=True)            # host arrays again: own copies, the caller's tensors may go away
    del dofs, verts, idofs
    torch.cuda.empty_cache()
    ctx.intersect(3, 1e-15)
    assert all(torch.equal(a, b) for a, b in zip(ctx.csr(), want))
    assert ctx.device_bytes() > held_inplace
    ctx.close(); ref.close()


@pytest.mark.parametrize("name,cycle", [("sphere", 3), ("disk", 5)])
def test_assemble_host_with_cubic_cells_as_corner_and_edge(name, cycle):
    """nm_host_problem.edge: cubic cells sent as lower corner + edge length give the matrix of the lo / hi form bit for bit
    (the upper corner is fl(lo + edge), exact for the dyadic coordinates of a hyper_cube refinement)."""
    bg, im = make_config(name, cycle)
    d = bg.spacedim
    edge = (bg.hi[:, 0] - bg.lo[:, 0]).contiguous()
    assert bool((bg.lo + edge[:, None] == bg.hi).all())
    args = (bg.dofs.to(torch.int32).contiguous(), bg.n_dofs, bg.c_dof.to(torch.int32).contiguous(), bg.c_ptr.contiguous(),
            bg.c_master.to(torch.int32).contiguous(), bg.c_weight.contiguous(), im.dim, im.verts.contiguous(), im.dofs.to(torch.int32).contiguous(),
            im.n_dofs, 3, 1e-15)
    a, b = _lib.Context(0), _lib.Context(0)
    ca = a.assemble_host(d, bg.lo.contiguous(), bg.hi.contiguous(), *args)
    cb = b.assemble_host(d, bg.lo.contiguous(), None, *args, edge=edge)
    assert ca == cb
    assert all(torch.equal(x, y) for x, y in zip(a.csr(), b.csr()))
    a.close(); b.close()


@pytest.mark.parametrize("name,cycle", [("sphere", 3), ("disk", 5)])
def test_a_dof_repeated_inside_one_cell_goes_to_the_general_merge(c_oracle, name, cycle):
    """Without constraint lines the single-pass merge adds the contributions of one pair side by side, which needs the
    dofs of a cell to be distinct.  A table that repeats one (no finite element does; distribute_local_to_global would
    simply add both, coupling_utilities.h:285-287) is detected when the rows are read and the general merge, which sums
    any multiset, builds the matrix: same result as the oracle's sequential scatter."""
    import dataclasses
    bg, im = make_config(name, cycle, band_only=True)
    assert bg.c_dof.numel() == 0                                          # the kernel without the line lookup
    plain = coupling.make_context(bg, im, boxes=True)
    plain.intersect(3, 1e-15)
    plain.build_sparsity()
    assert plain.info("merge_path") == 0
    cell = int(plain.pairs()[1][0])                                       # a background cell that holds an accepted pair
    plain.close()
    dofs = bg.dofs.clone()
    dofs[cell, 1] = dofs[cell, 0]
    bad = dataclasses.replace(bg, dofs=dofs)
    ref, ref_plain = c_oracle.assemble(bad, im), c_oracle.assemble(bg, im)
    ctx = coupling.make_context(bad, im, boxes=True)
    counts = ctx.intersect(3, 1e-15)
    rp, col, val = ctx.csr()
    assert ctx.info("merge_path") == 1
    assert np.array_equal(col.numpy().astype(np.uint32), ref["col"]) and np.array_equal(rp.numpy().astype(np.uint64), ref["rowptr"])
    # sub-simplices under the reference's |J| < 1e-12 threshold depend on the subdivision (see test_gpu_parity.py)
    allowed = 1e-12 + (counts["n_dropped"] + ref["dropped"]) * 1e-12 / np.linalg.norm(ref["val"])
    assert np.linalg.norm(val.numpy() - ref["val"]) <= allowed * np.linalg.norm(ref["val"])
    # the entries of the repeated dof carry both vertex functions of that cell: the row of the dof that was overwritten
    # lost this cell's share, the row of the repeated one gained it
    A = sp.csr_matrix((val.numpy(), col.numpy(), rp.numpy()), shape=(bg.n_dofs, im.n_dofs))
    B = sp.csr_matrix((ref_plain["val"], ref_plain["col"], ref_plain["rowptr"]), shape=(bg.n_dofs, im.n_dofs))
    kept, gone = int(bg.dofs[cell, 0]), int(bg.dofs[cell, 1])
    moved = (B[gone] - A[gone]).toarray().ravel()
    assert moved.max() > 0 and np.abs((A[kept] - B[kept]).toarray().ravel() - moved).max() <= 1e-12 * np.abs(B[kept].data).max() + allowed
    ctx.close()


def test_local_products_never_read_untouched_scratch_entries():
    """The products in local numbering run on scratch vectors of which only the addressed entries are written.  The
    padding slots of the SpMV kernels (0 x entry) used to read entry 0 of them: with recycled memory that held NaN bit
    patterns the whole result turned NaN.  Device memory is filled with NaNs first, dof 0 (a corner of the background,
    far from the interface) is untouched, and the local products must equal the global ones on the touched entries."""
    nan_fill = torch.full((1 << 27,), float("nan"), dtype=torch.float64, device="cuda")     # 1 GiB of NaNs, handed back
    del nan_fill
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    bg, im = make_config("sphere", 2, band_only=False)
    ctx = coupling.make_context(bg, im, boxes=True)
    ctx.intersect(3, 1e-15)
    rows = ctx.csr_compact()[0].long()
    assert int(rows[0]) != 0                                            # dof 0 has no entry
    g = torch.Generator().manual_seed(4)
    x = torch.rand(im.n_cells, dtype=torch.float64, generator=g)
    u = torch.rand(bg.n_dofs, dtype=torch.float64, generator=g)
    xg = torch.zeros(im.n_dofs, dtype=torch.float64); xg[im.dofs[:, 0].long()] = x
    ug = torch.zeros(bg.n_dofs, dtype=torch.float64); ug[rows] = u[rows]
    y_loc = ctx.spmv_local(x, transpose=False)
    z_loc = ctx.spmv_local(u[rows].contiguous(), transpose=True)
    assert bool(torch.isfinite(y_loc).all()) and bool(torch.isfinite(z_loc).all())
    assert torch.equal(y_loc, ctx.spmv(xg, transpose=False)[rows])
    assert torch.equal(z_loc, ctx.spmv(ug, transpose=True)[im.dofs[:, 0].long()])
    y2, z2 = ctx.spmv_local_pair(x.pin_memory(), u[rows].contiguous().pin_memory())
    assert torch.equal(y2, y_loc) and torch.equal(z2, z_loc)
    ctx.close()
