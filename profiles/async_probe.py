"""Where does the host time of an asynchronous read-back go?  (nm_get_csr_compact with NM_HOST_ASYNC must return at once.)"""
import sys, time, torch
sys.path.insert(0, '/root/repo')
from non_matching_test_suite_b200 import coupling, _lib
from non_matching_test_suite_b200.grids import make_config
import ctypes as C
bg, im = make_config("sphere", 7, band_only=True, device="cuda")
ctx = coupling.make_context(bg, im, on_device=True, boxes=True)
ctx.intersect(3, 1e-15); nnz = ctx.build_sparsity(); nr = ctx.n_rows_local()
pin = lambda n, dt: torch.empty(n, dtype=dt).pin_memory()
bufs = (pin(nr + 8, torch.int32), pin(nr + 8, torch.int32), pin(nnz + 8, torch.int32), pin(nnz + 8, torch.float64))
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); n = C.c_uint64(0)
    ctx._chk(ctx.L.nm_get_csr_compact(ctx.h, C.byref(n), None, None, None, None, 0)); t1 = time.perf_counter()
    ids, ln, col, val = bufs[0][:nr], bufs[1][:nr], bufs[2][:nnz], bufs[3][:nnz]
    pinned = all(t.is_pinned() for t in (ids, ln, col, val)); t2 = time.perf_counter()
    ctx._chk(ctx.L.nm_get_csr_compact(ctx.h, C.byref(n), _lib._ptr(ids)[0], _lib._ptr(ln)[0], _lib._ptr(col)[0], _lib._ptr(val)[0], 2)); t3 = time.perf_counter()
    ctx.sync_downloads(); t4 = time.perf_counter()
    print("nnz %d rows %d pinned %s | size query %.2f ms | slicing+is_pinned %.2f ms | async call %.2f ms | sync_downloads %.2f ms" % (
        nnz, nr, pinned, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3), flush=True)
# the same copy issued by torch on a side stream
s = torch.cuda.Stream()
d = torch.empty(nnz, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
with torch.cuda.stream(s):
    bufs[3][:nnz].copy_(d, non_blocking=True)
t1 = time.perf_counter(); s.synchronize(); t2 = time.perf_counter()
print("torch non_blocking D2H of %d MB: call %.2f ms, wait %.2f ms" % (nnz * 8 >> 20, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
