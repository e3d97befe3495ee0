"""PCIe host<->device bandwidth of this box: each direction alone and both at once (pinned memory, two streams).
Explains the end-to-end numbers of bench.py: a step uploads ~5 GB and reads back ~3.5 GB."""
import json, sys, torch
n = 1 << 30
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    s1.synchronize(); s2.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
import time
def wall(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
r = {"bytes_each": n, "h2d_ms": wall(h2d), "d2h_ms": wall(d2h), "both_ms": wall(both)}
r["h2d_GBs"] = n / r["h2d_ms"] / 1e6; r["d2h_GBs"] = n / r["d2h_ms"] / 1e6; r["both_sum_GBs"] = 2 * n / r["both_ms"] / 1e6
print(json.dumps(r))
