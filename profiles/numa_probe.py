"""Per-rank host->device and device->host bandwidth with all ranks copying at once, under different memory policies of the
pinned buffer (default first touch, or bound to NUMA node k with set_mempolicy).  Tells whether the multi-GPU end-to-end
numbers of bench.py are limited by where the pinned host memory lives.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 profiles/numa_probe.py"""
import ctypes, glob, json, os, sys, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
libc = ctypes.CDLL(None, use_errno=True)
SYS_set_mempolicy, MPOL_DEFAULT, MPOL_BIND = 238, 0, 2          # x86-64


def set_policy(node):
    if node is None:
        return libc.syscall(SYS_set_mempolicy, MPOL_DEFAULT, None, 0)
    mask = ctypes.c_ulong(1 << node)
    return libc.syscall(SYS_set_mempolicy, MPOL_BIND, ctypes.byref(mask), 64)


nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
info = {"rank": rank, "nodes": nodes, "cpus_allowed": len(os.sched_getaffinity(0))}
try:
    bus = torch.cuda.get_device_properties(local).pci_bus_id
    info["gpu_numa"] = open("/sys/bus/pci/devices/0000:%02x:00.0/numa_node" % bus).read().strip()
except Exception as e:
    info["gpu_numa"] = "?"
for k in nodes:
    try:
        info["node%d_cpus" % k] = open("/sys/devices/system/node/node%d/cpulist" % k).read().strip()
    except Exception:
        pass
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
res = {}
for pol in [None] + nodes:
    rc = set_policy(pol)
    if rc != 0:
        res[str(pol)] = "set_mempolicy failed errno %d" % ctypes.get_errno()
        continue
    try:
        h = torch.empty(n, dtype=torch.uint8).pin_memory()
    except Exception as e:
        res[str(pol)] = "alloc failed: %s" % str(e)[:80]
        set_policy(None)
        continue
    set_policy(None)
    out = {}
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            fn()
        torch.cuda.synchronize()
        out[name] = round(4 * n / (time.perf_counter() - t0) / 1e9, 1)
        if world > 1:
            dist.barrier()
    res["default" if pol is None else "node%d" % pol] = out
    del h
info["GBs_all_ranks_at_once"] = res
for r in range(world):
    if r == rank:
        print(json.dumps(info), flush=True)
    if world > 1:
        dist.barrier()
if world > 1:
    dist.destroy_process_group()
