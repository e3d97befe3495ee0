#!/usr/bin/env python
"""Benchmark of the hot path: coupling-matrix assembly by exact intersections + C / C^T SpMV.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one full pass over the workload: hand both grids to the library, candidate pairs,
clipping + quadrature + local coupling vectors, merge through the constraint lines into CSR
(both orientations), one Ct.vmult and one C.Tvmult.  With the immersed grid sharded over several
GPUs the Ct.vmult result is distributed by background-dof ownership like the reference's Trilinos
vector (`--ct-vmult owned`, the default: one fused kernel adds every row result into its owner's
vector over NVLink) or replicated on every rank (`--ct-vmult replicated`).

`value`  : accepted intersection pairs per second, whole job, inputs already resident in HBM.
`e2e`    : the same through the public API with HOST (pinned) buffers: host->device copies of both
           grids and of the product vectors, device->host read-back of the CSR matrix and of both
           product vectors, all inside the timed region.
Workload : refined 2D-in-3D sphere (data/lm_2d3d.smooth_sphere.prm geometry), true refinement cycles:
           cycle 8 on one GPU (BASELINE.json configs[3]), cycle 9 on two and four, cycle 10 on eight
           (configs[4], 1.0e9 pairs); synthetic band-only background, deterministic geometry.
The N = 1 line also carries short runs of the 2D configurations (`configs`: disk / flower at cycle 14,
configs[0..2]).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "coupling-matrix intersection pairs/s (+ C SpMV GB/s in `spmv`)"
UNIT = "pairs/s"
DEFAULT_CYCLE = {"sphere": 8, "disk": 14, "flower": 14}
NAMES = {"sphere": "lm_2d3d.smooth_sphere.prm geometry (2D sphere in 3D cube, quad-vs-hex clipping)",
         "disk": "lm_1d2d_disk.smooth.prm geometry (1D circle in 2D square)",
         "flower": "lm_1d2d_flower.smooth.prm geometry (flower polygon in 2D square)"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="sphere", choices=["sphere", "disk", "flower"])
    ap.add_argument("--cycle", type=int, default=None, help="refinement cycle of the single-GPU workload (default: 8 for sphere)")
    ap.add_argument("--full-background", action="store_true", help="whole background grid instead of the band near the interface (small cycles)")
    ap.add_argument("--ref-cells", type=int, default=None, help="immersed cells of the CPU sample per step (reference arm / cpu_baseline)")
    ap.add_argument("--ref-cycle", type=int, default=None, help="take the CPU sample from this cycle instead of the workload's (testing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sub-configs", action="store_true", help="skip the short 2D runs reported under `configs` (N = 1)")
    ap.add_argument("--no-dropin", action="store_true", help="skip the timing of the reference's three calls through the C++ drop-in header (N = 1)")
    ap.add_argument("--dropin-cycle", type=int, default=6, help="sphere cycle of the drop-in flow timing")
    ap.add_argument("--no-check", action="store_true", help="skip the multi-GPU correctness checks (N > 1)")
    ap.add_argument("--check-cycle", type=int, default=3, help="cycle of the shard-vs-single-GPU matrix check (N > 1)")
    ap.add_argument("--e2e-serial", action="store_true", help="end-to-end leg with a blocking matrix read-back (no overlap with the next step)")
    ap.add_argument("--e2e-global-vectors", action="store_true",
                    help="end-to-end leg with global-length product vectors and row pointers (the round-1 form)")
    ap.add_argument("--e2e-boxes", action="store_true", help="end-to-end leg: send lo/hi corners even when the cells are cubes")
    ap.add_argument("--ct-vmult", choices=["replicated", "owned", "owned-pull"], default="owned",
                    help="result of Ct.vmult on N > 1 GPUs: distributed by contiguous background-dof ranges like the reference's Trilinos "
                         "vector (reduce-scatter semantics, the default) or replicated on every rank (all-reduce semantics)")
    ap.add_argument("--nccl", action="store_true", help="reduce Ct.vmult with NCCL (reduce_scatter / all_reduce) instead of the fused NVLink push")
    ap.add_argument("--verts", action="store_true", help="send all 2^d vertices of every background cell instead of lo/hi boxes")
    ap.add_argument("--shard-of", type=int, default=None,
                    help="single-GPU run of rank 0's share of a W-GPU job (its own dof numbering, no process group): sizes the memory of one rank")
    ap.add_argument("--only-sub-config", default=None, help="run one 2D sub-configuration as the main line (profiling): disk, flower, disk_full")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def pin_to_gpu_numa_node(local):
    """Run this process (and therefore its pinned host allocations, first touch) on the NUMA node the GPU hangs off:
    with one rank per GPU every rank streams several GB per step over its own PCIe link, and a buffer on the remote
    socket halves that.  Best effort; returns a short description for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/" % (dom, bus, dev)
        node = int(open(path + "numa_node").read().strip())
        cpus = open(path + "local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if node >= 0 and ids:
            os.sched_setaffinity(0, ids)
            return "numa node %d (%d cpus)" % (node, len(ids))
        return "numa node %d: affinity unchanged" % node
    except Exception as e:  # no sysfs entry, no permission: keep the default placement
        return "unchanged (%s)" % type(e).__name__


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU restatement (oracle/nm_oracle.c) on the box's host cores
# --------------------------------------------------------------------------------------------
def cpu_sample(config, cycle, n_cells, steps, warmup, threads=None):
    """Times the CPU restatement on a bounded SAMPLE of the workload: the first `n_cells` immersed cells of the true
    grid of refinement `cycle` with the background band near them (same mesh sizes as the GPU arm)."""
    import torch
    from non_matching_test_suite_b200.grids import sample_problem
    from oracle import c_oracle as co
    co.build()
    co.lib().nmo_set_num_threads(threads or os.cpu_count() or 1)   # torchrun pins OMP_NUM_THREADS=1: set it explicitly
    cores = co.lib().nmo_num_threads()
    bg, im = sample_problem(config, cycle, n_cells)
    times, acc = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        r = co.assemble(bg, im, n1d=3, tol=1e-15)
        x = torch.ones(im.n_dofs, dtype=torch.float64).numpy()
        co.spmv(0, r["rowptr"], r["col"], r["val"], x, bg.n_dofs, im.n_dofs)
        u = torch.ones(bg.n_dofs, dtype=torch.float64).numpy()
        co.spmv(1, r["rowptr"], r["col"], r["val"], u, bg.n_dofs, im.n_dofs)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        acc = int(r["accepted"].shape[0])
    ms = 1e3 * sum(times) / len(times)
    sample = "%s cycle %d, first %d of %d immersed cells with their background band (%d cells): %d accepted pairs per step" % (
        config, cycle, im.n_cells, im.n_dofs, bg.n_cells, acc)
    return dict(value=acc / (ms * 1e-3), unit=UNIT, cores=cores, kind="port", sample=sample), ms, acc


def workload_cycle(args, world):
    from non_matching_test_suite_b200.grids import weak_scaling_sizes
    cycle = args.cycle if args.cycle is not None else DEFAULT_CYCLE[args.config]
    return weak_scaling_sizes(args.config, cycle, world)[0]


def ref_cells(args, threads):
    if args.ref_cells is not None:
        return args.ref_cells
    if args.config == "sphere":
        return 98304 if threads >= 8 else 24576          # ~1.0e6 / 2.5e5 pairs: a couple of seconds per step
    return 1 << 18


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cyc = args.ref_cycle if args.ref_cycle is not None else workload_cycle(args, args.gpus)
    n_thr = os.cpu_count() or 1
    base, ms, acc = cpu_sample(args.config, cyc, ref_cells(args, n_thr), max(1, args.steps), max(0, min(args.warmup, 1)))
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.config, workload_cycle(args, args.gpus), args.gpus, args.full_background),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference algorithm (oracle/nm_oracle.c, FP64, OpenMP over all host threads); the real "
                "deal.II/CGAL reference cannot be built in this image (SURVEY.md 8(c)). Each step is a bounded sample of the SAME "
                "workload (same refinement cycle and mesh sizes; `cpu_baseline.sample` says which cells).",
    }
    print(json.dumps(line), flush=True)


def workload_config(config, cycle, n_gpus, full_background=False):
    """Identical for the native and the reference arm of one (config, N)."""
    return {"workload": "%s, synthetic refinement cycle %d, %s" % (
                NAMES[config], cycle, "full background with hanging-node and Dirichlet constraint lines" if full_background else
                "band-only background (the finest-level cells near the interface: as in the full grid at this refinement, no "
                "constraint line touches a cut cell; lines are exercised by the parity tests and the full-background sub-config)"),
            "sharding": "immersed-cell ranges, %d shard(s)" % n_gpus, "degree": 3, "tol": 1e-15,
            "l2_policy": "inputs and intermediates are larger than L2 (no flush needed)"}


# --------------------------------------------------------------------------------------------
# native arm
# --------------------------------------------------------------------------------------------
class Problem:
    """One rank's grids on the device + the synthetic product vectors."""

    def __init__(self, torch, config, cycle, rank, world, dev, full_background=False, verts=False, shard_of=None):
        from non_matching_test_suite_b200.grids import make_config, shard_problem
        t0 = time.time()
        if full_background:
            bg, im = make_config(config, cycle, band_only=False, device=dev)
            info = {"global_background_cells": bg.n_cells, "global_immersed_cells": im.n_cells, "background_dofs": bg.n_dofs,
                    "rank0_background_cells": bg.n_cells, "rank0_immersed_cells": im.n_cells}
        elif shard_of:
            bg, im, info = shard_problem(config, cycle, 0, shard_of, device=dev, weak=True, want_coords=False, standalone=True)
        else:
            bg, im, info = shard_problem(config, cycle, rank, world, device=dev, weak=True, want_coords=False)
        torch.cuda.synchronize()
        self.bg, self.im, self.info, self.d = bg, im, info, bg.spacedim
        u32 = lambda t: t.to(torch.int32).contiguous()
        geom = dict(verts=bg.vertices().contiguous()) if verts else dict(lo=bg.lo.contiguous(), hi=bg.hi.contiguous())
        self.dev_in = dict(**geom, dofs=u32(bg.dofs), c_dof=u32(bg.c_dof), c_ptr=bg.c_ptr.contiguous(), c_master=u32(bg.c_master),
                           c_weight=bg.c_weight.contiguous(), im_verts=im.verts.contiguous(), im_dofs=u32(im.dofs).reshape(-1))
        # drop the generator's own copies: the library needs the memory
        bg.lo = bg.hi = bg.dofs = bg.level = bg.c_dof = bg.c_ptr = bg.c_master = bg.c_weight = bg.dof_coords = None
        im.verts = im.dofs = None
        g0, g1 = torch.Generator(device=dev).manual_seed(0), torch.Generator(device=dev).manual_seed(1)
        self.x = torch.rand(im.n_dofs, dtype=torch.float64, device=dev, generator=g0) * 2 - 1     # lambda: immersed dofs (global)
        self.u = torch.rand(bg.n_dofs, dtype=torch.float64, device=dev, generator=g1) * 2 - 1     # background dofs (global)
        self.y = torch.empty(bg.n_dofs, dtype=torch.float64, device=dev)
        self.z = torch.empty(im.n_dofs, dtype=torch.float64, device=dev)
        self.n_im = int(self.dev_in["im_verts"].shape[0])
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        self.gen_seconds = time.time() - t0


def run_native(args):
    import torch
    import torch.distributed as dist
    from non_matching_test_suite_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # NCCL writes its version banner / debug lines to stdout; rank 0 must print exactly one JSON line there, so NCCL's
    # own output is sent to stderr.  NCCL honours NCCL_DEBUG_FILE only above the VERSION level, hence VERSION -> WARN
    # (which still prints the banner, now to stderr).
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pin_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")

    if args.only_sub_config:
        name = args.only_sub_config
        cfgname, cyc, full = {"disk": ("disk", 14, False), "flower": ("flower", 14, False), "disk_full": ("disk", 6, True)}[name]
        args.config, args.cycle, args.full_background = cfgname, cyc, full

    cycle = workload_cycle(args, args.shard_of or world)
    P = Problem(torch, args.config, args.cycle if args.cycle is not None else DEFAULT_CYCLE[args.config], rank, world, dev,
                full_background=args.full_background, verts=args.verts, shard_of=args.shard_of)
    ctx = _lib.Context(local)
    fp64_peak = ctx.measure_fp64_peak(20.0)
    res = run_problem(args, torch, dist, _lib, ctx, P, rank, world, local, dev, peak, peak_src, fp64_peak,
                      steps=args.steps, warmup=args.warmup, e2e=not args.no_e2e)

    # ---- multi-GPU correctness, once, outside every timed region -----------------------------------
    checks = None
    if world > 1 and not args.no_check:
        checks = multi_gpu_checks(args, torch, dist, _lib, ctx, P, res, rank, world, local, dev)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.shard_of:
        n_thr = os.cpu_count() or 1
        cpu_base, _, _ = cpu_sample(args.config, cycle, 4 * ref_cells(args, n_thr) if args.config == "sphere" else ref_cells(args, n_thr), 1, 0)
        # the reference as shipped runs 1 MPI rank x 1 thread (code/scripts/run_all_tests.sh): same port, one thread, smaller sample
        one, _, _ = cpu_sample(args.config, cycle, ref_cells(args, 1), 1, 0, threads=1)
        cpu_base["single_thread"] = {"value": one["value"], "unit": UNIT, "cores": 1, "sample": one["sample"]}

    device_bytes = ctx.device_bytes()
    ctx.close()
    del ctx, P
    torch.cuda.empty_cache()

    # ---- the 2D configurations (BASELINE.json configs[0..2]), short runs on one GPU -----------------
    sub = None
    if world == 1 and not args.no_sub_configs and not args.shard_of and not args.only_sub_config and args.config == "sphere":
        sub = {}
        for name, (cfgname, cyc, full) in {"disk_c14": ("disk", 14, False), "flower_c14": ("flower", 14, False),
                                            "disk_full_background_c6": ("disk", 6, True)}.items():
            a2 = argparse.Namespace(**vars(args))
            a2.config, a2.full_background, a2.e2e_global_vectors, a2.e2e_serial = cfgname, full, False, False
            P2 = Problem(torch, cfgname, cyc, 0, 1, dev, full_background=full)
            c2 = _lib.Context(local)
            r2 = run_problem(a2, torch, dist, _lib, c2, P2, 0, 1, local, dev, peak, peak_src, fp64_peak,
                             steps=max(3, min(args.steps, 10)), warmup=3, e2e=not args.no_e2e)
            c2.close()
            sub[name] = {"config": workload_config(cfgname, cyc, 1, full), "value": r2["value"], "unit": UNIT, "ms_per_step": r2["ms"],
                         "pairs": r2["tot_acc"], "candidates": r2["tot_cand"], "nnz": r2["tot_nnz"], "phases_ms": r2["phases"],
                         "roofline": r2["roofline"], "spmv": r2["spmv"], "e2e": r2["e2e"], "gpu_launches": r2["launches"],
                         "host_syncs_per_step": r2["syncs"], "background_cells": P2.info.get("rank0_background_cells"),
                         "immersed_cells": P2.info.get("rank0_immersed_cells")}
            del P2, c2
            torch.cuda.empty_cache()

    dropin = None
    if world == 1 and rank == 0 and not args.no_dropin and not args.shard_of and not args.only_sub_config and args.config == "sphere":
        dropin = dropin_flow(torch, "sphere", args.dropin_cycle)

    failed = bool(checks and not checks.get("ok", True))
    if rank == 0:
        line = {
            "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.config, cycle, args.shard_of or world, args.full_background),
            "workload_info": res["info"],
            "pairs": res["tot_acc"], "candidates": res["tot_cand"], "nnz": res["tot_nnz"], "nnz_per_s": res["tot_nnz"] / (res["ms"] * 1e-3),
            "pairs_per_min": res["value"] * 60.0,
            "counts_rank0": res["counts"], "phases_ms_rank0": res["phases"], "roofline": res["roofline"], "spmv": res["spmv"],
            "clocks": res["clocks"], "e2e": res["e2e"], "gpu_launches": int(res["launches"]), "host_syncs_per_step": res["syncs"],
            "cpu_baseline": cpu_base, "device_bytes_rank0": device_bytes, "gen_seconds": res["gen_seconds"],
            "ct_vmult_reduction": res["ct_vmult_label"], "index": res["index"], "host_numa": numa,
        }
        if checks is not None:
            line["ct_vmult_check"] = checks["ct_vmult_check"]
            line["shard_check"] = checks["shard_check"]
        if sub is not None:
            line["configs"] = sub
        if dropin is not None:
            line["e2e_dropin"] = dropin
        if args.shard_of:
            line["note"] = "rank 0's share of a %d-GPU job run stand-alone on one GPU (memory sizing); not a scaling number" % args.shard_of
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if failed:
        sys.exit(1)


def run_problem(args, torch, dist, _lib, ctx, P, rank, world, local, dev, peak, peak_src, fp64_peak, steps, warmup, e2e):
    from non_matching_test_suite_b200.distributed import FusedCtVmult, from_context, owned_dof_range
    bg, im, d = P.bg, P.im, P.d
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    sharded_op = from_context(ctx)
    sharded_op.m, sharded_op.n = bg.n_dofs, im.n_dofs     # the context learns its sizes with the first assembly
    fused = {"op": None, "tried": False, "why": ""}
    owned = args.ct_vmult in ("owned", "owned-pull")
    pull = args.ct_vmult == "owned-pull"

    def assemble_device(inp):
        if "verts" in inp:   # all 2^d vertices of every background cell, as mapping.get_vertices() returns them
            ctx.set_background(d, verts=inp["verts"], dofs=inp["dofs"], n_dofs=bg.n_dofs, c_dof=inp["c_dof"], c_ptr=inp["c_ptr"],
                               c_master=inp["c_master"], c_weight=inp["c_weight"], inplace=True)
        else:                # lo/hi corners: what include/coupling_utilities.h sends for Cartesian background cells
            ctx.set_background(d, lo=inp["lo"], hi=inp["hi"], dofs=inp["dofs"], n_dofs=bg.n_dofs, c_dof=inp["c_dof"], c_ptr=inp["c_ptr"],
                               c_master=inp["c_master"], c_weight=inp["c_weight"], inplace=True)
        # NM_DEVICE_INPLACE: the inputs are resident in HBM; the DoF tables and the immersed vertices are read where they are
        ctx.set_immersed(im.dim, im.spacedim, inp["im_verts"], inp["im_dofs"], im.n_dofs, inplace=True)
        counts = ctx.intersect(3, 1e-15)
        ctx.build_sparsity()
        ctx.assemble()
        return counts

    def get_fused():
        if world > 1 and not fused["tried"] and not args.nccl:
            fused["tried"] = True
            op = FusedCtVmult(ctx)               # symmetric-memory result vector, peer / multicast mapped
            ok = torch.tensor([1 if op.available else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            fused["op"] = op if int(ok.item()) == 1 else None
            fused["why"] = "" if op.available else op.reason
        return fused["op"]

    def ct_vmult_device(x):
        """Ct.vmult on device vectors (global numbering); returns the result this rank keeps."""
        if world == 1:
            return ctx.spmv(x, transpose=False, out=P.y)
        op = get_fused()
        if owned:
            if op is None:
                return sharded_op.vmult_owned(x)
            return op.vmult_owned_pull(x) if pull else op.vmult_owned(x)
        if op is not None:
            return op.vmult(x)
        ctx.spmv(x, transpose=False, out=P.y)
        dist.all_reduce(P.y)
        return P.y

    def step_device():
        counts = assemble_device(P.dev_in)
        ct_vmult_device(P.x)
        ctx.spmv(P.u, transpose=True, out=P.z)       # C.Tvmult : this shard's immersed entries
        return counts

    def timed(fn, n_steps, n_warm, finish=None):
        for _ in range(n_warm):
            counts = fn()
        if finish:
            finish()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        phases = {}
        l0, s0 = ctx.launch_count(), ctx.info("host_syncs")
        e0.record(stream)
        for _ in range(n_steps):
            counts = fn()
            for k, v in ctx.timings().items():
                phases[k] = phases.get(k, 0.0) + v / n_steps
        if finish:
            finish()
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n_steps
        launches = (ctx.launch_count() - l0) // n_steps
        syncs = (ctx.info("host_syncs") - s0) // n_steps
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), counts, phases, launches, syncs

    sampler = ClockSampler(local)
    sampler.start()
    ms, counts, phases, launches, syncs = timed(step_device, steps, warmup)
    clocks = sampler.stop()

    n_acc, n_cand = counts["n_accepted"], counts["n_candidates"]
    nnz = ctx.nnz
    n_rows = ctx.n_rows_local()
    tot = torch.tensor([n_acc, n_cand, nnz], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    tot_acc, tot_cand, tot_nnz = (int(v) for v in tot.tolist())
    value = tot_acc / (ms * 1e-3)

    # ---- roofline of the dominant kernel (this rank; CUDA events on the library's stream) ---------
    per_cand = 296 if d == 3 else 104      # pair ids 8 B + background cell 2^d*d*8 B + immersed cell 2^(d-1)*d*8 B
    per_acc_vals = 64 if d == 3 else 32    # 2^d local values written per accepted pair
    clip_name = "clip_moments3_kernel" if d == 3 else "clip_local_kernel<2>"
    kernels = {
        clip_name: (phases["k_clip"], n_cand * per_cand + n_acc * per_acc_vals),
        "cand_probe+cand_place": (phases["k_cand_count"] + phases["k_cand_fill"], 0),
        "merge_rows_fast": (phases["k_merge"], n_acc * ((36 if d == 3 else 20) + per_acc_vals)),
    }
    dom = max(kernels, key=lambda k: kernels[k][0])
    dom_ms, dom_bytes = kernels[clip_name]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    roofline = {"kernel": clip_name, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                "kernel_ms": dom_ms, "largest_kernel_by_time": dom, "kernel_ms_all": {k: v[0] for k, v in kernels.items()},
                "whole_step_algorithmic_GBs": (dom_bytes + n_acc * (36 if d == 3 else 20)) / (phases_sum(phases) * 1e-3) / 1e9,
                "measured_limiter": None}
    tj = {}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    key = "clip3d" if d == 3 else "clip2d"
    if key in tj:
        t = tj[key]
        roofline["traffic"] = t.get("dram_bytes_per_candidate", 0) * n_cand or None
        roofline["traffic_source"] = t.get("source")
        roofline["measured_limiter"] = t.get("limiter")
        fl = t.get("fp64_flops_per_candidate")
        if fl and dom_ms > 0:
            flops = fl * n_cand
            roofline["fp64"] = {"flops": flops, "achieved": flops / (dom_ms * 1e-3) / 1e12, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                                "frac": flops / (dom_ms * 1e-3) / fp64_peak, "peak_source": "nm_measure_fp64_peak (DFMA chains, this run)",
                                "flops_source": t.get("flops_source")}
    if "fp64" not in roofline:
        roofline["fp64"] = {"flops": None, "peak": fp64_peak / 1e12, "unit": "TFLOP/s", "frac": None,
                            "peak_source": "nm_measure_fp64_peak (DFMA chains, this run)"}
    # SpMV bytes: entries + the rows the kernel walks (all dofs in full-row mode, the touched ones in compact-row mode: never
    # the global dof count of a sharded run) + the vector entries this rank's columns address
    stored_rows = ctx.info("stored_rows")
    a_bytes = 12 * nnz + 16 * stored_rows + 8 * P.n_im
    c_bytes = 12 * nnz + 16 * P.n_im + 8 * n_rows
    spmv = {"Ct_vmult_GBs": a_bytes / (phases["spmv"] * 1e-3) / 1e9 if phases["spmv"] > 0 else None,
            "C_Tvmult_GBs": c_bytes / (phases["spmv_t"] * 1e-3) / 1e9 if phases["spmv_t"] > 0 else None,
            "nnz": nnz, "stored_rows": stored_rows, "stored_rows_with_entries": n_rows, "immersed_cells": P.n_im,
            "bytes_formula": "Ct.vmult: 12*nnz + 16*stored_rows + 8*immersed_cells; C.Tvmult: 12*nnz + 16*immersed_cells + 8*rows_with_entries "
                             "(stored_rows = rows the kernel walks: every dof of this rank's background in full-row mode, touched dofs in compact-row mode)",
            "note": "per rank, kernel only (the cross-GPU reduction is inside the fused Ct.vmult kernel when it is used)"}
    spmv["Ct_vmult_frac_hbm"] = spmv["Ct_vmult_GBs"] / peak if spmv["Ct_vmult_GBs"] else None
    spmv["C_Tvmult_frac_hbm"] = spmv["C_Tvmult_GBs"] / peak if spmv["C_Tvmult_GBs"] else None

    # ---- end to end: host buffers in, matrix + vectors out ----------------------------------------
    e2e_res = None
    if e2e:
        e2e_res = run_e2e(args, torch, dist, _lib, ctx, P, rank, world, dev, timed, get_fused, sharded_op, owned, tot_acc, max(1, steps))

    return dict(value=value, ms=ms, counts=counts, phases=phases, launches=launches, syncs=syncs, clocks=clocks, roofline=roofline, spmv=spmv,
                e2e=e2e_res, tot_acc=tot_acc, tot_cand=tot_cand, tot_nnz=tot_nnz, info=P.info, gen_seconds=P.gen_seconds,
                ct_vmult_label=ct_vmult_label(world, args, fused), index={1: "per-cell chains", 2: "brick bitmasks"}.get(ctx.info("index_kind"), "?"),
                fused=fused, sharded_op=sharded_op)


def phases_sum(ph):
    return sum(ph[k] for k in ("upload", "index", "candidates", "split", "clip", "merge", "transpose", "spmv", "spmv_t"))


def run_e2e(args, torch, dist, _lib, ctx, P, rank, world, dev, timed, get_fused, sharded_op, owned, tot_acc, e_steps):
    from non_matching_test_suite_b200.distributed import owned_dof_range
    bg, im, d = P.bg, P.im, P.d
    pull = args.ct_vmult == "owned-pull"
    host_in = {k: v.cpu().pin_memory() for k, v in P.dev_in.items()}
    # Cubic cells whose upper corner is fl(lo + edge) bit for bit (the reference's hyper_cube refinements: dyadic coordinates)
    # travel as lower corner + one edge length (nm_host_problem.edge): checked here, once, outside the timed region
    cubes = False
    if "lo" in host_in and not args.e2e_boxes:
        edge = (host_in["hi"][:, 0] - host_in["lo"][:, 0]).contiguous()
        if bool((host_in["lo"] + edge[:, None] == host_in["hi"]).all()):
            host_in["edge"] = edge.pin_memory()
            del host_in["hi"]
            cubes = True
    h2d_in = sum(v.numel() * v.element_size() for v in host_in.values())
    nnz_cap = int(ctx.nnz * 1.02) + 1024
    rows_cap = int(ctx.n_rows_local() * 1.02) + 1024
    pin = lambda n, dt: torch.empty(n, dtype=dt).pin_memory()
    glob = args.e2e_global_vectors
    pipelined = not args.e2e_serial and not glob and "lo" in host_in
    # pinned host memory this rank is about to hold: inputs (already pinned above) + read-back buffers (two sets when the
    # read-back is pipelined) + vectors; all ranks of the node share the host RAM
    out_bytes = rows_cap * 8 + nnz_cap * 12
    vec_bytes = 8 * (2 * P.n_im + 2 * rows_cap + bg.n_dofs // max(1, world))
    try:
        avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
    except Exception:
        avail = None
    if avail is not None:
        share = 0.8 * avail / max(1, world)       # inputs are already resident: `avail` is what is left
        if pipelined and 2 * out_bytes + vec_bytes > share:
            pipelined = False                     # one buffer set, blocking read-back
        if out_bytes + vec_bytes > share:
            del host_in
            return {"skipped": "host memory: %.1f GB per rank needed for the read-back buffers, %.1f GB available per rank" % (
                (out_bytes + vec_bytes) / 1e9, share / 1e9)}
    io = {"h2d": 0, "d2h": 0}

    if glob:
        xh, uh = P.x.cpu().pin_memory(), P.u.cpu().pin_memory()
        yh, zh = pin(bg.n_dofs, torch.float64), pin(im.n_dofs, torch.float64)
        csr_buf = (pin(bg.n_dofs + 1, torch.int64), pin(nnz_cap, torch.int32), pin(nnz_cap, torch.float64))
    else:
        xh = (torch.rand(P.n_im, dtype=torch.float64, generator=torch.Generator().manual_seed(2)) * 2 - 1).pin_memory()    # local: one per cell
        uh = (torch.rand(rows_cap, dtype=torch.float64, generator=torch.Generator().manual_seed(3)) * 2 - 1).pin_memory()  # local: one per touched row
        lo_o, hi_o = owned_dof_range(bg.n_dofs, rank, world)
        yh = pin(rows_cap if world == 1 else max(1, hi_o - lo_o), torch.float64)
        zh = pin(P.n_im, torch.float64)
        xg = torch.empty(im.n_dofs, dtype=torch.float64, device=dev) if world > 1 else None
        n_buf = 2 if pipelined else 1
        csr_bufs = [(pin(rows_cap, torch.int32), pin(rows_cap, torch.int32), pin(nnz_cap, torch.int32), pin(nnz_cap, torch.float64))
                    for _ in range(n_buf)]
    state = {"k": 0, "out": None}

    def assemble_host():
        if "lo" in host_in:
            return ctx.assemble_host(d, host_in["lo"], host_in.get("hi"), host_in["dofs"], bg.n_dofs, host_in["c_dof"], host_in["c_ptr"],
                                     host_in["c_master"], host_in["c_weight"], im.dim, host_in["im_verts"], host_in["im_dofs"], im.n_dofs, 3, 1e-15,
                                     edge=host_in.get("edge"))
        ctx.set_background(d, verts=host_in["verts"], dofs=host_in["dofs"], n_dofs=bg.n_dofs, c_dof=host_in["c_dof"], c_ptr=host_in["c_ptr"],
                           c_master=host_in["c_master"], c_weight=host_in["c_weight"])
        ctx.set_immersed(im.dim, im.spacedim, host_in["im_verts"], host_in["im_dofs"], im.n_dofs)
        counts = ctx.intersect(3, 1e-15)
        ctx.build_sparsity()
        ctx.assemble()
        return counts

    def step_global():
        counts = assemble_host()
        ctx.spmv(xh, transpose=False, out=yh)
        if world > 1:
            yy = P.y.copy_(yh, non_blocking=True)
            dist.all_reduce(yy)
            yh.copy_(yy)
        ctx.spmv(uh, transpose=True, out=zh)
        state["out"] = ctx.csr(transpose=False, values=True, out=csr_buf)
        io["h2d"] = h2d_in + xh.numel() * 8 + uh.numel() * 8
        io["d2h"] = sum(t.numel() * t.element_size() for t in state["out"]) + yh.numel() * 8 + zh.numel() * 8
        return counts

    trace = os.environ.get("NMGPU_BENCH_TRACE") == "1" and rank == 0
    tr = []

    def mark(label, t0):
        if trace:
            tr.append("%s %.1f" % (label, (time.perf_counter() - t0) * 1e3))
        return time.perf_counter()

    def step_local():
        t0 = time.perf_counter()
        counts = assemble_host()
        t0 = mark("assemble_host", t0)
        nr = ctx.n_rows_local()
        t0 = mark("n_rows", t0)
        if world == 1:
            # Ct.vmult (touched rows only) and C.Tvmult in one call: the upload of u overlaps the download of Ct*x
            ctx.spmv_local_pair(xh, uh[:nr], out_ct=yh[:nr], out_c=zh)
            y_bytes = nr * 8
        else:
            ctx.local_to_global(_lib.NM_LOCAL_IMMERSED, xh, xg)                   # this shard's lambda entries
            op = get_fused()
            if owned:
                yy = (op.vmult_owned_pull(xg) if pull else op.vmult_owned(xg)) if op is not None else sharded_op.vmult_owned(xg)
                yh[:yy.numel()].copy_(yy)                                         # the slice of the dofs this rank owns
                y_bytes = yy.numel() * 8
            else:
                yy = op.vmult(xg) if op is not None else sharded_op.vmult(xg)
                ctx.global_to_local(_lib.NM_LOCAL_ROWS, yy, yh[:nr])              # replicated result: read the touched rows
                y_bytes = nr * 8
        t0 = mark("ct_vmult", t0)
        if world > 1:
            ctx.spmv_local(uh[:nr], transpose=True, out=zh)                       # C.Tvmult, this shard's cells
        t0 = mark("c_tvmult", t0)
        buf = csr_bufs[state["k"] % len(csr_bufs)]
        state["k"] += 1
        if pipelined:
            ctx.sync_downloads()     # the previous step's matrix: it left the device under this step's uploads and kernels
            t0 = mark("sync_downloads", t0)
            state["out"] = ctx.csr_compact(out=buf, asynchronous=True)
        else:
            state["out"] = ctx.csr_compact(out=buf)
        t0 = mark("readback_call", t0)
        if trace:
            sys.stderr.write("e2e step: " + " | ".join(tr) + "\n")
            del tr[:]
        io["h2d"] = h2d_in + xh.numel() * 8 + nr * 8
        io["d2h"] = sum(t.numel() * t.element_size() for t in state["out"]) + y_bytes + zh.numel() * 8
        return counts

    fn = step_global if glob else step_local
    ms_e, counts_e, _, _, _ = timed(fn, e_steps, 1, finish=ctx.sync_downloads if pipelined else None)
    tot = torch.tensor([io["h2d"], io["d2h"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    api = ("nm_assemble_host (host %s + dofs + constraint lines + immersed grid in one call, uploads overlapped with the phases that do "
           "not need them yet)" % ("lower corners + edge lengths of the cubic cells" if cubes else "lo/hi") if "lo" in host_in else "nm_set_background (host vertices) + nm_set_immersed + nm_intersect + nm_build_sparsity + nm_assemble")
    if glob:
        api += " + nm_spmv x2 (global-length host vectors) + nm_get_csr (global row pointers)"
    else:
        api += (" + nm_spmv_local_pair (both products, host vectors in local numbering: one entry per immersed cell / per touched row; the two "
                    "directions of the link used at once)" if world == 1 else
                " + nm_local_to_global + fused Ct.vmult (%s) + nm_spmv_local (C.Tvmult)" % args.ct_vmult) + \
               " + nm_get_csr_compact (rows with entries only" + (", asynchronous: the read-back of step k runs under the uploads and kernels of step k+1; "
                                                                 "the last one is waited for inside the timed region)" if pipelined else ")")
    res = {"value": tot_acc / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(tot[0].item()), "d2h_bytes_per_step": int(tot[1].item()),
           "bytes_are": "whole job (sum over ranks)", "ms_per_step": ms_e, "steps": e_steps, "pipelined_readback": pipelined,
           "api": "_lib.Context (C ABI): " + api}
    del host_in
    return res


def dropin_flow(torch, config, cycle):
    """The reference's own three calls through the C++ drop-in header (tests/cpp/bench_dropin.cpp on the deal.II stub):
    host objects in, tuples / pattern / matrix out, next to the direct C-ABI path on the same arrays."""
    import numpy as np
    import tempfile
    from non_matching_test_suite_b200.grids import make_config
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "bench_dropin")
    lib = os.path.join(ROOT, "non_matching_test_suite_b200")
    try:
        if not os.path.exists(exe):
            os.makedirs(os.path.dirname(exe), exist_ok=True)
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "tests", "cpp", "stub"),
                                   os.path.join(ROOT, "tests", "cpp", "bench_dropin.cpp"), "-o", exe, "-L" + lib, "-lnmgpu", "-Wl,-rpath," + lib])
        bg, im = make_config(config, cycle, band_only=True, device="cuda")
        with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
            path = f.name
            np.array([bg.spacedim, bg.n_cells, bg.n_dofs, im.n_cells, im.n_dofs, bg.c_dof.numel(), bg.c_master.numel()], dtype=np.int64).tofile(f)
            for t, dt in ((bg.lo, np.float64), (bg.hi, np.float64), (bg.dofs, np.uint32), (bg.c_dof, np.uint32), (bg.c_ptr, np.uint64),
                          (bg.c_master, np.uint32), (bg.c_weight, np.float64), (im.verts, np.float64), (im.dofs, np.uint32)):
                t.cpu().numpy().astype(dt).tofile(f)
        del bg, im
        torch.cuda.empty_cache()
        r = subprocess.run([exe, path], capture_output=True, text=True, timeout=900)
        os.unlink(path)
        if r.returncode != 0:
            return {"error": "bench_dropin exited %d: %s" % (r.returncode, r.stderr[-300:])}
        d = json.loads(r.stdout.strip().splitlines()[-1])
        d["workload"] = "%s cycle %d (band-only)" % (config, cycle)
        d["value"] = d["pairs"] / (d["three_calls_ms"] * 1e-3)
        d["unit"] = UNIT
        d["ratio_three_calls_to_direct_plus_quadrature_readback"] = d["three_calls_ms"] / (d["direct_assemble_host_plus_readback_ms"] + d["quadrature_readback_ms"])
        d["note"] = ("host side is the deal.II STUB (tests/cpp/stub): cell loops, Quadrature objects and the pattern / matrix sinks are plain "
                     "std::vector code, pageable memory; the three calls use the device-resident quadrature of the collect call (fingerprint of the "
                     "returned vector) -- only the tuples the signature demands travel back")
        return d
    except Exception as e:
        return {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}


def multi_gpu_checks(args, torch, dist, _lib, ctx, P, res, rank, world, local, dev):
    """(a) the fused Ct.vmult kernels against nm_spmv + NCCL on the benchmark matrix; (b) on a small cycle, the union of the
    ranks' shards against the matrix one GPU assembles from the whole grids.  Outside every timed region."""
    import numpy as np
    from non_matching_test_suite_b200.distributed import FusedCtVmult, owned_dof_range
    from non_matching_test_suite_b200.grids import make_config, shard_problem
    from non_matching_test_suite_b200 import coupling
    bg = P.bg
    sharded_op, fused = res["sharded_op"], res["fused"]
    out = {"ok": True}
    # (a)
    lam = P.x
    ctx.spmv(lam, transpose=False, out=P.y)
    y_ref = P.y.clone()
    dist.all_reduce(y_ref)                                   # nm_spmv + NCCL all-reduce: the reference result
    scale = float(y_ref.abs().max())
    lo, hi = owned_dof_range(bg.n_dofs, rank, world)
    errs = {}
    op = fused["op"]
    if op is None and not args.nccl:
        op_try = FusedCtVmult(ctx)
        ok = torch.tensor([1 if op_try.available else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        op = op_try if int(ok.item()) == 1 else None
    if op is not None:
        y_own = op.vmult_owned(lam)
        errs["fused_owned"] = float((y_own - y_ref[lo:hi]).abs().max()) / scale if hi > lo else 0.0
        y_pull = op.vmult_owned_pull(lam)
        errs["fused_owned_pull"] = float((y_pull - y_ref[lo:hi]).abs().max()) / scale if hi > lo else 0.0
        y_pull0 = y_pull.clone()
        errs["fused_owned_pull_not_bitwise_reproducible"] = 0.0 if torch.equal(op.vmult_owned_pull(lam), y_pull0) else 1.0
        y_rep = op.vmult(lam)
        errs["fused_replicated_%s" % ("multicast" if op.multicast else "peer")] = float((y_rep - y_ref).abs().max()) / scale
    y_rs = sharded_op.vmult_owned(lam)
    errs["nccl_reduce_scatter"] = float((y_rs - y_ref[lo:hi]).abs().max()) / scale if hi > lo else 0.0
    names = sorted(errs)
    t = torch.tensor([errs[k] for k in names], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    errs = {k: float(v) for k, v in zip(names, t.tolist())}
    mx = max(errs.values()) if errs else 0.0
    out["ct_vmult_check"] = {"max_rel_err": mx, "per_variant": errs, "against": "nm_spmv + NCCL all_reduce on the benchmark matrix",
                             "tolerance": 1e-12, "fused_available": op is not None}
    out["ok"] &= mx <= 1e-12
    # (b)
    cyc = args.check_cycle
    bgs, ims, _ = shard_problem(args.config, cyc, rank, world, device=dev, weak=False)
    c2 = coupling.make_context(bgs, ims, device=local, boxes=True, on_device=True)
    c2.intersect(3, 1e-15)
    ids, ln, col, val = c2.csr_compact(device="cuda")
    rows = torch.repeat_interleave(ids.long(), ln.long())
    trip = torch.stack([rows.double(), col.double(), val]).cpu().numpy()
    n_acc = c2.counts["n_accepted"]
    c2.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, (trip, n_acc))
    sc = None
    if rank == 0:
        bgf, imf = make_config(args.config, cyc, band_only=True, device=dev)
        c3 = coupling.make_context(bgf, imf, device=local, boxes=True, on_device=True)
        c3.intersect(3, 1e-15)
        ids, ln, col, val = c3.csr_compact(device="cuda")
        rows = torch.repeat_interleave(ids.long(), ln.long()).cpu().numpy()
        full = (rows, col.cpu().numpy().astype(np.int64), val.cpu().numpy())
        n_full = c3.counts["n_accepted"]
        c3.close()
        allt = np.concatenate([g[0] for g in gathered], axis=1)
        r, c, v = allt[0].astype(np.int64), allt[1].astype(np.int64), allt[2]
        order = np.lexsort((c, r))
        r, c, v = r[order], c[order], v[order]
        pattern = r.shape == full[0].shape and bool(np.array_equal(r, full[0]) and np.array_equal(c, full[1]))
        rel = float(np.linalg.norm(v - full[2]) / np.linalg.norm(full[2])) if pattern else None
        sc = {"cycle": cyc, "pairs_sum_of_shards": int(sum(g[1] for g in gathered)), "pairs_single_gpu": int(n_full),
              "pattern_identical": pattern, "rel_frobenius_err": rel, "bitwise_identical_values": bool(pattern and np.array_equal(v, full[2])),
              "tolerance": 1e-12}
        ok = pattern and rel is not None and rel <= 1e-12 and sc["pairs_sum_of_shards"] == sc["pairs_single_gpu"]
    else:
        ok = True
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["shard_check"] = sc
    out["ok"] &= bool(int(flag.item()))
    return out


def ct_vmult_label(world, args, fused):
    """How the partial Ct.vmult results of the ranks were combined in the timed step."""
    if world == 1:
        return "none (1 GPU)"
    why = " (fused path unavailable: %s)" % fused["why"] if fused["why"] else ""
    if args.ct_vmult == "owned-pull" and fused["op"] is not None:
        return ("result distributed by dof ownership, bit-reproducible: nm_spmv_rows + nm_pull_rows (every owner reads its segment of each "
                "rank's send buffer over NVLink and adds in rank order; %s)" % fused["op"].sync_kind)
    if args.ct_vmult in ("owned", "owned-pull"):
        if fused["op"] is not None:
            return "result distributed by dof ownership: nm_spmv_push_owned over NVLink (red.add into the owner's peer-mapped vector; %s)" % fused["op"].sync_kind
        return "result distributed by dof ownership: nccl reduce_scatter" + why
    if fused["op"] is not None:
        return "replicated result: nm_spmv_push over NVLink, %s; %s" % ("multimem.red on the multicast address" if fused["op"].multicast
                                                                        else "red.add into peer-mapped vectors", fused["op"].sync_kind)
    return "replicated result: nccl all_reduce" + why


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
