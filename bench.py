#!/usr/bin/env python
"""Benchmark of the hot path: coupling-matrix assembly by exact intersections + C / C^T SpMV.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one full pass over the workload: hand both grids to the library, candidate pairs,
clipping + quadrature + local coupling vectors, merge through the constraint lines into CSR
(both orientations), one Ct.vmult and one C.Tvmult (+ all-reduce of the partial Ct.vmult result
when the immersed grid is sharded over several GPUs).

`value`  : accepted intersection pairs per second, whole job, inputs already resident in HBM.
`e2e`    : the same through the public API with HOST (pinned) buffers, host->device copies of
           both grids and device->host read-back of the CSR matrix and both product vectors
           inside the timed region.
Workload : refined 2D-in-3D sphere (data/lm_2d3d.smooth_sphere.prm geometry, BASELINE.json
           configs[3]); synthetic band-only background, deterministic geometry, no RNG in the mesh.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="sphere", choices=["sphere", "disk", "flower"])
    ap.add_argument("--cycle", type=int, default=None, help="refinement cycle of the workload (default: 8 for sphere)")
    ap.add_argument("--ref-cycle", type=int, default=None, help="refinement cycle of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-two-in-flight", action="store_true",
                    help="additionally report the end-to-end throughput with two independent steps in flight (two contexts, two host "
                         "threads): PCIe is full duplex, so the uploads of one step run under the read-back of the other")
    ap.add_argument("--e2e-separate-calls", action="store_true",
                    help="end-to-end leg through nm_set_background_boxes + nm_set_immersed + nm_intersect + ... instead of nm_assemble_host")
    ap.add_argument("--ct-vmult", choices=["replicated", "owned"], default="replicated",
                    help="result of Ct.vmult on N > 1 GPUs: replicated on every rank (all-reduce semantics, the default) or "
                         "distributed by contiguous background-dof ranges like the reference's Trilinos vector (reduce-scatter semantics)")
    ap.add_argument("--nccl-allreduce", action="store_true", help="reduce Ct.vmult with NCCL all_reduce instead of the fused NVLink push")
    ap.add_argument("--verts", action="store_true", help="send all 2^d vertices of every background cell instead of lo/hi boxes")
    return ap.parse_args()


DEFAULT_CYCLE = {"sphere": 8, "disk": 14, "flower": 14}


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


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU restatement (oracle/nm_oracle.c) on the box's host cores
# --------------------------------------------------------------------------------------------
def cpu_sample(config, cycle, steps, warmup, threads=None):
    import torch
    from non_matching_test_suite_b200.grids import make_config
    from oracle import c_oracle as co
    co.build()
    co.lib().nmo_set_num_threads(threads or os.cpu_count() or 1)   # torchrun pins OMP_NUM_THREADS=1: set it explicitly
    cores = co.lib().nmo_num_threads()
    bg, im = make_config(config, cycle, band_only=True)
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
    sample = "%s cycle %d: %d background cells (band only), %d immersed cells, %d accepted pairs per step" % (
        config, cycle, bg.n_cells, im.n_cells, acc)
    return dict(value=acc / (ms * 1e-3), unit=UNIT, cores=cores, kind="port", sample=sample), ms, acc


def pick_ref_cycle(args):
    if args.ref_cycle is not None:
        return args.ref_cycle
    n = os.cpu_count() or 8
    if args.config == "sphere":
        return 6 if n >= 24 else 5
    return 10


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cyc = pick_ref_cycle(args)
    base, ms, acc = cpu_sample(args.config, cyc, max(1, args.steps), max(0, min(args.warmup, 1)))
    workload_cycle = args.cycle if args.cycle is not None else DEFAULT_CYCLE[args.config]
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.config, workload_cycle, args.gpus, None),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference algorithm (oracle/nm_oracle.c, FP64, OpenMP over all host threads); the real "
                "deal.II/CGAL reference cannot be built in this image (SURVEY.md 8(c)). Each step is a bounded sample of the workload.",
    }
    print(json.dumps(line), flush=True)


def workload_config(config, cycle, n_gpus, extra):
    names = {"sphere": "lm_2d3d.smooth_sphere.prm geometry (2D sphere in 3D cube, quad-vs-hex clipping)",
             "disk": "lm_1d2d_disk.smooth.prm geometry", "flower": "lm_1d2d_flower.smooth.prm geometry"}
    c = {"workload": "%s, synthetic refinement cycle %d, band-only background with hanging-node constraints" % (names[config], cycle),
         "sharding": "immersed-cell ranges, %d shard(s)" % n_gpus, "degree": 3, "tol": 1e-15,
         "l2_policy": "inputs and intermediates are larger than L2 (no flush needed)"}
    if extra:
        c.update(extra)
    return c


# --------------------------------------------------------------------------------------------
# native arm
# --------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist
    from non_matching_test_suite_b200 import _lib, coupling
    from non_matching_test_suite_b200.grids import make_config, shard_problem

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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cycle = args.cycle if args.cycle is not None else DEFAULT_CYCLE[args.config]

    # ---- workload: generated on the GPU, this rank's shard kept -----------------------------------
    t0 = time.time()
    bg, im, info = shard_problem(args.config, cycle, rank, world, device=dev, weak=True, want_coords=False)
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    d = bg.spacedim
    u32 = lambda t: t.to(torch.int32).contiguous()
    # background cells as lo/hi corners -- the format include/coupling_utilities.h hands to nm_set_background_boxes
    # (pass --verts to feed all 2^d vertices per cell through nm_set_background instead)
    geom = dict(verts=bg.vertices().contiguous()) if args.verts else dict(lo=bg.lo.contiguous(), hi=bg.hi.contiguous())
    dev_in = dict(**geom, dofs=u32(bg.dofs), c_dof=u32(bg.c_dof), c_ptr=bg.c_ptr.contiguous(),
                  c_master=u32(bg.c_master), c_weight=bg.c_weight.contiguous(), im_verts=im.verts.contiguous(), im_dofs=u32(im.dofs))
    x_dev = (torch.rand(im.n_dofs, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(0)) * 2 - 1)
    u_dev = (torch.rand(bg.n_dofs, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 2 - 1)
    y_dev = torch.empty(bg.n_dofs, dtype=torch.float64, device=dev)
    z_dev = torch.empty(im.n_dofs, dtype=torch.float64, device=dev)

    ctx = _lib.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    from non_matching_test_suite_b200.distributed import from_context, owned_dof_range
    sharded_op = from_context(ctx)

    csr_out = {}
    fused = {"op": None, "tried": False, "why": ""}

    def step(inp, x, u, y, z, host):
        if host and "lo" in inp and not args.e2e_separate_calls:
            # one call from host arrays: uploads queued up front, each phase waits only for what it reads
            counts = ctx.assemble_host(d, inp["lo"], inp["hi"], inp["dofs"], bg.n_dofs, inp["c_dof"], inp["c_ptr"], inp["c_master"],
                                       inp["c_weight"], im.dim, inp["im_verts"], inp["im_dofs"], im.n_dofs, 3, 1e-15)
            return finish_step(counts, x, u, y, z, host)
        if "verts" in inp:   # all 2^d vertices of every background cell, as mapping.get_vertices() returns them
            ctx.set_background(d, verts=inp["verts"], dofs=inp["dofs"], n_dofs=bg.n_dofs, c_dof=inp["c_dof"], c_ptr=inp["c_ptr"],
                               c_master=inp["c_master"], c_weight=inp["c_weight"])
        else:                # lo/hi corners: what include/coupling_utilities.h sends for Cartesian background cells
            ctx.set_background(d, lo=inp["lo"], hi=inp["hi"], dofs=inp["dofs"], n_dofs=bg.n_dofs, c_dof=inp["c_dof"], c_ptr=inp["c_ptr"],
                               c_master=inp["c_master"], c_weight=inp["c_weight"])
        ctx.set_immersed(im.dim, im.spacedim, inp["im_verts"], inp["im_dofs"], im.n_dofs)
        counts = ctx.intersect(3, 1e-15)
        ctx.build_sparsity()
        ctx.assemble()
        return finish_step(counts, x, u, y, z, host)

    def finish_step(counts, x, u, y, z, host):
        if world > 1 and fused["op"] is None and not args.nccl_allreduce and fused["tried"] is False:
            from non_matching_test_suite_b200.distributed import FusedCtVmult
            fused["tried"] = True
            op = FusedCtVmult(ctx)               # symmetric-memory result vector, peer / multicast mapped
            ok = torch.tensor([1 if op.available else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            fused["op"] = op if int(ok.item()) == 1 else None
            fused["why"] = "" if op.available else op.reason
        if world > 1 and fused["op"] is not None and args.ct_vmult == "owned":
            # reduce-scatter semantics: every row result crosses NVLink once, to the rank that owns the dof
            xx = x if not host else x_dev.copy_(x, non_blocking=True)
            yy = fused["op"].vmult_owned(xx)
            if host:
                lo_o, hi_o = owned_dof_range(bg.n_dofs, rank, world)
                y[lo_o:hi_o].copy_(yy)
        elif world > 1 and args.ct_vmult == "owned":
            xx = x if not host else x_dev.copy_(x, non_blocking=True)
            yy = sharded_op.vmult_owned(xx)      # SpMV + NCCL reduce_scatter
            if host:
                lo_o, hi_o = owned_dof_range(bg.n_dofs, rank, world)
                y[lo_o:hi_o].copy_(yy)
        elif world > 1 and fused["op"] is not None:
            # Ct.vmult fused with its reduction: one kernel adds this rank's rows into every rank's result over NVLink
            xx = x if not host else x_dev.copy_(x, non_blocking=True)
            yy = fused["op"].vmult(xx)
            if host:
                y.copy_(yy)
        else:
            ctx.spmv(x, transpose=False, out=y)  # Ct.vmult : partial background vector
            if world > 1:
                yy = y if not host else y_dev.copy_(y, non_blocking=True)
                dist.all_reduce(yy)              # the only collective of the path
                if host:
                    y.copy_(yy)
        ctx.spmv(u, transpose=True, out=z)       # C.Tvmult : this shard's immersed entries
        out = None
        if host:
            if "buf" not in csr_out or csr_out["buf"][1].numel() < ctx.nnz:   # pinned, reused read-back buffers
                cap = int(ctx.nnz * 1.05) + 16
                csr_out["buf"] = (torch.empty(bg.n_dofs + 1, dtype=torch.int64).pin_memory(),
                                  torch.empty(cap, dtype=torch.int32).pin_memory(), torch.empty(cap, dtype=torch.float64).pin_memory())
            out = ctx.csr(transpose=False, values=True, out=csr_out["buf"])   # D2H of the assembled matrix
        return counts, out

    def timed(inp, x, u, y, z, host, steps, warmup):
        for _ in range(warmup):
            counts, out = step(inp, x, u, y, z, host)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        phases = {}
        l0 = ctx.launch_count()
        e0.record(stream)
        for _ in range(steps):
            counts, out = step(inp, x, u, y, z, host)
            for k, v in ctx.timings().items():
                phases[k] = phases.get(k, 0.0) + v / steps
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        launches = (ctx.launch_count() - l0) // steps
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), counts, phases, launches, out

    sampler = ClockSampler(local)
    sampler.start()
    ms, counts, phases, launches, _ = timed(dev_in, x_dev, u_dev, y_dev, z_dev, False, args.steps, args.warmup)
    clocks = sampler.stop()

    n_acc, n_cand = counts["n_accepted"], counts["n_candidates"]
    nnz = ctx.nnz
    tot = torch.tensor([n_acc, n_cand, nnz], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    tot_acc, tot_cand, tot_nnz = (int(v) for v in tot.tolist())
    value = tot_acc / (ms * 1e-3)

    # ---- roofline of the dominant kernel (this rank; CUDA events on the library's stream) ---------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback")
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
                "kernel_ms": dom_ms, "largest_kernel_by_time": dom,
                "kernel_ms_all": {k: v[0] for k, v in kernels.items()}}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            tj = json.load(open(traffic_file))
            roofline["traffic"] = tj.get("clip_kernel_dram_bytes_per_candidate", 0) * n_cand or None
            roofline["traffic_source"] = tj.get("source")
        except Exception:
            pass
    a_bytes = 12 * nnz + 16 * bg.n_dofs + 8 * im.n_dofs
    c_bytes = 12 * nnz + 16 * im.n_cells + 8 * bg.n_dofs
    spmv = {"Ct_vmult_GBs": a_bytes / (phases["spmv"] * 1e-3) / 1e9, "C_Tvmult_GBs": c_bytes / (phases["spmv_t"] * 1e-3) / 1e9,
            "Ct_vmult_frac_hbm": a_bytes / (phases["spmv"] * 1e-3) / 1e9 / peak, "C_Tvmult_frac_hbm": c_bytes / (phases["spmv_t"] * 1e-3) / 1e9 / peak,
            "nnz": nnz, "bytes_formula": "12*nnz + 16*n_rows + 8*n_cols", "note": "per rank, excludes the all-reduce"}

    # ---- end to end: host buffers in, matrix + vectors out ----------------------------------------
    e2e = None
    if not args.no_e2e:
        host_in = {k: v.cpu().pin_memory() for k, v in dev_in.items()}
        xh, uh = x_dev.cpu().pin_memory(), u_dev.cpu().pin_memory()
        yh = torch.empty(bg.n_dofs, dtype=torch.float64).pin_memory()
        zh = torch.empty(im.n_dofs, dtype=torch.float64).pin_memory()
        e_steps = max(1, min(args.steps, 3))
        ms_e, counts_e, _, _, out = timed(host_in, xh, uh, yh, zh, True, e_steps, 1)
        h2d = sum(v.numel() * v.element_size() for v in host_in.values()) + xh.numel() * 8 + uh.numel() * 8
        d2h = sum(t.numel() * t.element_size() for t in out) + yh.numel() * 8 + zh.numel() * 8
        e2e = {"value": tot_acc / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": ms_e, "steps": e_steps,
               "api": ("_lib.Context (C ABI): nm_set_background_boxes(host lo/hi + dofs + constraint lines) + nm_set_immersed + nm_intersect + "
                       "nm_build_sparsity + nm_assemble" if (args.e2e_separate_calls or args.verts) else
                       "_lib.Context (C ABI): nm_assemble_host (host lo/hi + dofs + constraint lines + immersed grid in one call, uploads "
                       "overlapped with the phases that do not need them yet)") +
                      " + nm_spmv x2 (host vectors) + nm_get_csr read-back into pinned host memory"}
        if args.e2e_two_in_flight and world == 1 and "lo" in host_in:
            e2e["two_steps_in_flight"] = two_in_flight(_lib, local, ctx, host_in, xh, uh, bg, im, d, e_steps, tot_acc)
        del host_in

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base, _, _ = cpu_sample(args.config, pick_ref_cycle(args), 1, 0)
        # the reference as shipped runs 1 MPI rank x 1 thread (code/scripts/run_all_tests.sh): same port, one thread, smaller sample
        one, _, _ = cpu_sample(args.config, max(0, pick_ref_cycle(args) - 2), 1, 0, threads=1)
        cpu_base["single_thread"] = {"value": one["value"], "unit": UNIT, "cores": 1, "sample": one["sample"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.config, cycle, world, info),
            "pairs": tot_acc, "candidates": tot_cand, "nnz": tot_nnz, "nnz_per_s": tot_nnz / (ms * 1e-3),
            "pairs_per_min": value * 60.0,
            "counts_rank0": counts, "phases_ms_rank0": phases, "roofline": roofline, "spmv": spmv,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "cpu_baseline": cpu_base,
            "device_bytes_rank0": ctx.device_bytes(), "gen_seconds": t_gen,
            "ct_vmult_reduction": ct_vmult_label(world, args, fused),
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def two_in_flight(_lib, local, ctx0, host_in, xh, uh, bg, im, d, steps, pairs_per_step):
    """End-to-end throughput for a STREAM of problems: two contexts, each driven by its own host thread through the same
    calls as the e2e leg (every step uploads its inputs from pinned memory and reads its matrix and vectors back), so the
    host->device traffic of one step overlaps the device->host traffic and the kernels of the other.  Wall clock between
    device synchronisations; reported next to `e2e`, never instead of it."""
    import threading
    import torch
    ctxs = [ctx0, _lib.Context(local)]
    outs = [{}, {}]

    def one(c, o):
        c.assemble_host(d, host_in["lo"], host_in["hi"], host_in["dofs"], bg.n_dofs, host_in["c_dof"], host_in["c_ptr"], host_in["c_master"],
                        host_in["c_weight"], im.dim, host_in["im_verts"], host_in["im_dofs"], im.n_dofs, 3, 1e-15)
        if "y" not in o:
            cap = int(c.nnz * 1.05) + 16
            o["y"] = torch.empty(bg.n_dofs, dtype=torch.float64).pin_memory()
            o["z"] = torch.empty(im.n_dofs, dtype=torch.float64).pin_memory()
            o["csr"] = (torch.empty(bg.n_dofs + 1, dtype=torch.int64).pin_memory(), torch.empty(cap, dtype=torch.int32).pin_memory(),
                        torch.empty(cap, dtype=torch.float64).pin_memory())
        c.spmv(xh, transpose=False, out=o["y"])
        c.spmv(uh, transpose=True, out=o["z"])
        c.csr(transpose=False, values=True, out=o["csr"])

    for c, o in zip(ctxs, outs):       # warm-up: allocations of both contexts
        one(c, o)
    torch.cuda.synchronize()
    errs = []

    def worker(i):
        try:
            for _ in range(steps):
                one(ctxs[i], outs[i])
        except Exception as e:  # surfaced after the join
            errs.append(e)

    t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if errs:
        raise errs[0]
    same = all(torch.equal(a, b) for a, b in zip(outs[0]["csr"], outs[1]["csr"])) and torch.equal(outs[0]["y"], outs[1]["y"])
    ctxs[1].close()
    return {"value": 2 * steps * pairs_per_step / dt, "unit": UNIT, "ms_per_step": dt / (2 * steps) * 1e3, "steps": 2 * steps,
            "contexts": 2, "results_identical": bool(same), "timing": "wall clock between device synchronisations"}


def ct_vmult_label(world, args, fused):
    """How the partial Ct.vmult results of the ranks were combined in the timed step."""
    if world == 1:
        return "none (1 GPU)"
    why = " (fused path unavailable: %s)" % fused["why"] if fused["why"] else ""
    if args.ct_vmult == "owned":
        if fused["op"] is not None:
            return "result distributed by dof ownership: nm_spmv_push_owned over NVLink (red.add into the owner's peer-mapped vector)"
        return "result distributed by dof ownership: nccl reduce_scatter" + why
    if fused["op"] is not None:
        return "replicated result: nm_spmv_push over NVLink, %s" % ("multimem.red on the multicast address" if fused["op"].multicast
                                                                    else "red.add into peer-mapped vectors")
    return "replicated result: nccl all_reduce" + why


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
