"""bench.py contract on the CPU side: the reference arm (`--impl reference`) needs no GPU, prints exactly one JSON
line with the keys the driver reads, and ranks other than 0 stay silent under torchrun-style environments."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(extra_env=None, *flags):
    env = dict(os.environ, **(extra_env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", *flags],
                          capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = _run(None, "--ref-cycle", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d)
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["value"] > 0 and d["unit"] == "pairs/s" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2", "--ref-cycle", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_committed_gpu_lines_carry_the_contract():
    """The lines of the round that DESIGN.md quotes (profiles/r2_bench_n*.json, written on the B200 boxes) hold every key
    of the bench contract, a roofline that is consistent with itself and the clocks sampled during the timed region."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_bench_n[1248].json")))
    assert len(files) == 4
    for f in files:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        assert (KEYS - {"impl"}) | {"clocks", "gpu_launches", "roofline"} <= set(d), f
        assert d["unit"] == "pairs/s" and d["dtype"] == "f64" and d["vs_baseline"] is None and d["scaling"] == "weak"
        assert d["warmup"] >= 2 and d["steps"] >= 1 and d["gpu_launches"] > 0 and "workload" in d["config"]
        assert abs(d["value"] - d["pairs"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert r["traffic"] is None or r["traffic"] > 0
        e = d["e2e"]
        assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
        c = d["clocks"]
        assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if d["n_gpus"] == 1:
            cb = d["cpu_baseline"]
            assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
        else:
            assert d["ct_vmult_check"]["max_rel_err"] <= 1e-12 and d["shard_check"]["bitwise_identical_values"] is True
