"""Host-side unit test of the brick-index helpers the candidate kernels are built from
(non_matching_test_suite_b200/csrc/nm_brick.cuh): exact per-axis index ranges, per-brick query masks and a whole query
on a two-level band of grid cells, each against brute force on the same doubles (closed boxes, touching counts:
reference code/src/coupling_utilities.cc:53-55)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_brick_helpers_against_brute_force():
    build = os.path.join(ROOT, "tests", "cpp", "_build")
    os.makedirs(build, exist_ok=True)
    exe = os.path.join(build, "test_brick")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_brick.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "test_brick: ok" in r.stdout
