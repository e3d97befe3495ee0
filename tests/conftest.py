import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.dirname(os.path.abspath(__file__))
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on the CPU")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: on a CPU-only box they are skipped instead of failing."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import c_oracle as co
    co.build()
    return co


@pytest.fixture(scope="session", autouse=True)
def _native_library():
    """The tests load non_matching_test_suite_b200/libnmgpu.so; on a fresh checkout (the .so is not in the history)
    build it once with the repo's own Makefile (nvcc cross-compiles sm_100a without a GPU)."""
    lib = os.path.join(ROOT, "non_matching_test_suite_b200", "libnmgpu.so")
    if not os.path.exists(lib):
        import shutil
        import subprocess
        if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "non_matching_test_suite_b200", "csrc")])
        # without nvcc the tests that load the library fail on their own; the pure-Python oracle tests still run
    yield
