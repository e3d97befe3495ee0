"""The C-ABI library loads and exports every symbol include/nmgpu.h declares (no compute without a GPU),
and the product fails loudly instead of falling back when no CUDA device is present."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "nmgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(nm_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_are_exported():
    from non_matching_test_suite_b200 import _lib
    L = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_lib.SYMBOLS) == names
    assert L.nm_version() >= 100


def test_no_cpu_fallback():
    from non_matching_test_suite_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.NmError) as e:
        _lib.Context(0)
    assert e.value.code == 2  # NM_ERR_CUDA


def test_product_does_not_touch_the_oracle():
    """Nothing under the package (or its C sources) may import, link or execute oracle/."""
    pkg = os.path.join(ROOT, "non_matching_test_suite_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_obj" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.replace("exact_oracle.py", "").lower() or f in (), (dirpath, f)


def test_wrapper_argument_checks_mirror_the_reference():
    from non_matching_test_suite_b200 import coupling
    from non_matching_test_suite_b200.grids import make_config
    bg, im = make_config("disk", 0, band_only=False)
    with pytest.raises(ValueError, match="Invalid quadrature degree"):
        coupling.collect_quadratures_on_overlapped_grids(bg, im, 0)


def test_struct_layouts_match_the_header(tmp_path):
    """nm_counts and nm_host_problem as a C compiler lays them out against the ctypes mirrors in _lib.py."""
    import subprocess
    from non_matching_test_suite_b200 import _lib
    src = tmp_path / "layout.c"
    fields = [n for n, _ in _lib.HostProblem._fields_]
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "nmgpu.h"\nint main(void){\n'
                   'printf("%zu %zu\\n", sizeof(nm_counts), sizeof(nm_host_problem));\n' +
                   "".join('printf("%%zu\\n", offsetof(nm_host_problem, %s));\n' % f for f in fields) + "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert int(out[0]) == ctypes.sizeof(_lib.Counts) and int(out[1]) == ctypes.sizeof(_lib.HostProblem)
    for f, off in zip(fields, out[2:]):
        assert getattr(_lib.HostProblem, f).offset == int(off), f
