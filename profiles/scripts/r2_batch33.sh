#!/bin/bash
# pair-by-pair accumulation also in the kernel that looks the lines up (rounds none of whose dofs is constrained):
# tests on the default library, then merge times on the plain and the graded band for the default build and _ps0
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_graded.py tests/test_gpu_nitsche.py -q -m gpu 2>&1 | tail -5
for v in "" _ps0; do
NMGPU_LIB=$PWD/non_matching_test_suite_b200/libnmgpu$v.so timeout 600 python profiles/graded_timing.py 6 1 > $O/r2ap_graded$v.json 2> $O/r2ap_graded$v.err; echo "timing '$v' rc $?"
python - "$O/r2ap_graded$v.json" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
for k,v in d.items():
    if isinstance(v,dict) and "merge_ms" in v: print("  ",k,{q:v[q] for q in v if q.endswith("_ms") or q in ("pairs","constraint_lines")})
PY
done
