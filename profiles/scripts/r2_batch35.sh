#!/bin/bash
# warp-cooperative row sort of the transpose (default) against the thread-per-row kernel (_tc0)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_graded.py tests/test_gpu_codim0.py tests/test_gpu_fe.py tests/test_gpu_nitsche.py -q -m gpu 2>&1 | tail -5
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-dropin --no-check"
for v in "" _tc0; do
NMGPU_LIB=$PWD/non_matching_test_suite_b200/libnmgpu$v.so timeout 300 python bench.py $B > $O/r2ar_bench$v.json 2> $O/r2ar_bench$v.err; echo "bench '$v' rc $?"
python - "$O/r2ar_bench$v.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
p=d["phases_ms_rank0"]; print("  step %.2f transpose %.3f"%(d["ms_per_step"],p["transpose"]), {k:(round(v["ms_per_step"],3), round(v["phases_ms"]["transpose"],3)) for k,v in d.get("configs",{}).items()})
PY
done
