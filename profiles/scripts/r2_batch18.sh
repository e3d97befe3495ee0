#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
for v in "" _w4m6 _w4m7 _w2m12 _w16m1; do
  NMGPU_LIB=$PWD/non_matching_test_suite_b200/libnmgpu$v.so timeout 300 python bench.py $B > $O/r2t_bench$v.json 2> $O/r2t_bench$v.err
  echo "variant '$v' rc $?"
  python - "$O/r2t_bench$v.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
p=d["phases_ms_rank0"]; print("  step %.2f clip %.3f merge %.2f"%(d["ms_per_step"],p["k_clip"],p["merge"]))
PY
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "error_behaviour" > $O/r2t_pytest.log 2>&1; echo "pytest rc $?"; tail -2 $O/r2t_pytest.log
