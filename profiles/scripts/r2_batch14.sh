#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
for v in "" _g16 _s128; do
  NMGPU_LIB=$PWD/non_matching_test_suite_b200/libnmgpu$v.so timeout 300 python bench.py $B > $O/r2p_bench$v.json 2> $O/r2p_bench$v.err
  echo "variant '$v' rc $?"
  python - "$O/r2p_bench$v.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
p=d["phases_ms_rank0"]; print("  step %.2f clip %.3f merge %.2f k_merge %.3f pairs %d nnz %d"%(d["ms_per_step"],p["k_clip"],p["merge"],p["k_merge"],d["pairs"],d["nnz"]))
PY
done
NMGPU_LIB=$PWD/non_matching_test_suite_b200/libnmgpu_g16.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x > $O/r2p_pytest_g16.log 2>&1; echo "pytest g16 rc $?"; tail -2 $O/r2p_pytest_g16.log
