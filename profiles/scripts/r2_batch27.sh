#!/bin/bash
# merge with the DoF rows of a chunk staged in shared memory (default), with only the background cell ids staged (_cbg1),
# and with neither (_cbg0): parity tests on the default library, then the phases of the benchmark step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_graded.py -x -q -m gpu 2>&1 | tail -4
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
for v in "" _cbg1 _cbg0; do
  NMGPU_LIB=$PWD/non_matching_test_suite_b200/libnmgpu$v.so timeout 300 python bench.py $B > $O/r2ai_bench$v.json 2> $O/r2ai_bench$v.err
  echo "variant '$v' rc $?"
  python - "$O/r2ai_bench$v.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
p=d["phases_ms_rank0"]; print("  step %.2f merge %.3f transpose %.3f spmv %.3f %.3f mem %.1f GB"%(d["ms_per_step"],p["merge"],p["transpose"],p["spmv"],p["spmv_t"],d["device_bytes_rank0"]/1e9))
PY
done
