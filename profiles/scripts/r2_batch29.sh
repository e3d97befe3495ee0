#!/bin/bash
# merge: pair-by-pair accumulation + one-key rank passes with a ballot tail (default), without the pair steps (_ps0),
# without the ballot tail (_rt0): GPU tests that touch the merge on the default library, then the benchmark phases
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_graded.py tests/test_gpu_fe.py tests/test_gpu_codim0.py -q -m gpu 2>&1 | tail -4
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
for v in "" _ps0 _rt0; do
  NMGPU_LIB=$PWD/non_matching_test_suite_b200/libnmgpu$v.so timeout 300 python bench.py $B > $O/r2ak_bench$v.json 2> $O/r2ak_bench$v.err
  echo "variant '$v' rc $?"
  python - "$O/r2ak_bench$v.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
p=d["phases_ms_rank0"]; print("  step %.2f merge %.3f transpose %.3f spmv %.3f %.3f mem %.1f GB"%(d["ms_per_step"],p["merge"],p["transpose"],p["spmv"],p["spmv_t"],d["device_bytes_rank0"]/1e9))
PY
done
