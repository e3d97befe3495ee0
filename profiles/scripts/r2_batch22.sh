#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
for v in "" _ts6 _ts8; do
  NMGPU_LIB=$PWD/non_matching_test_suite_b200/libnmgpu$v.so timeout 300 python bench.py $B > $O/r2ab_bench$v.json 2> $O/r2ab_bench$v.err
  echo "variant '$v' rc $?"
  python - "$O/r2ab_bench$v.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
p=d["phases_ms_rank0"]; print("  step %.2f index %.3f cand %.3f merge %.3f transpose %.3f"%(d["ms_per_step"],p["index"],p["candidates"],p["merge"],p["transpose"]))
PY
done
