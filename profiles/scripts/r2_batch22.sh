#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
for v in "" _mb6 _mb7 _mb8; do
  NMGPU_LIB=$PWD/non_matching_test_suite_b200/libnmgpu$v.so timeout 300 python bench.py $B > $O/r2aa_bench$v.json 2> $O/r2aa_bench$v.err
  echo "variant '$v' rc $?"
  python - "$O/r2aa_bench$v.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
p=d["phases_ms_rank0"]; print("  step %.2f clip %.3f merge %.2f k_merge %.3f"%(d["ms_per_step"],p["k_clip"],p["merge"],p["k_merge"]))
PY
done
