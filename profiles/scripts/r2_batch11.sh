#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
for v in "" _pslot _sp _spaxl; do
  NMGPU_LIB=$PWD/non_matching_test_suite_b200/libnmgpu$v.so timeout 300 python bench.py $B > $O/r2n_bench$v.json 2> $O/r2n_bench$v.err
  echo "variant '$v' rc $?"
  python - "$O/r2n_bench$v.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
p=d["phases_ms_rank0"]; print("  step %.2f clip %.3f merge %.2f pairs %d"%(d["ms_per_step"],p["k_clip"],p["merge"],d["pairs"]))
PY
done
A6="--cycle 6 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
export NMGPU_LIB=$PWD/non_matching_test_suite_b200/libnmgpu_sp.so
timeout 600 python bench.py $A6 > $O/plain6.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'clip_moments3' -s 1 -c 1 -o $O/r2n_clip6_sp -f python bench.py $A6 > $O/ncu_clip6.log 2>&1; echo "prof rc $?"
