#!/bin/bash
# SpMV padding slots masked by 32-bit predicates (default) against the unmasked kernel (_pad0): regression tests on the
# default library, then the product timings of both builds
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_fe.py -q -m gpu 2>&1 | tail -6
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-sub-configs --no-dropin --no-check"
for v in "" _pad0; do
NMGPU_LIB=$PWD/non_matching_test_suite_b200/libnmgpu$v.so timeout 300 python bench.py $B > $O/r2ao_bench$v.json 2> $O/r2ao_bench$v.err; echo "bench '$v' rc $?"
python - "$O/r2ao_bench$v.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
p=d["phases_ms_rank0"]; print("  step %.2f spmv %.3f spmv_t %.3f"%(d["ms_per_step"],p["spmv"],p["spmv_t"]), d["spmv"]["Ct_vmult_frac_hbm"], d["spmv"]["C_Tvmult_frac_hbm"])
PY
done
