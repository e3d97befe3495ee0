#!/bin/bash
# warp-aggregated brick insertion (default) against one claim + one OR per cell (_agg0)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_graded.py tests/test_gpu_codim0.py -q -m gpu 2>&1 | tail -5
B="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-dropin --no-check"
for v in "" _agg0; do
NMGPU_LIB=$PWD/non_matching_test_suite_b200/libnmgpu$v.so timeout 300 python bench.py $B > $O/r2aq_bench$v.json 2> $O/r2aq_bench$v.err; echo "bench '$v' rc $?"
python - "$O/r2aq_bench$v.json" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
p=d["phases_ms_rank0"]; print("  step %.2f index %.3f cand %.3f"%(d["ms_per_step"],p["index"],p["candidates"]), {k:(round(v["ms_per_step"],3), round(v["phases_ms_rank0"]["index"],3)) for k,v in d.get("configs",{}).items()})
PY
done
