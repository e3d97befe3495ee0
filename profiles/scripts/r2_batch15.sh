#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -x > $O/r2q_pytest.log 2>&1; echo "pytest rc $?"; tail -3 $O/r2q_pytest.log
NMGPU_BENCH_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sub-configs --no-dropin > $O/r2q_bench.json 2> $O/r2q_bench.err; echo "bench rc $?"
grep -a "e2e step" $O/r2q_bench.err | tail -6
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2q_bench.json") if l.startswith("{")][-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["e2e"]["steps"])
PY
